#!/usr/bin/env python
"""TEST TOOLING — NOT PRODUCT CODE. Dry run of bench.py's single-GPU arm on CPU: torch.cuda is replaced by no-op stand-ins
and khronos_b200.lib() by the CUDA-on-CPU build (tools/cuda_emu), so the whole measurement script (descriptor building,
batched calls, event sampling, e2e windows, CPU baseline, JSON assembly) is executed end to end on a tiny workload.
The numbers it prints mean nothing; it exists to catch host-side bugs in bench.py where no GPU is at hand.
    python tools/cuda_emu/run_bench_dry.py [bench.py args, default: --small --steps 2 --warmup 1 ...]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 tools/cuda_emu/run_bench_dry.py \
        --gpus 2 --small ...          (the N > 1 arms over gloo instead of NCCL)"""
import contextlib
import ctypes
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main(argv):
    import build_emu
    lib = ctypes.CDLL(build_emu.build())
    import torch
    import khronos_b200 as kb
    kb._LIB = lib
    torch.set_num_threads(1)
    real_device = torch.device

    class FakeStream:
        cuda_stream = 0
        def __init__(self, *a, **k): pass
        def wait_stream(self, *a): pass
        def wait_event(self, *a): pass

    class FakeEvent:
        def __init__(self, *a, **k): self.t = None
        def record(self, *a): self.t = time.perf_counter()
        def elapsed_time(self, other): return (other.t - self.t) * 1e3 + 1e-3

    def fake_device(kind, *a):
        return real_device("cpu")

    torch.device = fake_device
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Stream = FakeStream
    torch.cuda.Event = FakeEvent
    torch.cuda.current_stream = lambda *a, **k: FakeStream()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    real_empty, real_zeros, real_tensor = torch.empty, torch.zeros, torch.tensor
    torch.empty = lambda *a, pin_memory=False, **k: real_empty(*a, **k)
    torch.zeros = lambda *a, pin_memory=False, **k: real_zeros(*a, **k)
    import torch.distributed as dist
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend=None, **k: real_init("gloo", **{a: b for a, b in k.items() if a != "device_id"})
    sys.argv = ["bench.py"] + (list(argv) or ["--small", "--steps", "2", "--warmup", "1", "--lap-frames", "48", "--frames-per-step", "16",
                                              "--batch", "8", "--e2e-frames", "16", "--cpu-sample-frames", "12", "--cpu-sample-seconds", "1"])
    import bench
    bench.main()


if __name__ == "__main__":
    main(sys.argv[1:])
