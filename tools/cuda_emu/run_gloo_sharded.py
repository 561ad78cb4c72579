#!/usr/bin/env python
"""TEST TOOLING — NOT PRODUCT CODE. World-size-2 gloo run of the sharded per-frame pipeline with the *product's*
kernels (CUDA-on-CPU build, tools/cuda_emu) on every rank: khronos_b200.distributed.ShardedActiveWindow + DistComm
(pixel-flag all-reduce, pending / halo all-gathers) exactly as the NCCL test drives it on GPUs; rank 0 compares the
union of the shards and the dynamic images with the unsharded oracle.   python tools/cuda_emu/run_gloo_sharded.py"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def worker(rank, world, port, q, libpath):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    from khronos_b200 import capi, distributed as kd
    import harness as hs
    import test_sharded_pipeline as tsp
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        emu = ctypes.CDLL(libpath)
        cam = hs.small_camera(8)
        n = 20
        frames, poses, stamps = tsp.dynamic_scenario(cam, n)
        mot = capi.default_motion_config(min_cluster_size=5, min_separation_distance=2.0, num_threads=2)
        h = hs.make_handle(emu, "kb_", cam=cam, mot_cfg=mot)
        h.set_shard(rank, world)
        win = kd.ShardedActiveWindow([h], kd.DistComm(world), device="cpu")
        ref = None
        if rank == 0:
            ref = hs.make_handle(ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so")), "ko_", cam=cam, mot_cfg=mot)
        ok, dyn = True, 0
        for i in range(n):
            d, l = frames[i]
            (img, ns, nc), = win.spin_once([h.make_frame(d, poses[i], stamps[i], label=l)])
            if ref is not None:
                img_o, ns_o, nc_o = ref.spin_once(ref.make_frame(d, poses[i], stamps[i], label=l))
                ok = ok and (ns, nc) == (ns_o, nc_o) and bool((img == img_o).all())
                dyn += int((img_o > 0).sum())
        gathered = [None] * world
        dist.all_gather_object(gathered, h.export_blocks())
        if rank == 0:
            msg = "ok" if ok and dyn > 30 else "dynamic image / counts differ (or no motion)"
            try:
                tsp.assert_union_equals(gathered, ref.export_blocks(), "gloo product shards")
            except AssertionError as e:
                msg = "mismatch: " + str(e)[:300]
            q.put((msg, [g.n for g in gathered], ref.export_blocks().n, dyn))
    except Exception as e:
        q.put(("error on rank %d: %r" % (rank, e), [], 0, 0))
        os._exit(1)
    dist.destroy_process_group()


def main():
    sys.path.insert(0, HERE)
    import build_emu
    libpath = build_emu.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=worker, args=(r, 2, port, q, libpath), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = q.get(timeout=900)
    finally:
        for p in procs:
            p.join(timeout=30)
        for p in procs:
            if p.is_alive():
                p.kill()
    print(res)
    return 0 if res[0] == "ok" else 1


if __name__ == "__main__":
    sys.exit(main())
