#!/bin/bash
# Multi-GPU measurements round 1 left open. Run with N GPUs of one box:
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_checklist_multi.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "== 1. NCCL tests (sharded fusion + sharded per-frame pipeline)"
timeout 400 python -m pytest tests/test_multigpu_nccl.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/nccl_tests.log 2>&1; tail -4 gpurun_out/nccl_tests.log
echo "== 2. fusion-only bench, f32 wire vs lossless 5 B/px wire"
for w in f32 f32u8; do
  timeout 300 $TR bench.py --gpus $N --wire $w > gpurun_out/bench_n${N}_$w.json 2> gpurun_out/bench_n${N}_$w.err
  python -c "import json;d=json.load(open('gpurun_out/bench_n${N}_$w.json'));print('$w', round(d['value']), 'fps', d['ms_per_step'], 'ms/step')" || tail -3 gpurun_out/bench_n${N}_$w.err
done
echo "== 2b. is the frame broadcast channel-limited? (the f32 wire is broadcast bound at ~280 GB/s)"
for ch in 16 32; do
  NCCL_MIN_NCHANNELS=$ch timeout 300 $TR bench.py --gpus $N --wire f32 > gpurun_out/bench_n${N}_f32_ch$ch.json 2> gpurun_out/bench_n${N}_f32_ch$ch.err
  python -c "import json;d=json.load(open('gpurun_out/bench_n${N}_f32_ch$ch.json'));print('NCCL_MIN_NCHANNELS=$ch', round(d['value']), 'fps')" || true
done
echo "== 2c. NVLS multicast frame broadcast (kb_multicast_copy into a symmetric buffer; first time on hardware; smaller steps)"
for w in f32 f32u8; do
  timeout 300 $TR bench.py --gpus $N --wire $w --bcast multimem --frames-per-step 1000 > gpurun_out/bench_n${N}_${w}_mm.json 2> gpurun_out/bench_n${N}_${w}_mm.err
  python -c "import json;d=json.load(open('gpurun_out/bench_n${N}_${w}_mm.json'));print('multimem $w', round(d['value']), 'fps')" || tail -4 gpurun_out/bench_n${N}_${w}_mm.err
  timeout 300 $TR bench.py --gpus $N --wire $w --frames-per-step 1000 > gpurun_out/bench_n${N}_${w}_f1000.json 2> gpurun_out/bench_n${N}_${w}_f1000.err
  python -c "import json;d=json.load(open('gpurun_out/bench_n${N}_${w}_f1000.json'));print('nccl     $w', round(d['value']), 'fps (same step size)')" || true
done
echo "== 3. per-frame pipeline (config[2]) sharded over $N GPUs"
timeout 300 $TR bench.py --gpus $N --workload dynamic --steps 4 --warmup 2 > gpurun_out/bench_n${N}_dynamic.json 2> gpurun_out/bench_n${N}_dynamic.err
tail -c 700 gpurun_out/bench_n${N}_dynamic.json || tail -3 gpurun_out/bench_n${N}_dynamic.err
echo "== 3b. the same with the peer-memory exchange (symmetric memory; first time on hardware)"
timeout 300 $TR bench.py --gpus $N --workload dynamic --steps 4 --warmup 2 --exchange peers > gpurun_out/bench_n${N}_dynamic_peers.json 2> gpurun_out/bench_n${N}_dynamic_peers.err
tail -c 400 gpurun_out/bench_n${N}_dynamic_peers.json || tail -5 gpurun_out/bench_n${N}_dynamic_peers.err
