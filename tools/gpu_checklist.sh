#!/bin/bash
# One gpurun call that covers everything round 1 left unmeasured (run from the repo root on a B200 box):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_checklist.sh'
# Outputs land in gpurun_out/ (copy the summaries worth keeping to profiles/).
set -u
mkdir -p gpurun_out
echo "== 1. GPU tests not yet run on hardware (object detection, experimental fuse variants)"
timeout 300 python -m pytest tests/test_zz_edge_cases.py tests/test_zz_object_detection.py tests/test_zzy_peer_exchange.py tests/test_zzz_fuse_variants.py -m gpu -q --tb=short -p no:cacheprovider \
  > gpurun_out/zz_tests.log 2>&1; tail -5 gpurun_out/zz_tests.log
echo "== 2. headline bench, default vs compacted heaviest-first item lists"
timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
KB_FUSE_ITEM_LIST=1 timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_item_list.json 2> gpurun_out/bench_item_list.err
KB_FUSE_MLP=2 timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_mlp2.json 2> gpurun_out/bench_mlp2.err
KB_FUSE_MLP=4 timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_mlp4.json 2> gpurun_out/bench_mlp4.err
KB_FUSE_MLP=4 KB_FUSE_ITEM_LIST=1 timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_mlp4_list.json 2> gpurun_out/bench_mlp4_list.err
for c in 10 8 6; do
  KB_PIPELINE=1 KB_FUSE_CTAS_PER_SM=$c timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_pipe$c.json 2> gpurun_out/bench_pipe$c.err
done
KB_PIPELINE=1 KB_FUSE_CTAS_PER_SM=4 KB_FUSE_MLP=4 timeout 200 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_pipe_mlp4.json 2> gpurun_out/bench_pipe_mlp4.err
python - <<'PY'
import json
for n in ("default", "item_list", "mlp2", "mlp4", "mlp4_list", "pipe10", "pipe8", "pipe6", "pipe_mlp4"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, round(d["value"]), "fps", d["roofline"]["launch_us"], "us/launch", d["clocks"])
    except Exception as e:
        print(n, "failed:", e)
PY
echo "== 2b. e2e (host f32 + i32 frames through the C ABI): default vs labels narrowed to u8 on the host (5 instead of 8 B/px over PCIe)"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_e2e_default.json 2> gpurun_out/bench_e2e_default.err
for t in 8 16 32; do
  KB_H2D_NARROW_LABELS=1 KB_H2D_THREADS=$t timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_e2e_narrow$t.json 2> gpurun_out/bench_e2e_narrow$t.err
done
python - <<'PY'
import json
for n in ("default", "narrow8", "narrow16", "narrow32"):
    try:
        d = json.load(open(f"gpurun_out/bench_e2e_{n}.json"))
        print("e2e", n, round(d["e2e"]["value"]), "fps")
    except Exception as e:
        print("e2e", n, "failed:", e)
PY
echo "== 3. per-frame pipeline (config[2]) on one GPU"
timeout 200 python bench.py --workload dynamic --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_dynamic.json 2> gpurun_out/bench_dynamic.err
tail -c 600 gpurun_out/bench_dynamic.json
timeout 200 python bench.py --workload dynamic --steps 4 --warmup 2 --no-cpu-baseline --force-cull > gpurun_out/bench_dynamic_cull.json 2> gpurun_out/bench_dynamic_cull.err
KB_EVERFREE_V2=1 timeout 200 python bench.py --workload dynamic --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_dynamic_efv2.json 2> gpurun_out/bench_dynamic_efv2.err
python -c "import json;print('dynamic with KB_EVERFREE_V2', round(json.load(open('gpurun_out/bench_dynamic_efv2.json'))['value']), 'fps')"
KB_MOTION_SPARSE=1 timeout 200 python bench.py --workload dynamic --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_dynamic_sparse.json 2> gpurun_out/bench_dynamic_sparse.err
python -c "import json;print('dynamic with KB_MOTION_SPARSE', round(json.load(open('gpurun_out/bench_dynamic_sparse.json'))['value']), 'fps')"
python -c "import json;print('dynamic', round(json.load(open('gpurun_out/bench_dynamic.json'))['value']), 'fps; with single-frame culling', round(json.load(open('gpurun_out/bench_dynamic_cull.json'))['value']), 'fps')"
echo "== 4. launch list of the sharded / object kernels (2 shards on one device)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_sharded.csv \
  -k regex:everFree\|halo\|ghost\|exportPending\|motionFinalize\|motionLookup\|os.*Kernel\|o2.*Kernel \
  python -m pytest "tests/test_sharded_pipeline.py::test_product_shards_equal_unsharded_oracle[2-2.0]" tests/test_zz_object_detection.py -m gpu -q -x -p no:cacheprovider \
  > gpurun_out/ncu_sharded.log 2>&1; tail -3 gpurun_out/ncu_sharded.log
echo "== 5. track measurements (written after the round's GPU minutes were spent): parity on hardware + latency of the detector / tracker calls"
timeout 300 python -m pytest tests/test_zz_track_measurements.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/zz_tracks.log 2>&1; tail -3 gpurun_out/zz_tracks.log
timeout 200 python tools/time_detectors.py > gpurun_out/time_detectors.log 2>&1; tail -2 gpurun_out/time_detectors.log
echo "== 6. ray index (RayVerificator on the device; written without GPU minutes): parity on hardware"
timeout 300 python -m pytest tests/test_zz_ray_index.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/zz_rays.log 2>&1; tail -3 gpurun_out/zz_rays.log
