#!/bin/bash
# round 2, call 1: A/Bs of the prepared variants on hardware (sections 2, 2b, 3 of gpu_checklist.sh) + detector timing
set -u
mkdir -p gpurun_out/ab1
O=gpurun_out/ab1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
run() { n=$1; shift; env "$@" timeout 200 python bench.py --no-e2e --no-cpu-baseline > $O/bench_$n.json 2> $O/bench_$n.err; }
run default A=1
run item_list KB_FUSE_ITEM_LIST=1
run mlp2 KB_FUSE_MLP=2
run mlp4 KB_FUSE_MLP=4
run mlp2_list KB_FUSE_MLP=2 KB_FUSE_ITEM_LIST=1
run pipe10 KB_PIPELINE=1 KB_FUSE_CTAS_PER_SM=10
run pipe8 KB_PIPELINE=1 KB_FUSE_CTAS_PER_SM=8
run pipe8_list KB_PIPELINE=1 KB_FUSE_CTAS_PER_SM=8 KB_FUSE_ITEM_LIST=1
run default2 A=1
python - <<'PY'
import json
for n in ("default", "item_list", "mlp2", "mlp4", "mlp2_list", "pipe10", "pipe8", "pipe8_list", "default2"):
    try:
        d = json.load(open(f"gpurun_out/ab1/bench_{n}.json"))
        print(n, round(d["value"]), "fps", d["roofline"]["launch_us"], "us/launch", d["clocks"])
    except Exception as e:
        print(n, "failed:", e)
PY
echo "== e2e"
timeout 300 python bench.py --no-cpu-baseline > $O/bench_e2e_default.json 2> $O/bench_e2e_default.err
KB_H2D_NARROW_LABELS=1 KB_H2D_THREADS=16 timeout 300 python bench.py --no-cpu-baseline > $O/bench_e2e_narrow16.json 2> $O/bench_e2e_narrow16.err
python - <<'PY'
import json
for n in ("default", "narrow16"):
    try:
        d = json.load(open(f"gpurun_out/ab1/bench_e2e_{n}.json"))
        print("e2e", n, round(d["e2e"]["value"]), "fps")
    except Exception as e:
        print("e2e", n, "failed:", e)
PY
echo "== dynamic"
dyn() { n=$1; shift; env "$@" timeout 200 python bench.py --workload dynamic --steps 4 --warmup 2 --no-cpu-baseline $EXTRA > $O/dyn_$n.json 2> $O/dyn_$n.err; python -c "import json;d=json.load(open('$O/dyn_$n.json'));print('dynamic $n', round(d['value']), 'fps')"; }
EXTRA="" dyn default A=1
EXTRA="--force-cull" dyn cull A=1
EXTRA="" dyn efv2 KB_EVERFREE_V2=1
EXTRA="" dyn sparse KB_MOTION_SPARSE=1
EXTRA="" dyn both KB_EVERFREE_V2=1 KB_MOTION_SPARSE=1
timeout 200 python tools/time_detectors.py > $O/time_detectors.log 2>&1; tail -4 $O/time_detectors.log
