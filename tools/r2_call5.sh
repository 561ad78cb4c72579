#!/bin/bash
# round 2, call 5 (1 GPU): tests touched since call 4, N=1 bench (whole-step calls, MLE update, vectorised tile max),
# ncu full captures of the prologue kernels
set -u
O=gpurun_out/r2c5
mkdir -p $O
echo "== gpu tests (subset)"
timeout 900 python -m pytest tests/test_bench_shape_parity.py tests/test_instance_forwarding.py tests/test_parity_gpu.py tests/test_zzz_fuse_variants.py tests/test_golden.py -m gpu -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "rc=$?"; tail -6 $O/gpu_tests.log
run() { n=$1; shift; env "$@" timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-legs > $O/bench_$n.json 2> $O/bench_$n.err; python -c "
import json;d=json.load(open('$O/bench_$n.json'));print('$n',round(d['value']),'fps',d['checksum']['sum'],round(d['roofline']['launch_us'],1),'us/group', d['roofline'].get('launch_us_sampled'))" || tail -3 $O/bench_$n.err; }
run default A=1
run batch32 A=1 --batch 32
K='regex:fuseKernel|selectBlocks|itemCull|itemCompact|tileMax|tilePyramid'
A="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-legs"
echo "== ncu launch list + metrics (second lap)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k "$K" -s 1200 -c 360 --csv --log-file $O/launches_metrics.csv python bench.py $A > $O/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full capture: selectBlocks, itemCull, tileMax (2 launches each), fuse (3)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:selectBlocks -s 220 -c 2 -o $O/select_full python bench.py $A > $O/ncu_select.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:itemCull -s 220 -c 2 -o $O/cull_full python bench.py $A > $O/ncu_cull.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fuseKernel -s 220 -c 3 -o $O/fuse_full python bench.py $A > $O/ncu_fuse.log 2>&1; echo "rc=$?"
ls -la $O
