#!/bin/bash
# round 2, call 2 (1 GPU): full -m gpu suite with the new defaults, headline bench, ncu launch list + metrics + full capture
set -u
O=gpurun_out/r2c2
mkdir -p $O
echo "== gpu tests"
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "rc=$?"; tail -4 $O/gpu_tests.log
echo "== bench N=1"
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_n1.json'));print(round(d['value']),'fps e2e',d['e2e'] and round(d['e2e']['value']),'cpu',d['cpu_baseline'] and round(d['cpu_baseline']['value']),d['checksum'],d['roofline']['frac'],d['roofline']['launch_us'],d['clocks'])"
K='regex:fuseKernel|selectBlocks|itemCull|itemCompact|tileMax|tilePyramid'
A="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
echo "== ncu launch list (second lap: map allocated, steady state)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k "$K" -s 1200 -c 360 --csv --log-file $O/launches_metrics.csv python bench.py $A > $O/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full capture of the fuse kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fuseKernel -s 220 -c 3 -o $O/fuse_full python bench.py $A > $O/ncu_full.log 2>&1; echo "rc=$?"
ls -la $O
