#!/bin/bash
# round 2, call 6 (1 GPU): final validation: whole -m gpu suite, default bench with all legs, ncu launch metrics + full
# captures exported to CSV on the box (the .ncu-rep files stay there: gpurun_out is limited to 64 MiB)
set -u
O=gpurun_out/r2c6
mkdir -p $O
echo "== gpu tests (all)"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "rc=$?"; tail -6 $O/gpu_tests.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log
echo "== full default bench"
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_full.json'));print(round(d['value']),'fps; e2e',d['e2e'] and round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']), d['checksum'], 'frac', d['roofline']['frac'], d['clocks']);print('tick',d['output_tick']['mesh_tick_ms'], d['output_tick']['mirror_back_tick_ms']);print('dynamic',d['configs']['dynamic']['value'], d['configs']['dynamic']['flagged_pixel_fraction_last_frame'], d['configs']['dynamic'].get('cpu_baseline'))" || tail -5 $O/bench_full.err
echo "== CTA-cooperative fuse kernel (KB_FUSE_COOP=1): parity + A/B"
KB_FUSE_COOP=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_bench_shape_parity.py tests/test_golden.py -m gpu -q -p no:cacheprovider > $O/gpu_tests_coop.log 2>&1; echo "rc=$?"; tail -4 $O/gpu_tests_coop.log
for v in 1 0; do KB_FUSE_COOP=$v timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-legs > $O/bench_coop$v.json 2> $O/bench_coop$v.err; python -c "
import json;d=json.load(open('$O/bench_coop$v.json'));print('coop=$v',round(d['value']),'fps',d['checksum']['sum'],round(d['roofline']['launch_us'],1),'us/group')" || tail -3 $O/bench_coop$v.err; done
echo "== reference arm"
timeout 300 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; head -c 300 $O/bench_ref.json; echo
K='regex:fuseKernel|selectBlocks|itemCull|itemCompact|tileMax|tilePyramid'
A="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-legs"
echo "== ncu launch list + metrics (second lap)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k "$K" -s 1200 -c 360 --csv --log-file $O/launches_metrics.csv python bench.py $A > $O/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full captures -> CSV"
for k in fuseKernel selectBlocks itemCull; do
  timeout 500 ncu --set full --clock-control none -k regex:$k -s 220 -c 2 -o /tmp/${k}_full python bench.py $A > $O/ncu_$k.log 2>&1; echo "$k rc=$?"
  ncu -i /tmp/${k}_full.ncu-rep --page raw --csv > $O/${k}_full_raw.csv 2>/dev/null
done
cuobjdump -sass khronos_b200/csrc/libkhronos_b200.so 2>/dev/null | grep -E "UBLKCP|UTMA|SYNCS|LDGSTS" | awk '{print $2}' | sort | uniq -c > $O/sass_async_ops.txt
du -sh $O
