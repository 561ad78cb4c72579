#!/bin/bash
# round 2, call 4 (1 GPU): re-run of the tests that failed / are new, prefetch A/B, full default bench with legs
set -u
O=gpurun_out/r2c4
mkdir -p $O
echo "== gpu tests (subset)"
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_parity_gpu_color.py tests/test_zz_mesh.py tests/test_zzz_fuse_variants.py tests/test_instance_forwarding.py tests/test_host_adaptor.py tests/test_bench_shape_parity.py tests/test_cell_sharded_replay.py -m gpu -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "rc=$?"; tail -8 $O/gpu_tests.log
run() { n=$1; shift; env "$@" timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-legs > $O/bench_$n.json 2> $O/bench_$n.err; python -c "
import json;d=json.load(open('$O/bench_$n.json'));print('$n',round(d['value']),'fps',d['checksum']['sum'],round(d['roofline']['launch_us'],1),'us/group')" || tail -3 $O/bench_$n.err; }
run pref0 KB_FUSE_PREFETCH=0
run pref1 KB_FUSE_PREFETCH=1
run pref2 KB_FUSE_PREFETCH=2
run pref1_mb12 KB_FUSE_PREFETCH=1 KB_PRODUCT_LIB_VARIANT=mb12
echo "== full default bench with legs"
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_full.json'));print(round(d['value']),'fps; e2e',d['e2e'] and round(d['e2e']['value']), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']));print('tick',d.get('output_tick'));print('dynamic',d.get('configs'));print('next',d.get('next_rows'))" || tail -5 $O/bench_full.err
