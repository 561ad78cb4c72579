#!/bin/bash
# round 2, call 3 (1 GPU): full -m gpu suite (new defaults, mesh, virtual-rank replay), A/B of item classes / CTAs per SM
set -u
O=gpurun_out/r2c3
mkdir -p $O
echo "== gpu tests"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "rc=$?"; tail -15 $O/gpu_tests.log
run() { n=$1; shift; env "$@" timeout 200 python bench.py --no-e2e --no-cpu-baseline > $O/bench_$n.json 2> $O/bench_$n.err; python -c "
import json;d=json.load(open('$O/bench_$n.json'));print('$n',round(d['value']),'fps',d['checksum']['sum'],round(d['roofline']['launch_us'],1),'us/group')" || tail -3 $O/bench_$n.err; }
run default A=1
run mb12 KB_PRODUCT_LIB_VARIANT=mb12
run default_ctas8 KB_FUSE_CTAS_PER_SM=8
run nolist KB_FUSE_ITEM_LIST=0
run default2 A=1
echo "== full default bench with legs"
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_full.json'));print(round(d['value']),'fps; e2e',d['e2e'] and round(d['e2e']['value']));print('tick',d.get('output_tick'));print('dynamic',d.get('configs'));print('next',d.get('next_rows'))" || tail -5 $O/bench_full.err
