#!/bin/bash
# round 2, 4-GPU validation of the layout selection (bisect table) and the default multi-GPU command line
set -u
O=gpurun_out/r2m4b
mkdir -p $O
timeout 400 python -m pytest tests/test_cell_sharded_replay.py -m gpu -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "rc=$?"; tail -3 $O/gpu_tests.log
run() { np=$1; n=$2; port=$3; shift 3; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" > $O/bench_$n.json 2> $O/bench_$n.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$n.json'));print('$n',round(d['value']),'fps',d['checksum']['sum'],'|',d['config']['parallelism'][:150]);print('   proxy',d['config'].get('layout_max_frames_per_rank'));print('   shards',d.get('shards'));print('   exchange gbps',d.get('exchange') and d['exchange']['gbps_per_rank']);print('   e2e',d.get('e2e') and round(d['e2e']['value']))" || tail -5 $O/bench_$n.err; }
run 4 n4_auto 29911
run 2 n2_auto 29912 --no-e2e
run 4 n4_tiling 29913 --layout tiling --no-e2e
