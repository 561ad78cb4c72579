#!/bin/bash
# round 2, 4-GPU call: routed ingest and fusion-CTA budget experiments
set -u
O=gpurun_out/r2m4
mkdir -p $O
run() { np=$1; n=$2; port=$3; shift 3; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np --no-e2e "$@" > $O/bench_$n.json 2> $O/bench_$n.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$n.json'));print('$n',round(d['value']),'fps',d['checksum']['sum']);print('   shards',d.get('shards'));print('   exchange gbps',d.get('exchange') and d['exchange']['gbps_per_rank'], 'gather ms', d['exchange']['gather_ms_per_rank'])" || tail -5 $O/bench_$n.err; }
run 4 routed 29811
run 4 routed_ctas8 29812 --fuse-ctas-per-sm 8
run 4 routed_ctas6 29813 --fuse-ctas-per-sm 6
run 4 striped 29814 --ingest striped
run 4 routed_c20 29815 --cell-blocks 20
