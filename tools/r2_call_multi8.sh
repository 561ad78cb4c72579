#!/bin/bash
# round 2, 8-GPU call: cell-sharded replay at N = 8 (and N = 4 with the time left)
set -u
N=8
O=gpurun_out/r2m8
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
run() { np=$1; n=$2; port=$3; shift 3; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" > $O/bench_$n.json 2> $O/bench_$n.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$n.json'));print('$n',round(d['value']),'fps',d['checksum']['sum'],'|',d['config']['parallelism'][:70]);print('   shards',d.get('shards'));print('   exchange gbps',d.get('exchange') and d['exchange']['gbps_per_rank']);print('   e2e',d.get('e2e') and round(d['e2e']['value']))" || tail -5 $O/bench_$n.err; }
run 8 n8_c16 29711
run 8 n8_c12 29712 --cell-blocks 12 --no-e2e
run 8 n8_c20 29713 --cell-blocks 20 --no-e2e
run 8 n8_rank0 29714 --ingest rank0 --no-e2e
run 4 n4_c16 29715
run 4 n4_c12 29716 --cell-blocks 12 --no-e2e
