#!/bin/bash
# round 2, final 8-GPU validation of the default configuration (the command line the driver uses)
set -u
O=gpurun_out/r2f8
mkdir -p $O
for np in 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 2991$np bench.py --gpus $np --steps 10 --warmup 3 > $O/bench_n$np.json 2> $O/bench_n$np.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_n$np.json'));print('N=$np',round(d['value']),'fps',d['checksum']['sum'],'|',d['config']['parallelism'][:100]);print('   shards',d.get('shards'));print('   exchange gbps',d.get('exchange') and d['exchange']['gbps_per_rank']);print('   e2e',d.get('e2e') and round(d['e2e']['value']), d['clocks'])" || tail -5 $O/bench_n$np.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29921 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $O/bench_ref8.json 2> $O/bench_ref8.err; echo "ref rc=$?"; head -c 600 $O/bench_ref8.json
