#!/bin/bash
# round 2, multi-GPU call: bash tools/r2_call_multi.sh N  (under gpurun --gpus N)
set -u
N=${1:-2}
O=gpurun_out/r2m$N
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
run() { n=$1; port=$2; shift 2; timeout 420 $TR $port bench.py --gpus $N "$@" > $O/bench_$n.json 2> $O/bench_$n.err; echo "rc=$?"; python -c "
import json;d=json.load(open('$O/bench_$n.json'));print('$n',round(d['value']),'fps',d['checksum']['sum'],'|',d['config']['parallelism'][:60]);print('   shards',d.get('shards'));print('   exchange',d.get('exchange'));print('   e2e',d.get('e2e') and round(d['e2e']['value']))" || tail -5 $O/bench_$n.err; }
if [ "$N" = "2" ]; then
  echo "== NCCL tests (skipped on 1-GPU boxes)"
  timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -q -p no:cacheprovider > $O/nccl_tests.log 2>&1; echo "rc=$?"; tail -4 $O/nccl_tests.log
fi
run cells_ce 29611 --gather ce
run cells_bulk 29612 --gather bulk --no-e2e
run cells_sm 29613 --gather sm --no-e2e
run cells_rank0 29614 --ingest rank0 --no-e2e
run hash_nccl 29615 --shard hash --no-e2e
run cells_c8 29616 --cell-blocks 8 --no-e2e
run cells_c24 29617 --cell-blocks 24 --no-e2e
