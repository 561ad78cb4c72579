#!/bin/bash
# round 2, call 7 (1 GPU): validation of the shared-memory pose staging in selectBlocksKernel: parity + bench + launch metrics
set -u
O=gpurun_out/r2c7
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_bench_shape_parity.py tests/test_golden.py tests/test_parity_gpu_color.py -m gpu -q -p no:cacheprovider > $O/gpu_tests.log 2>&1; echo "rc=$?"; tail -3 $O/gpu_tests.log
timeout 200 python bench.py --no-e2e --no-cpu-baseline --no-legs > $O/bench.json 2> $O/bench.err; python -c "
import json;d=json.load(open('$O/bench.json'));print(round(d['value']),'fps',d['checksum']['sum'],round(d['roofline']['launch_us'],1),'us/group')" || tail -3 $O/bench.err
K='regex:fuseKernel|selectBlocks|itemCull|itemCompact|tileMax|tilePyramid'
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "$K" -s 1200 -c 120 --csv --log-file $O/launches_metrics.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-legs > $O/ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
from collections import defaultdict
rows=list(csv.reader(l for l in open('gpurun_out/r2c7/launches_metrics.csv') if l.startswith('"')))
h=rows[0]; t=defaultdict(list)
for r in rows[1:]:
    if r[h.index("Metric Name")].startswith("gpu__time"):
        v=float(r[h.index("Metric Value")].replace(",","")); u=r[h.index("Metric Unit")]
        t[r[h.index("Kernel Name")][:40]].append(v*{"ns":1e-3,"us":1,"ms":1e3}.get(u,1))
for k,v in t.items(): print(k, len(v), round(sum(v)/len(v),1),"us")
PY
