#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for r in 1 8; do timeout 100 python tests/cuda/shard_time.py $r; done 2>&1 | tee gpurun_out/s11_shard.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/s11_launches_shard8.csv python tests/cuda/shard_time.py 8 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/s11_launches_shard8.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} mean {sum(v) / len(v) / 1e3:8.1f} us")
PY
