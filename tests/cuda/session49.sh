#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s49_*
export ACMIL_B200_NO_REBUILD=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s49_launches_bench.csv python bench.py --steps 3 --warmup 3 --launch eager --no-train-step --no-fp16 --no-gpu-eager --no-cpu-baseline --e2e-steps 0 > gpurun_out/s49_b.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/s49_launches_bench.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:80]].append(float(r[vi].replace(',', '')))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:80s} n={len(v):3d} mean {sum(v) / len(v) / 1e3:8.1f} us  total {sum(v)/1e3:9.1f} us")
PY
