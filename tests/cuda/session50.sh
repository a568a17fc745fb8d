#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 200 python -m pytest tests/test_gated_pool_gpu.py -q -m gpu -k "backward or training_step" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s50_launches.csv python tests/cuda/train_step_time.py 50000 fused:eager > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/s50_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
for k, v in agg.items():
    if 'gp_bwd' in k or 'transpose' in k or 'tm_gemm' in k: print(f"{k:60s} n={len(v):3d} mean {sum(v) / len(v) / 1e3:8.1f} us")
PY
