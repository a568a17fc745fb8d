#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 400 python -m pytest tests/test_gated_pool_gpu.py tests/test_consumers_gpu.py tests/test_mha_gpu.py -q -m gpu > gpurun_out/s42_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/s42_pytest.log
timeout 120 python tests/cuda/shard_time.py 8
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s42_launches.csv python tests/cuda/shard_time.py 8 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/s42_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
for k, v in agg.items():
    if 'gp_' in k: print(f"{k:70s} n={len(v):3d} mean {sum(v) / len(v) / 1e3:8.1f} us  max {max(v)/1e3:8.1f}")
PY
timeout 60 python tests/cuda/umma_time.py 16 10
