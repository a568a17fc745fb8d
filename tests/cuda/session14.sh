#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gp_finish_kernel" -s 4 -c 1 -f -o gpurun_out/s14_finish python tests/cuda/shard_time.py 8 > gpurun_out/s14_ncu.log 2>&1
ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_nofence timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/s14_launches_nofence.csv python tests/cuda/shard_time.py 8 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/s14_launches_nofence.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:60]].append(float(r[vi].replace(',', '')))
for k, v in agg.items():
    if 'gp_' in k: print(f"nofence {k:60s} n={len(v):3d} mean {sum(v) / len(v) / 1e3:8.1f} us")
PY
