#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s23_*
export ACMIL_B200_NO_REBUILD=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/s23_pytest_all.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/s23_pytest_all.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/s23_bench.json 2> gpurun_out/s23_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s23_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s23_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'warmup', 'gpu_launches', 'launch')}); print(d['roofline']); print(d['e2e']); print(d.get('gpu_eager_baseline')); print(d.get('train_step'))
PY
