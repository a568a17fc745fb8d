#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s7_*
timeout 200 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu -k "exchange or sharded" > gpurun_out/s7_pytest_x.log 2>&1
echo "rc=$?" >> gpurun_out/s7_pytest_x.log; tail -15 gpurun_out/s7_pytest_x.log
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/s7_pytest_all.log 2>&1
echo "rc=$?" >> gpurun_out/s7_pytest_all.log; tail -5 gpurun_out/s7_pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s7_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s7_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'warmup', 'gpu_launches', 'launch')}); print(d['roofline']); print(d['e2e']); print(d.get('gpu_eager_baseline')); print(d.get('cpu_baseline'))
PY
