#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s38_*
export ACMIL_B200_NO_REBUILD=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/s38_pytest_all.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/s38_pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/s38_bench.json 2> gpurun_out/s38_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/s38_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s38_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'warmup', 'gpu_launches', 'launch')}); print(d['roofline']); print(d['e2e']['value'], d['e2e']['fp16_features']['value']); print(d.get('fp16_features', {}).get('value'), d.get('fp16_features', {}).get('roofline', {}).get('frac')); print(d.get('gpu_eager_baseline', {}).get('value'), d.get('train_step'))
PY
