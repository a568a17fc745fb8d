#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s25_*
export ACMIL_B200_NO_REBUILD=1
timeout 600 python -m pytest tests/test_gated_pool_gpu.py tests/test_stream_gpu.py -q -m gpu > gpurun_out/s25_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/s25_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-train-step > gpurun_out/s25_bench.json 2> gpurun_out/s25_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s25_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s25_bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}); print(d['roofline']['frac']); print(d['e2e']['value'], d['e2e']['fp16_features']); print(d.get('fp16_features'))
PY
