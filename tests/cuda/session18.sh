#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s18_*
export ACMIL_B200_NO_REBUILD=1
timeout 200 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu > gpurun_out/s18_pytest_gp.log 2>&1
echo "pytest gp rc=$?"; tail -3 gpurun_out/s18_pytest_gp.log
for bags in 8 16 32; do
  echo "== dbuf bags=$bags" >> gpurun_out/s18_time.log
  timeout 60 python tests/cuda/umma_time.py $bags 0 10 >> gpurun_out/s18_time.log 2>&1
done
cat gpurun_out/s18_time.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/s18_bench16.json 2> gpurun_out/s18_bench16.err
echo "bench rc=$?"; tail -3 gpurun_out/s18_bench16.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gp_main_umma_kernel" -s 3 -c 1 -f -o gpurun_out/s18_umma_train16 python tests/cuda/umma_time.py 16 10 > gpurun_out/s18_ncu.log 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s18_bench16.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'warmup', 'gpu_launches', 'launch')}); print(d['roofline']); print(d['e2e']); print(d.get('gpu_eager_baseline')); print(d.get('cpu_baseline')); print(d.get('clocks'))
PY
