#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s21_*
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py -q -m gpu -k "backward or training_step" > gpurun_out/s21_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/s21_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s21_launches.csv python tests/cuda/train_step_time.py 50000 kernel:eager > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/s21_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
names = [r[ki] for r in rows[1:]]; t = [float(r[vi].replace(',', '')) / 1e3 for r in rows[1:]]
# the last training step: find the last occurrence of the AdamW kernel and walk back to the previous one
idx = [i for i, n in enumerate(names) if 'multi_tensor' in n or 'adam' in n.lower()]
end = idx[-1]; prev = max(i for i in idx if i < end - 20)
tot = 0
for i in range(prev + 1, end + 1):
    tot += t[i]
    print(f"{t[i]:9.1f} us  {names[i][:110]}")
print(f"one step: {tot:.1f} us in {end - prev} launches")
PY
