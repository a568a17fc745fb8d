#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s16_*
export ACMIL_B200_NO_REBUILD=1
timeout 200 python -m pytest tests/test_gated_pool_gpu.py tests/test_consumers_gpu.py -x -q -m gpu > gpurun_out/s16_pytest_gp.log 2>&1
echo "pytest gp rc=$?" >> gpurun_out/s16_pytest_gp.log
tail -8 gpurun_out/s16_pytest_gp.log
for bags in 8 16; do
  echo "== dbuf bags=$bags" >> gpurun_out/s16_time.log
  timeout 60 python tests/cuda/umma_time.py $bags 0 10 >> gpurun_out/s16_time.log 2>&1
done
for nm in 0 10; do
  echo "== prof dbuf n_masked=$nm" >> gpurun_out/s16_prof.log
  ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 60 python tests/cuda/umma_prof.py 8 $nm >> gpurun_out/s16_prof.log 2>&1
done
cat gpurun_out/s16_time.log; grep -v "^ *e[0-9]\|^wait\|^epi\|^soft\|^pool\|^flush\|^total" gpurun_out/s16_prof.log | head -70
