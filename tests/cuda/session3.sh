#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 90 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu > gpurun_out/s3_pytest_gp.log 2>&1
echo "pytest gp rc=$?" >> gpurun_out/s3_pytest_gp.log
tail -15 gpurun_out/s3_pytest_gp.log
for bags in 8 16; do
  echo "== v3 bags=$bags" >> gpurun_out/s3_time.log
  timeout 40 python tests/cuda/umma_time.py $bags 0 10 >> gpurun_out/s3_time.log 2>&1
done
echo "== v2 bags=8" >> gpurun_out/s3_time.log
ACMIL_GP_UMMA_VARIANT=2 timeout 40 python tests/cuda/umma_time.py 8 0 10 >> gpurun_out/s3_time.log 2>&1
for nm in 0 10; do
  echo "== prof v3 n_masked=$nm" >> gpurun_out/s3_prof.log
  ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 40 python tests/cuda/umma3_prof.py 8 $nm >> gpurun_out/s3_prof.log 2>&1
done
cat gpurun_out/s3_time.log; cat gpurun_out/s3_prof.log
