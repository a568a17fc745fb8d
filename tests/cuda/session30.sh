#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s30_*
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu > gpurun_out/s30_pytest_gp.log 2>&1
echo "pytest gp rc=$?"; tail -3 gpurun_out/s30_pytest_gp.log
for v in poly0 poly2 default poly0 default; do
  echo "== $v bags=16" >> gpurun_out/s30_time.log
  if [ "$v" = "default" ]; then
    timeout 60 python tests/cuda/umma_time.py 16 0 10 >> gpurun_out/s30_time.log 2>&1
  else
    ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 60 python tests/cuda/umma_time.py 16 0 10 >> gpurun_out/s30_time.log 2>&1
  fi
done
cat gpurun_out/s30_time.log
