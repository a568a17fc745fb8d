#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
for v in prof exp1; do
  echo "== $v n_masked=0" >> gpurun_out/s2_prof.log
  ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 120 python tests/cuda/umma_prof.py 8 0 >> gpurun_out/s2_prof.log 2>&1
  ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 120 python tests/cuda/umma_time.py 8 0 >> gpurun_out/s2_prof.log 2>&1
done
timeout 600 python -m pytest tests/test_transmil_gpu.py -x -q -m gpu > gpurun_out/s2_pytest_tm.log 2>&1
tail -3 gpurun_out/s2_pytest_tm.log
grep -A12 "epilogue warps" gpurun_out/s2_prof.log; grep "row pass" gpurun_out/s2_prof.log
