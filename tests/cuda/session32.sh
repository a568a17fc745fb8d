#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s32_*
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu > gpurun_out/s32_pytest_gp.log 2>&1
echo "pytest gp rc=$?"; tail -3 gpurun_out/s32_pytest_gp.log
for v in warpmma default warpmma default; do
  echo "== $v bags=16" >> gpurun_out/s32_time.log
  if [ "$v" = "default" ]; then
    timeout 60 python tests/cuda/umma_time.py 16 0 10 >> gpurun_out/s32_time.log 2>&1
  else
    ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 60 python tests/cuda/umma_time.py 16 0 10 >> gpurun_out/s32_time.log 2>&1
  fi
done
cat gpurun_out/s32_time.log
echo "== prof one-thread issuer n_masked=0" >> gpurun_out/s32_prof.log
ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 60 python tests/cuda/umma_prof.py 16 0 >> gpurun_out/s32_prof.log 2>&1
grep -v "^ *e[0-9]\|^wait\|^epi\|^soft\|^pool\|^flush\|^total" gpurun_out/s32_prof.log | head -24
