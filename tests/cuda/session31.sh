#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s31_*
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu > gpurun_out/s31_pytest_gp.log 2>&1
echo "pytest gp rc=$?"; tail -3 gpurun_out/s31_pytest_gp.log
for v in noepi default noepi default; do
  echo "== $v bags=16" >> gpurun_out/s31_time.log
  if [ "$v" = "default" ]; then
    timeout 60 python tests/cuda/umma_time.py 16 0 10 >> gpurun_out/s31_time.log 2>&1
  else
    ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 60 python tests/cuda/umma_time.py 16 0 10 >> gpurun_out/s31_time.log 2>&1
  fi
done
cat gpurun_out/s31_time.log
for nm in 0 10; do
  echo "== prof cvt-epi1 n_masked=$nm" >> gpurun_out/s31_prof.log
  ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 60 python tests/cuda/umma_prof.py 16 $nm >> gpurun_out/s31_prof.log 2>&1
done
grep -v "^ *e[0-9]\|^wait\|^epi\|^soft\|^pool\|^flush\|^total\|cta rank 1" gpurun_out/s31_prof.log | head -60
