#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s17_*
export ACMIL_B200_NO_REBUILD=1
for nm in 0 10; do
  echo "== prof dbuf n_masked=$nm" >> gpurun_out/s17_prof.log
  ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 60 python tests/cuda/umma_prof.py 8 $nm >> gpurun_out/s17_prof.log 2>&1
done
grep "SM clock\|==" gpurun_out/s17_prof.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv
