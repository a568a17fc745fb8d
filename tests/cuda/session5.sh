#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
rm -f gpurun_out/s5_*.log
for v in exp3 exp4 exp5; do
echo "== prof v3 $v n_masked=0" >> gpurun_out/s5_prof.log
ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 40 python tests/cuda/umma3_prof.py 8 0 >> gpurun_out/s5_prof.log 2>&1
ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 40 python tests/cuda/umma_time.py 8 0 >> gpurun_out/s5_prof.log 2>&1
done
cat gpurun_out/s5_prof.log
