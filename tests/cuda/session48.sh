#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s48_*
export ACMIL_B200_NO_REBUILD=1
echo "== 16 x 50000 train" >> gpurun_out/s48_prof.log
ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 60 python tests/cuda/umma_prof.py 16 10 >> gpurun_out/s48_prof.log 2>&1
echo "== 128 x 6250 train" >> gpurun_out/s48_prof.log
UMMA_PROF_ROWS=6250 ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 60 python tests/cuda/umma_prof.py 128 10 >> gpurun_out/s48_prof.log 2>&1
echo "== 128 x 6250 eval" >> gpurun_out/s48_prof.log
UMMA_PROF_ROWS=6250 ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 60 python tests/cuda/umma_prof.py 128 0 >> gpurun_out/s48_prof.log 2>&1
grep "==\|cta rank 0\|MMA idle\|MMA total\|EPI \|SM clock" gpurun_out/s48_prof.log | grep -v "mean            0"
