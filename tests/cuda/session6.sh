#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
rm -f gpurun_out/s6_*.log
timeout 90 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu > gpurun_out/s6_pytest_gp.log 2>&1
echo "pytest gp rc=$?" >> gpurun_out/s6_pytest_gp.log
tail -4 gpurun_out/s6_pytest_gp.log
echo "== v3 bags=8" >> gpurun_out/s6_prof.log
timeout 40 python tests/cuda/umma_time.py 8 0 10 >> gpurun_out/s6_prof.log 2>&1
for NM in 0 10; do v=prof
echo "== prof v3 $v n_masked=0" >> gpurun_out/s6_prof.log
ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 40 python tests/cuda/umma3_prof.py 8 $NM >> gpurun_out/s6_prof.log 2>&1
ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$v timeout 40 python tests/cuda/umma_time.py 8 0 >> gpurun_out/s6_prof.log 2>&1
done
cat gpurun_out/s6_prof.log
