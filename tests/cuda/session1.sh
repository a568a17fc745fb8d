#!/bin/bash
# GPU session 1 of round 2: correctness of the reworked masking path + timings / cycle counters of the row pass
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt
timeout 600 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu > gpurun_out/s1_pytest_gp.log 2>&1
echo "pytest gp rc=$?" >> gpurun_out/s1_pytest_gp.log
for v in "" _oldsleep; do
  for bags in 8 16; do
    echo "== lib$v bags=$bags" >> gpurun_out/s1_time.log
    ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib$v timeout 120 python tests/cuda/umma_time.py $bags 0 10 >> gpurun_out/s1_time.log 2>&1
  done
done
for nm in 0 10; do
  echo "== prof n_masked=$nm" >> gpurun_out/s1_prof.log
  ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_prof timeout 120 python tests/cuda/umma_prof.py 8 $nm >> gpurun_out/s1_prof.log 2>&1
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s1_pytest_all.log 2>&1
echo "pytest all rc=$?" >> gpurun_out/s1_pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -3 gpurun_out/s1_pytest_gp.log; cat gpurun_out/s1_time.log; tail -2 gpurun_out/s1_pytest_all.log
