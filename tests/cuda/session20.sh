#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s20_*
export ACMIL_B200_NO_REBUILD=1
timeout 600 python -m pytest tests/test_gated_pool_gpu.py tests/test_consumers_gpu.py tests/test_mha_gpu.py -q -m gpu > gpurun_out/s20_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/s20_pytest.log
timeout 300 python tests/cuda/train_step_time.py > gpurun_out/s20_train.log 2>&1; grep -v Warning gpurun_out/s20_train.log | tail -8
timeout 60 python tests/cuda/umma_time.py 16 0 10
