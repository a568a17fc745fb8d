#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s19_*
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py tests/test_consumers_gpu.py tests/test_mha_gpu.py tests/test_transmil_gpu.py -q -m gpu -k "backward or training_step or clam or transmil" > gpurun_out/s19_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/s19_pytest.log
timeout 200 python tests/cuda/train_step_time.py > gpurun_out/s19_train.log 2>&1; cat gpurun_out/s19_train.log | tail -5
