#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s37_*
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py tests/test_consumers_gpu.py -q -m gpu -k "front_projection or golden or clam or ibmil or attmil or abmil" > gpurun_out/s37_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/s37_pytest.log
timeout 300 python tests/cuda/dims_time.py > gpurun_out/s37_dims.log 2>&1; cat gpurun_out/s37_dims.log
