#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py -q -m gpu -k "backward" > gpurun_out/s45_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/s45_pytest.log; grep -n "AssertionError: (" gpurun_out/s45_pytest.log | head
