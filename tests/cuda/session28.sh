#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s28_*
export ACMIL_B200_NO_REBUILD=1
timeout 600 python -m pytest tests/test_transmil_gpu.py -x -q -m gpu > gpurun_out/s28_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/s28_pytest.log
