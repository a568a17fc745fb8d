#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/s44_pytest_all.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/s44_pytest_all.log
