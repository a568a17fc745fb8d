#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s15_*
export ACMIL_B200_NO_REBUILD=1
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/s15_pytest_all.log 2>&1
echo "rc=$?" >> gpurun_out/s15_pytest_all.log; tail -15 gpurun_out/s15_pytest_all.log
