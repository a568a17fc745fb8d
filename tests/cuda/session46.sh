#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py -q -m gpu -k "sorted or overflow or rescue or full_size or golden" > gpurun_out/s46_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/s46_pytest.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/s46_launches.csv python tests/cuda/umma_time.py 16 10 > /dev/null 2>&1
grep "gp_" gpurun_out/s46_launches.csv | awk -F'","' '{print $5, $NF}' | tail -5
timeout 60 python tests/cuda/umma_time.py 16 10
