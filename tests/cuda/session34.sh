#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s34_*
export ACMIL_B200_NO_REBUILD=1
timeout 400 python -m pytest tests/test_gated_pool_gpu.py tests/test_consumers_gpu.py tests/test_mha_gpu.py tests/test_stream_gpu.py -q -m gpu > gpurun_out/s34_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/s34_pytest.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s34_launches.csv python tests/cuda/umma_time.py 1 10 > /dev/null 2>&1
grep "gp_" gpurun_out/s34_launches.csv | awk -F'","' '{print $5, $NF}' | tail -4
timeout 60 python tests/cuda/umma_time.py 16 10
timeout 200 python tests/cuda/train_step_time.py 50000 fused:graph,fused:eager 2>&1 | grep "ms per"
