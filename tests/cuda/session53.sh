#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py -q -m gpu 2>&1 | tail -2
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s53_launches.csv python tests/cuda/umma_time.py 1 10 > /dev/null 2>&1
grep "gp_" gpurun_out/s53_launches.csv | awk -F'","' '{print $5, $NF}' | tail -4
timeout 200 python tests/cuda/train_step_time.py 50000 fused:graph 2>&1 | grep "ms per"
timeout 60 python tests/cuda/umma_time.py 16 10
timeout 120 python tests/cuda/shard_time.py 8
