#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gated_pool_gpu.py -x -q -m gpu -k "exchange or sharded or golden" > gpurun_out/s10_pytest_x.log 2>&1
echo "rc=$?" >> gpurun_out/s10_pytest_x.log; tail -4 gpurun_out/s10_pytest_x.log
bash tests/cuda/session9.sh $1
