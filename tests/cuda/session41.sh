#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gp_reduce_kernel" -s 20 -c 1 -f -o gpurun_out/s41_reduce8 python tests/cuda/shard_time.py 8 > gpurun_out/s41_ncu.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gp_finish_kernel" -s 10 -c 1 -f -o gpurun_out/s41_finish8 python tests/cuda/shard_time.py 8 >> gpurun_out/s41_ncu.log 2>&1
echo done
