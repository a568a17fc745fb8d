#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gp_reduce_kernel" -s 6 -c 1 -f -o gpurun_out/s36_reduce1 python tests/cuda/umma_time.py 1 10 > gpurun_out/s36_ncu.log 2>&1
echo done
