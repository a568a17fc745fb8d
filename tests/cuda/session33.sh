#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s33_*
export ACMIL_B200_NO_REBUILD=1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gp_reduce_kernel" -s 6 -c 1 -f -o gpurun_out/s33_reduce1 python tests/cuda/umma_time.py 1 10 > gpurun_out/s33_ncu.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/s33_launches.csv python tests/cuda/umma_time.py 1 10 > /dev/null 2>&1
grep "gp_" gpurun_out/s33_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
