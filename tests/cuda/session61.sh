#!/bin/bash
# eight converter warps in the fp16-split kernels: parity, timings of the three tilings, wait-time breakdown
cd "$(dirname "$0")/../.."
export ACMIL_B200_NO_REBUILD=1
timeout 120 python -m pytest tests/test_transmil_gpu.py -q -m gpu -x -k "fp16_split" 2>&1 | tail -3
for v in 0 1 2; do
  echo "ACMIL_GEMM_PAIR=$v"
  ACMIL_GEMM_PAIR=$v timeout 120 python tests/cuda/gemm_split_time.py 2>&1 | tail -10 | cut -c1-150
done
ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_gprof ACMIL_GEMM_PAIR=1 timeout 100 python tests/cuda/gemm_h_prof.py 2>&1 | tail -36 | head -24
