#!/bin/bash
# wait-time breakdown of the fp16-split kernels (profiling build), one-CTA and CTA-pair tilings
cd "$(dirname "$0")/../.."
export ACMIL_B200_NO_REBUILD=1
export ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_gprof
for v in 0 1 2; do
  ACMIL_GEMM_PAIR=$v timeout 100 python tests/cuda/gemm_h_prof.py 2>&1 | tail -36
done
