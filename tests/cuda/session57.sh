#!/bin/bash
# A/B of the chunked softmax on both paths
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_transmil_gpu.py -q -m gpu -x 2>&1 | tail -3
for v in 0 1; do
  echo "ACMIL_CHUNKED_SOFTMAX=$v"
  ACMIL_CHUNKED_SOFTMAX=$v timeout 120 python bench.py --workload transmil --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | cut -c1-200
  ACMIL_CHUNKED_SOFTMAX=$v timeout 120 python bench.py --workload vit --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | cut -c1-200
done
