#!/bin/bash
# CTA-pair fp16-split kernel: parity (wrapped in a short timeout: a barrier mistake would hang), then timings
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 120 python -m pytest tests/test_transmil_gpu.py -q -m gpu -x -k "fp16_split" 2>&1 | tail -6
echo "pair tests rc=$?"
timeout 120 python tests/cuda/gemm_split_time.py 2>&1 | tail -11
ACMIL_GEMM_PAIR=0 timeout 120 python tests/cuda/gemm_split_time.py 2>&1 | tail -11 | head -4
timeout 300 python -m pytest tests/test_transmil_gpu.py tests/test_vit_gpu.py tests/test_resnet_gpu.py tests/test_stream_gpu.py tests/test_gated_pool_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 120 python bench.py --workload vit --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | cut -c1-200
timeout 120 python bench.py --workload transmil --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | cut -c1-200
timeout 120 python bench.py --workload resnet --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | cut -c1-200
