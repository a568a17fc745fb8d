#!/bin/bash
# fp16-split GEMM mode: parity of the kernel, the modules on it, and timings against the 3xTF32 engine
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 240 python -m pytest tests/test_transmil_gpu.py -q -m gpu -x -k "gemm" 2>&1 | tail -15
timeout 120 python tests/cuda/gemm_split_time.py 2>&1 | tail -12
timeout 400 python -m pytest tests/test_transmil_gpu.py tests/test_vit_gpu.py tests/test_resnet_gpu.py tests/test_stream_gpu.py tests/test_extract_gpu.py -q -m gpu 2>&1 | tail -15
timeout 120 python bench.py --workload vit --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | cut -c1-600
timeout 120 python bench.py --workload transmil --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | cut -c1-600
timeout 120 python bench.py --workload resnet --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | cut -c1-600
