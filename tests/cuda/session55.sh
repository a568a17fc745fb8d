#!/bin/bash
# fp16-split GEMM mode, second pass: v^T products through the transposed store, ncu evidence of the new kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 400 python -m pytest tests/test_transmil_gpu.py tests/test_vit_gpu.py tests/test_resnet_gpu.py tests/test_stream_gpu.py tests/test_gated_pool_gpu.py -q -m gpu -x 2>&1 | tail -8
timeout 120 python bench.py --workload vit --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | cut -c1-240
timeout 120 python bench.py --workload transmil --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | cut -c1-240
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tm_gemm_h_kernel -s 3 -c 1 -f -o gpurun_out/gemm_h_fc2 python tests/cuda/gemm_h_one.py > /dev/null 2>&1; echo "ncu rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 220 --csv --log-file gpurun_out/launches_r2_vit_h.csv python bench.py --workload vit --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1; echo "ncu list rc=$?"
