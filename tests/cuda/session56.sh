#!/bin/bash
# chunked softmax across the score / value products, 64-wide tiles for small problems: parity + timings + launch lists
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 400 python -m pytest tests/test_transmil_gpu.py tests/test_vit_gpu.py tests/test_resnet_gpu.py tests/test_stream_gpu.py -q -m gpu -x 2>&1 | tail -8
timeout 120 python bench.py --workload vit --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | cut -c1-240
timeout 120 python bench.py --workload transmil --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 | cut -c1-240
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_r2_transmil_h.csv python bench.py --workload transmil --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1; echo "ncu list rc=$?"
