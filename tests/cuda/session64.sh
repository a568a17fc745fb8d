#!/bin/bash
# PPEG with a shared-memory halo tile, register-blocked depth-wise conv: parity + TransMIL timing + launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_transmil_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 120 python bench.py --workload transmil --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | cut -c1-200
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 230 --csv --log-file gpurun_out/launches_r2_transmil_h2.csv python bench.py --workload transmil --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1; echo "ncu list rc=$?"
