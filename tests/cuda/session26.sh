#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s26_*
export ACMIL_B200_NO_REBUILD=1
timeout 300 python -m pytest tests/test_gated_pool_gpu.py -q -m gpu -k "fp16 or bags" > gpurun_out/s26_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/s26_pytest.log
timeout 400 python bench.py --workload stream --no-cpu-baseline > gpurun_out/s26_stream_1.json 2> gpurun_out/s26_stream_1.err
echo "stream rc=$?"; tail -3 gpurun_out/s26_stream_1.err; cut -c1-1200 gpurun_out/s26_stream_1.json
