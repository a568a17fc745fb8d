#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
timeout 80 python bench.py --workload stream --no-cpu-baseline > gpurun_out/s68_stream.json 2> gpurun_out/s68_stream.err; echo "stream rc=$?"
tail -c 600 gpurun_out/s68_stream.json | cut -c1-400
