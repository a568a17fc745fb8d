#!/bin/bash
cd "$(dirname "$0")/../.."
export ACMIL_B200_NO_REBUILD=1
timeout 200 python -m pytest tests/test_transmil_gpu.py -q -m gpu -x -k "other_tilings or chunked" 2>&1 | tail -4
