#!/bin/bash
# builds kernel variants next to the default library: tests/cuda/build_variants.sh name "-DFLAG=1 ..." [name flags ...]
set -e
cd "$(dirname "$0")/../.."
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ACMIL_B200_LIB_DIR=$PWD/acmil_b200/lib_$name ACMIL_NVCC_EXTRA="$flags" python -m acmil_b200.build >/dev/null
  echo "built acmil_b200/lib_$name ($flags)"
done
