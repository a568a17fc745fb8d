#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
N=${1:-8}
rm -f gpurun_out/s51_vit_$N.*
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --workload vit --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s51_vit_$N.json 2> gpurun_out/s51_vit_$N.err
echo "vit rc=$?"; tail -2 gpurun_out/s51_vit_$N.err
python - $N <<'PY'
import json, sys
d = json.loads([l for l in open(f'gpurun_out/s51_vit_{sys.argv[1]}.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k: d.get(k) for k in ('metric', 'value', 'ms_per_step', 'n_gpus', 'scaling', 'slides_per_sec_100k_patches')}, d['roofline']['frac'], d['e2e']['value'])
PY
