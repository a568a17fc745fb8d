#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
N=${1:-2}
rm -f gpurun_out/s29_tm_$N.*
if [ "$N" = "1" ]; then
timeout 400 python bench.py --workload transmil --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s29_tm_$N.json 2> gpurun_out/s29_tm_$N.err
else
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --workload transmil --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s29_tm_$N.json 2> gpurun_out/s29_tm_$N.err
fi
echo "bench rc=$?"; tail -4 gpurun_out/s29_tm_$N.err
python - $N <<'PY'
import json, sys
d = json.loads([l for l in open(f'gpurun_out/s29_tm_{sys.argv[1]}.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'scaling', 'gpu_launches', 'parity')}, d['roofline']['frac'], d['e2e']['value'], d['config']['parallelism'])
PY
