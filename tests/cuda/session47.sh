#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
N=${1:-1}
rm -f gpurun_out/s47_bench_$N.*
if [ "$N" = "1" ]; then
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/s47_bench_$N.json 2> gpurun_out/s47_bench_$N.err
else
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s47_bench_$N.json 2> gpurun_out/s47_bench_$N.err
fi
echo "bench rc=$?"
python - $N <<'PY'
import json, sys
d = json.loads([l for l in open(f'gpurun_out/s47_bench_{sys.argv[1]}.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'gpu_launches')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], (d.get('parity') or {}).get('parity_ok'))
print(d.get('fp16_features'))
PY
