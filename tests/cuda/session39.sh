#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export ACMIL_B200_NO_REBUILD=1
N=${1:-8}
rm -f gpurun_out/s39_*_$N.*
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s39_bench_$N.json 2> gpurun_out/s39_bench_$N.err
echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --workload transmil --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s39_tm_$N.json 2> gpurun_out/s39_tm_$N.err
echo "transmil rc=$?"; tail -2 gpurun_out/s39_tm_$N.err
python - $N <<'PY'
import json, sys
for nm in ("bench", "tm"):
    try:
        d = json.loads([l for l in open(f'gpurun_out/s39_{nm}_{sys.argv[1]}.json').read().strip().splitlines() if l.startswith('{')][-1])
        print(nm, {k: d.get(k) for k in ('value', 'ms_per_step', 'n_gpus', 'scaling', 'gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d.get('parity'))
    except Exception as e:
        print(nm, "failed", e)
PY
