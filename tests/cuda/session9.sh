#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s9_bench_peer_$N.json 2> gpurun_out/s9_bench_peer_$N.err
echo "peer rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/s9_bench_peer_$N.err | tail -5
python - <<PY
import json
d = json.loads(open(f'gpurun_out/s9_bench_peer_$N.json').read().strip().splitlines()[-1])
print({q: d.get(q) for q in ('value', 'ms_per_step', 'launch', 'parity_ok')}, d['roofline']['kernel_ms'], d['e2e']['value'] if d.get('e2e') else None)
PY
