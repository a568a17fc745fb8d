#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s8_*
N=${1:-2}
nvidia-smi topo -m > gpurun_out/s8_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s8_bench_peer_$N.json 2> gpurun_out/s8_bench_peer_$N.err
echo "peer rc=$?"; tail -5 gpurun_out/s8_bench_peer_$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --exchange nccl > gpurun_out/s8_bench_nccl_$N.json 2> gpurun_out/s8_bench_nccl_$N.err
echo "nccl rc=$?"; tail -5 gpurun_out/s8_bench_nccl_$N.err
python - <<PY
import json
for k in ("peer", "nccl"):
    try:
        d = json.loads(open(f'gpurun_out/s8_bench_{k}_$N.json').read().strip().splitlines()[-1])
        print(k, {q: d.get(q) for q in ('value', 'ms_per_step', 'launch', 'parity_ok', 'exchange')}, d['roofline']['kernel_ms'], d['e2e']['value'] if d.get('e2e') else None, d.get('parity'))
    except Exception as e:
        print(k, 'no line', e)
PY
