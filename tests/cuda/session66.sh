#!/bin/bash
# verification of the tree with the CTA-pair fp16-split weight products: full GPU suite, smoke, bench lines
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/s66_*
export ACMIL_B200_NO_REBUILD=1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s66_pytest_all.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/s66_pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for wl in vit resnet transmil; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/s66_$wl.json 2> gpurun_out/s66_$wl.err
  echo "$wl rc=$?"
done
timeout 500 python bench.py > gpurun_out/s66_bench.json 2> gpurun_out/s66_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for nm in ("vit", "resnet", "transmil", "bench"):
    try:
        d = json.loads([l for l in open(f'gpurun_out/s66_{nm}.json').read().strip().splitlines() if l.startswith('{')][-1])
        print(nm, d.get('metric'), round(d.get('value'), 1), d.get('unit'), 'steps', d.get('steps'), 'warmup', d.get('warmup'), 'frac', round((d.get('roofline') or {}).get('frac', 0), 4), 'e2e', round(d['e2e']['value'], 1), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(nm, "failed", e)
PY
