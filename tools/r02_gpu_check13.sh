#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r02v
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02v/bench_k20.json 2> gpurun_out/r02v/bench_k20.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02v/bench_k20.json')); print(round(d['value'],2), d['steps'], d['warmup'], round(d['e2e']['value'],2), d['design'], {k:round(v['us_per_sweep'],2) for k,v in d['slab'].items()}, d['cpu_baseline']['value'], d['clocks'])
PY
