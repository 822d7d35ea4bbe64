#!/bin/bash
mkdir -p gpurun_out/r02g
cd /root/repo
timeout 400 python bench.py --gpus 8 --no-cpu > gpurun_out/r02g/bench_g8.json 2> gpurun_out/r02g/bench_g8.err; echo "bench g8 rc=$?"
timeout 300 python bench.py --gpus 4 --no-cpu > gpurun_out/r02g/bench_g4.json 2> gpurun_out/r02g/bench_g4.err; echo "bench g4 rc=$?"
for ce in 64 512; do
  timeout 200 python bench.py --gpus 8 --workload c5slab --steps 16 --warmup 1 --check-every $ce > gpurun_out/r02g/slab_g8_ce$ce.json 2> gpurun_out/r02g/slab_g8_ce$ce.err
done
python - <<'PY'
import json,glob
def load(f):
    for line in open(f):
        if line.startswith('{'): return json.loads(line)
for f in sorted(glob.glob('gpurun_out/r02g/*.json')):
    try:
        d=load(f); s=d.get('slab',{}); print(f, round(d['value'],2), {k:(round(v['us_per_sweep'],2), v['mode']) for k,v in s.items()})
    except Exception as e: print(f, 'ERR', e)
PY
