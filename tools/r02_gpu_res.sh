#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r02z
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sor_resident -s 1 -c 1 -o gpurun_out/r02z/resident_r02b -f python tools/gpu_probe.py resident 1024 1024 300 > gpurun_out/r02z/ncu_res_b.log 2>&1; tail -1 gpurun_out/r02z/ncu_res_b.log
timeout 600 python bench.py > gpurun_out/r02z/bench.json 2> gpurun_out/r02z/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-slab > gpurun_out/r02z/bench_k20.json 2>/dev/null
python - <<'PY'
import json
for f in ('bench','bench_k20'):
    d=json.load(open(f'gpurun_out/r02z/{f}.json')); print(f, round(d['value'],2), d['steps'], round(d['e2e']['value'],2), round(d['design']['iters_per_s'],2), d['roofline']['us_per_sweep'], d['roofline']['frac'], d['roofline'].get('on_chip',{}).get('frac_of_exchange_floor'), d.get('cpu_baseline',{}).get('value'))
PY
