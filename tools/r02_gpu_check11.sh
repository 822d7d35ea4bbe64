#!/bin/bash
mkdir -p gpurun_out/r02p
cd /root/repo
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02p/bench.json 2> gpurun_out/r02p/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-slab > gpurun_out/r02p/bench_k20.json 2> gpurun_out/r02p/bench_k20.err
timeout 300 python bench.py --backend dct --no-cpu --no-slab > gpurun_out/r02p/bench_dct.json 2> gpurun_out/r02p/bench_dct.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:sor_wave -s 2 -c 1 -o gpurun_out/r02p/wave_tma_8192 -f python tools/wave_time.py 8192x8192 --sweeps 64 > gpurun_out/r02p/ncu_tma.log 2>&1; tail -2 gpurun_out/r02p/ncu_tma.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02p/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02p/bench_under_ncu.log 2>&1
python - <<'PY'
import json
for f in ['bench','bench_k20','bench_dct']:
    try:
        d=json.load(open(f'gpurun_out/r02p/{f}.json')); print(f, round(d['value'],2), d['steps'], round(d['e2e']['value'],2), d['design']['iters_per_s'], d['roofline']['us_per_sweep'], {k:round(v['us_per_sweep'],2) for k,v in d.get('slab',{}).items()}, d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
