#!/bin/bash
mkdir -p gpurun_out/r02e
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_slab.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -8 | cut -c1-300
for wl in c2k c4k; do
  timeout 300 python bench.py --workload $wl --steps 2 --no-cpu --no-slab > gpurun_out/r02e/bench_$wl.json 2> gpurun_out/r02e/bench_$wl.err
  PCD_WAVE_LAUNCH_PER_PASS=1 timeout 300 python bench.py --workload $wl --steps 2 --no-cpu --no-slab > gpurun_out/r02e/bench_${wl}_perpass.json 2> gpurun_out/r02e/bench_${wl}_perpass.err
done
timeout 300 python bench.py --no-cpu --steps 1 > gpurun_out/r02e/bench.json 2> gpurun_out/r02e/bench.err
PCD_WAVE_LAUNCH_PER_PASS=1 timeout 300 python bench.py --no-cpu --steps 1 > gpurun_out/r02e/bench_perpass.json 2> gpurun_out/r02e/bench_perpass.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02e/bench*.json')):
    try:
        d=json.load(open(f)); print(f, 'it/s', round(d['value'],3), 'us/sweep', d['roofline']['us_per_sweep'], 'launches', d['gpu_launches'], d.get('slab',{}).get('c5',{}).get('us_per_sweep'))
    except Exception as e: print(f, 'ERR', e)
PY
