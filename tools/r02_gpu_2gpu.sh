#!/bin/bash
# 2-GPU box: the multi-GPU tests the 1-GPU runs skip, bench --gpus 2, slab timing with different check intervals
mkdir -p gpurun_out/r02f
cd /root/repo
nvidia-smi topo -m > gpurun_out/r02f/topo.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -8 | cut -c1-300
timeout 600 python bench.py --gpus 2 --no-cpu > gpurun_out/r02f/bench_g2.json 2> gpurun_out/r02f/bench_g2.err; echo "bench g2 rc=$?"; tail -3 gpurun_out/r02f/bench_g2.err
for ce in 64 256 1024; do
  timeout 300 python bench.py --gpus 2 --workload c5slab --steps 16 --warmup 1 --check-every $ce > gpurun_out/r02f/slab_g2_ce$ce.json 2> gpurun_out/r02f/slab_g2_ce$ce.err
done
python - <<'PY'
import json,glob
d=json.load(open('gpurun_out/r02f/bench_g2.json')); print('bench g2', d['value'], json.dumps(d.get('slab'))[:1200])
for f in sorted(glob.glob('gpurun_out/r02f/slab_g2_ce*.json')):
    try:
        d=json.load(open(f)); print(f, d['value'], d['slab']['c5']['us_per_sweep'], d['slab']['c5']['mode'])
    except Exception as e: print(f, 'ERR', e)
PY
