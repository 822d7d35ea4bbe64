#!/bin/bash
mkdir -p gpurun_out/r02y
cd /root/repo
for n in 8 4 2; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/r02y/bench_g$n.err | grep '^{' > gpurun_out/r02y/bench_g$n.json; echo "g$n rc=$?"
done
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.load(open(f'gpurun_out/r02y/bench_g{n}.json')); print(n, round(d['value'],2), d['steps'], round(d['e2e']['value'],2), {k:(round(v['us_per_sweep'],2), v['bit_identical_to_1gpu']) for k,v in d['slab'].items()}, d['clocks']['reasons'])
    except Exception as e: print(n, 'ERR', e)
PY
