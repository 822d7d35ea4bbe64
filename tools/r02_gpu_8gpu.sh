#!/bin/bash
mkdir -p gpurun_out/r02r
cd /root/repo
timeout 200 python bench.py --gpus 8 --no-cpu 2>gpurun_out/r02r/bench_g8.err | grep '^{' > gpurun_out/r02r/bench_g8.json; echo "g8 rc=$?"
timeout 150 python bench.py --gpus 4 --no-cpu 2>gpurun_out/r02r/bench_g4.err | grep '^{' > gpurun_out/r02r/bench_g4.json; echo "g4 rc=$?"
timeout 150 python bench.py --gpus 2 --no-cpu 2>gpurun_out/r02r/bench_g2.err | grep '^{' > gpurun_out/r02r/bench_g2.json; echo "g2 rc=$?"
PCD_WAVE_TRACE=gpurun_out/r02r/trace8 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29551 tools/slab_run.py --W 8192 --H 8192 --sweeps 128 --check_every 64 2>&1 | grep us_per_sweep
python tools/wave_trace.py gpurun_out/r02r/trace8_row*.bin --strips 17 --json gpurun_out/r02r/trace8_summary.json > /dev/null
python - <<'PY'
import json
for n in (8,4,2):
    try:
        d=json.load(open(f'gpurun_out/r02r/bench_g{n}.json')); print(n, round(d['value'],2), round(d['e2e']['value'],2), {k:(round(v['us_per_sweep'],2), v['mode'], v['bit_identical_to_1gpu']) for k,v in d['slab'].items()})
    except Exception as e: print(n, 'ERR', e)
PY
