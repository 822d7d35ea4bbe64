#!/bin/bash
# 4 GPUs: 8192 x 4096 problem -> 1024-row slabs (the geometry of 8192^2 on 8 GPUs), ranks 1 and 2 have two neighbours
mkdir -p gpurun_out/r02i
cd /root/repo
export MASTER_PORT=29540
PCD_WAVE_TRACE=gpurun_out/r02i/trace4 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29541 tools/slab_run.py --W 8192 --H 4096 --sweeps 128 --check_every 64 > gpurun_out/r02i/slab4_trace.log 2>&1
tail -2 gpurun_out/r02i/slab4_trace.log
python tools/wave_trace.py gpurun_out/r02i/trace4_row*.bin --json gpurun_out/r02i/trace4_summary.json
for ce in 64 512; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29542 tools/slab_run.py --W 8192 --H 4096 --sweeps 1024 --check_every $ce 2>&1 | grep us_per_sweep
done
timeout 300 python bench.py --gpus 4 --no-cpu 2>/dev/null | grep '^{' > gpurun_out/r02i/bench_g4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02i/bench_g4.json')); print('g4', d['value'], {k:(round(v['us_per_sweep'],2)) for k,v in d['slab'].items()})
PY
