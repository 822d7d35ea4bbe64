#!/bin/bash
mkdir -p gpurun_out/r02t
cd /root/repo
run() { PCD_WAVE_TOP_CREDIT=$1 PCD_WAVE_TAIL_ROWS=$2 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port $3 tools/slab_run.py --W 8192 --H 4096 --sweeps 1024 --check_every 64 2>&1 | grep us_per_sweep | sed "s/^/credit=$1 tail=$2 /"; }
run 0 0 29601
run 8 0 29602
run 14 0 29603
run 20 0 29604
run 14 32 29605
run 0 32 29606
PCD_WAVE_TOP_CREDIT=14 PCD_WAVE_TRACE=gpurun_out/r02t/t4c14 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29607 tools/slab_run.py --W 8192 --H 4096 --sweeps 128 --check_every 64 2>&1 | grep us_per_sweep
python tools/wave_trace.py gpurun_out/r02t/t4c14_row1024.bin --strips 17 | cut -c1-700
