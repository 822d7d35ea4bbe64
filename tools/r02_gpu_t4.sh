#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r02q
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02q/pytest.log 2>&1
tail -4 gpurun_out/r02q/pytest.log | cut -c1-300
timeout 300 python tools/res_time.py 1024x1024 400x400 1024x512 1280x720 300x157 2>&1 | tail -5
timeout 300 python tools/wave_time.py 2048x2048 4096x4096 8192x1024 8192x8192 --sweeps 256 2>&1 | tail -4
