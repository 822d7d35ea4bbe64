#!/bin/bash
mkdir -p gpurun_out/r02n
cd /root/repo
timeout 120 python tools/sanitize_run.py tiled 6 2>&1 | tail -3
timeout 120 python tools/sanitize_run.py peer 6 2>&1 | tail -2
echo "--- TMA"; timeout 150 python tools/wave_time.py 8192x1024 8192x8192 2048x2048 4096x4096 2>&1 | tee gpurun_out/r02n/wave_tma.txt
echo "--- no TMA"; PCD_WAVE_NO_TMA=1 timeout 150 python tools/wave_time.py 8192x1024 8192x8192 2048x2048 4096x4096 2>&1 | tee gpurun_out/r02n/wave_notma.txt
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_slab.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
