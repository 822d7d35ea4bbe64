#!/bin/bash
mkdir -p gpurun_out/r02u
cd /root/repo
PCD_WAVE_TMA_MIN_ROWS=1 PCD_WAVE_TRACE=gpurun_out/r02u/tma timeout 100 python tools/wave_time.py 8192x1024 --sweeps 64 | tail -1
PCD_WAVE_NO_TMA=1 PCD_WAVE_TRACE=gpurun_out/r02u/notma timeout 100 python tools/wave_time.py 8192x1024 --sweeps 64 | tail -1
python tools/wave_trace.py gpurun_out/r02u/tma_solver8192x1024.bin gpurun_out/r02u/notma_solver8192x1024.bin --strips 17 | cut -c1-900
