#!/bin/bash
mkdir -p gpurun_out/r02m
cd /root/repo
PCD_WAVE_TRACE=gpurun_out/r02m/t1 timeout 200 python tools/wave_time.py 8192x1024 --sweeps 64 2>&1 | tail -1
PCD_WAVE_TRACE=gpurun_out/r02m/t8 timeout 200 python tools/wave_time.py 8192x8192 --sweeps 64 2>&1 | tail -1
python tools/wave_trace.py gpurun_out/r02m/t1_solver8192x1024.bin gpurun_out/r02m/t8_solver8192x8192.bin --strips 17 --json gpurun_out/r02m/trace_1gpu.json
timeout 600 python -m pytest tests/test_gpu_soak.py -m gpu -q 2>&1 | tail -3
