#!/bin/bash
mkdir -p gpurun_out/r02h
cd /root/repo
./tools/tma_stride_probe > gpurun_out/r02h/tma_probe.txt 2>&1; cat gpurun_out/r02h/tma_probe.txt
timeout 300 python tools/wave_time.py 8192x1024 8192x2048 8192x4096 2048x2048 1024x1024 > gpurun_out/r02h/wave_time.txt 2>&1; cat gpurun_out/r02h/wave_time.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | cut -c1-300
