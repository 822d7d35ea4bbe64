#!/bin/bash
mkdir -p gpurun_out/r02j
cd /root/repo
./tools/tma_stride_probe > gpurun_out/r02j/tma_probe.txt 2>&1; cat gpurun_out/r02j/tma_probe.txt
timeout 300 python tools/wave_time.py 8192x1024 8192x2048 8192x4096 8192x8192 2048x2048 4096x4096 1024x1024 > gpurun_out/r02j/wave_time.txt 2>&1; cat gpurun_out/r02j/wave_time.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | cut -c1-300
