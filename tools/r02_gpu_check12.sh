#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r02q
for b in 0 3 5 8; do echo "bias $b"; PCD_WAVE_ROW_BIAS=$b timeout 120 python tools/wave_time.py 8192x1024 8192x8192 8192x2048 2>&1 | tee -a gpurun_out/r02q/bias_$b.txt; done
PCD_WAVE_ROW_BIAS=5 timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "8192 or tiled" 2>&1 | tail -2
