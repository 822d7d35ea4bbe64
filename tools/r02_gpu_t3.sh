#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_solver.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
timeout 300 python tools/res_time.py 1024x1100 1024x1332 1024x1280 720x1280 2>&1 | tail -4
timeout 200 python tools/wave_time.py 1024x1100 1024x1332 --sweeps 256 2>&1 | tail -3
