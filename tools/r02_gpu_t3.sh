#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_solver.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
timeout 300 python tools/res_time.py 1280x720 1332x1024 1100x400 1280x960 2>&1 | tail -4
timeout 200 python tools/wave_time.py 1280x720 1332x1024 --sweeps 256 2>&1 | tail -3
