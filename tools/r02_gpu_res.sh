#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "deep_halo or production_sizes or resident_1024" 2>&1 | tail -15 | cut -c1-400
timeout 300 python tools/res_time.py 1024x1024 1000x1000 400x400 1024x512 1024x896 2>&1 | tail -8
