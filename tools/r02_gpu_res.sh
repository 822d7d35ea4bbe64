#!/bin/bash
cd /root/repo
timeout 120 python tools/gpu_probe.py resident 1024 1024 3000 2>&1 | tail -3
timeout 120 python tools/gpu_probe.py resident 1000 1000 3000 2>&1 | tail -1
timeout 120 python tools/gpu_probe.py resident 400 400 3000 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_soak.py tests/test_gpu_pipeline.py -m gpu -q -x 2>&1 | tail -3
