#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi.py tests/test_gpu_dct.py -m gpu -q 2>&1 | tail -6 | cut -c1-300
