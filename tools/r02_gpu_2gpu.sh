#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q -k "soak or production or two_gpu" 2>&1 | tail -6 | cut -c1-300
