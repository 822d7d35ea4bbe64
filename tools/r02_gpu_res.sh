#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "c5_stages or c4_first or height_tol" 2>&1 | tail -12 | cut -c1-300
