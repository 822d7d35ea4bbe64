#!/bin/bash
# A/B timing of resident-kernel build variants + the bit-exactness test of each
cd /root/repo
V=poisson_caustic_design_b200/variants
for v in "$@"; do
  echo "== variant $v"
  PCD_LIB=/root/repo/$V/libpcd_$v.so timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "deep_halo or resident_1024" 2>&1 | tail -3 | cut -c1-300
  PCD_LIB=/root/repo/$V/libpcd_$v.so timeout 200 python tools/res_time.py 1024x1024 1000x1000 400x400 1024x512 2>&1 | tail -4
done
