#!/bin/bash
cd /root/repo
V=poisson_caustic_design_b200/variants
for v in "$@"; do
  echo "== variant $v"
  PCD_LIB=/root/repo/$V/libpcd_$v.so timeout 400 python -m pytest tests/test_gpu_solver.py tests/test_gpu_soak.py -m gpu -q -x -k "deep_halo_kernel_bit or soak or production" 2>&1 | tail -2 | cut -c1-300
  PCD_LIB=/root/repo/$V/libpcd_$v.so timeout 300 python tools/res_time.py 1024x1024 1000x1000 400x400 1024x512 1280x720 1024x1332 300x157 2>&1 | tail -7
done
