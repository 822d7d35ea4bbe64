#!/bin/bash
mkdir -p gpurun_out/r02k
cd /root/repo
for v in F N512; do
  echo "variant $v"
  PCD_LIB=/root/repo/poisson_caustic_design_b200/variants/libpcd_$v.so timeout 200 python tools/wave_time.py 8192x1024 8192x8192 2048x2048 2>&1 | tee -a gpurun_out/r02k/variants3_$v.txt
done
