#!/bin/bash
cd /root/repo
V=poisson_caustic_design_b200/variants
mkdir -p gpurun_out/r02o
for v in abc d; do
  echo "== variant $v"
  PCD_LIB=/root/repo/$V/libpcd_$v.so timeout 300 python -m pytest tests/test_gpu_solver.py -m gpu -q -x -k "deep_halo or resident_1024" 2>&1 | tail -2 | cut -c1-300
  PCD_LIB=/root/repo/$V/libpcd_$v.so timeout 300 python tools/res_time.py 1024x1024 1000x1000 400x400 1024x512 512x512 1024x700 1024x880 800x600 600x800 2>&1 | tail -9
done
PCD_LIB=/root/repo/$V/libpcd_d.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:sor_resident_deep -s 1 -c 1 -o gpurun_out/r02o/deep_d -f python tools/res_time.py 1024x1024 --sweeps 300 > gpurun_out/r02o/ncu.log 2>&1
tail -2 gpurun_out/r02o/ncu.log
