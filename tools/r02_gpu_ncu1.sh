#!/bin/bash
mkdir -p gpurun_out/r02l
cd /root/repo
for v in E A; do
  PCD_LIB=/root/repo/poisson_caustic_design_b200/variants/libpcd_$v.so timeout 500 ncu --set full --clock-control none --import-source on -k regex:sor_wave -s 2 -c 1 -o gpurun_out/r02l/wave_$v -f python tools/wave_time.py 8192x1024 --sweeps 64 > gpurun_out/r02l/ncu_$v.log 2>&1
  tail -3 gpurun_out/r02l/ncu_$v.log
done
ls -la gpurun_out/r02l/
