#!/bin/bash
mkdir -p gpurun_out/r02m
cd /root/repo
timeout 500 ncu --set full --clock-control none --import-source on -k regex:sor_resident_deep -s 1 -c 1 -o gpurun_out/r02m/deep -f python tools/res_time.py 1024x1024 --sweeps 300 > gpurun_out/r02m/ncu.log 2>&1
tail -3 gpurun_out/r02m/ncu.log
ls -la gpurun_out/r02m/
