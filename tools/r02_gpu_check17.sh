#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r02z
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sor_resident -s 1 -c 1 -o gpurun_out/r02z/resident_r02 -f python tools/gpu_probe.py resident 1024 1024 300 > gpurun_out/r02z/ncu_res.log 2>&1; tail -2 gpurun_out/r02z/ncu_res.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dct_gemm_tma -s 4 -c 1 -o gpurun_out/r02z/dct_gemm_r02 -f python bench.py --backend dct --no-cpu --no-slab --steps 2 > gpurun_out/r02z/ncu_dct.log 2>&1; tail -2 gpurun_out/r02z/ncu_dct.log
ls -la gpurun_out/r02z
