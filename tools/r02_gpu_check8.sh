#!/bin/bash
mkdir -p gpurun_out/r02k
cd /root/repo
./tools/tma_stride_probe > gpurun_out/r02k/tma_probe.txt 2>&1; head -5 gpurun_out/r02k/tma_probe.txt
for v in E C F; do
  echo "variant $v"
  PCD_LIB=/root/repo/poisson_caustic_design_b200/variants/libpcd_$v.so timeout 200 python tools/wave_time.py 8192x1024 8192x8192 2048x2048 2>&1 | tee -a gpurun_out/r02k/variants2_$v.txt
done
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_slab.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
