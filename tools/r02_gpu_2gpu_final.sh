#!/bin/bash
# two-GPU confirmation: the tests that skip below two GPUs, then the bench as the driver launches it at N=2
cd /root/repo
mkdir -p gpurun_out/r02r
( time timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02r/pytest_2gpu.log 2>&1
tail -5 gpurun_out/r02r/pytest_2gpu.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02r/bench_n2.json 2> gpurun_out/r02r/bench_n2.err
tail -c 300 gpurun_out/r02r/bench_n2.err
cut -c1-300 gpurun_out/r02r/bench_n2.json
