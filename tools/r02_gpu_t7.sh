#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r02q
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02q/pytest.log 2>&1
tail -4 gpurun_out/r02q/pytest.log | cut -c1-300
timeout 300 python tools/res_time.py 1024x1024 400x400 1024x512 1280x720 300x157 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r02q/bench_n1.json 2> gpurun_out/r02q/bench_n1.err
tail -c 300 gpurun_out/r02q/bench_n1.err; cut -c1-160 gpurun_out/r02q/bench_n1.json
timeout 400 python bench.py --workload c1 --no-slab --no-cpu > gpurun_out/r02q/bench_c1.json 2> gpurun_out/r02q/bench_c1.err
cut -c1-160 gpurun_out/r02q/bench_c1.json
