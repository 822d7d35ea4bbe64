#!/bin/bash
# round-2 confirmation run: whole GPU suite, the default bench, ncu of the deep-halo resident kernel, launch list
mkdir -p gpurun_out/r02n
cd /root/repo
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02n/pytest.log 2>&1
tail -6 gpurun_out/r02n/pytest.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02n/bench_n1.json 2> gpurun_out/r02n/bench_n1.err
tail -c 600 gpurun_out/r02n/bench_n1.err
cut -c1-700 gpurun_out/r02n/bench_n1.json
timeout 300 python tools/res_time.py 1024x1024 1000x1000 400x400 1024x512 300x157 2>&1 | tail -5
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sor_resident_deep -s 1 -c 1 -o gpurun_out/r02n/deep2 -f python tools/res_time.py 1024x1024 --sweeps 300 > gpurun_out/r02n/ncu.log 2>&1
tail -2 gpurun_out/r02n/ncu.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02n/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-slab > gpurun_out/r02n/bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02n/launches_bench.csv | cut -c1-300
ls -la gpurun_out/r02n/
