#!/bin/bash
# end-of-round records: default bench, the bench on the reference's own README case (C1), ncu of the shipped resident kernel, launch list
mkdir -p gpurun_out/r02z
cd /root/repo
timeout 600 python bench.py > gpurun_out/r02z/bench_n1.json 2> gpurun_out/r02z/bench_n1.err
tail -c 300 gpurun_out/r02z/bench_n1.err; cut -c1-160 gpurun_out/r02z/bench_n1.json
timeout 400 python bench.py --workload c1 --no-slab > gpurun_out/r02z/bench_c1.json 2> gpurun_out/r02z/bench_c1.err
tail -c 300 gpurun_out/r02z/bench_c1.err; cut -c1-160 gpurun_out/r02z/bench_c1.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sor_resident_deep -s 1 -c 1 -o gpurun_out/r02z/deep_final -f python tools/res_time.py 1024x1024 --sweeps 300 > gpurun_out/r02z/ncu.log 2>&1
tail -2 gpurun_out/r02z/ncu.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02z/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-slab > gpurun_out/r02z/bench_under_ncu.log 2>&1
grep -c sor_resident_deep gpurun_out/r02z/launches_bench.csv
