#!/bin/bash
mkdir -p gpurun_out/r02d
cd /root/repo
timeout 120 python -X faulthandler -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02d/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02d/smoke.log | cut -c1-200
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02d/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02d/pytest.log | cut -c1-300
timeout 600 python bench.py --no-cpu > gpurun_out/r02d/bench.json 2> gpurun_out/r02d/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02d/bench.json
for wl in c2k; do timeout 300 python bench.py --workload $wl --steps 3 --no-cpu --no-slab > gpurun_out/r02d/bench_$wl.json 2> gpurun_out/r02d/bench_$wl.err; tail -c 400 gpurun_out/r02d/bench_$wl.json; done
