#!/bin/bash
set -x
mkdir -p gpurun_out/r02b
cd /root/repo
timeout 300 python -X faulthandler -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02b/smoke.log
tail -30 gpurun_out/r02b/smoke.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02b/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b/pytest.log
tail -15 gpurun_out/r02b/pytest.log
timeout 300 python bench.py --backend dct --no-slab --no-cpu > gpurun_out/r02b/bench_dct.json 2> gpurun_out/r02b/bench_dct.err
tail -c 600 gpurun_out/r02b/bench_dct.json
