#!/bin/bash
# round-2 first GPU call: parity suite, smoke, bench N=1, sanitizers (bounded)
set -x
mkdir -p gpurun_out/r02a
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest.log
tail -5 gpurun_out/r02a/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02a/smoke.log
timeout 600 python bench.py > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02a/bench_k20.json 2> gpurun_out/r02a/bench_k20.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a/bench_ref.json 2> gpurun_out/r02a/bench_ref.err
for c in design resident; do
  ( time timeout 90 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py $c 4 ) > gpurun_out/r02a/san_memcheck_${c}.txt 2>&1
  echo "memcheck $c rc=$?" >> gpurun_out/r02a/san_summary.txt
done
cat gpurun_out/r02a/san_summary.txt
tail -c 1500 gpurun_out/r02a/bench.json
