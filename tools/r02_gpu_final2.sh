#!/bin/bash
# whole GPU suite, sanitizers over both resident kernels, small-grid timings, default bench
mkdir -p gpurun_out/r02p
cd /root/repo
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02p/pytest.log 2>&1
tail -6 gpurun_out/r02p/pytest.log | cut -c1-300
timeout 300 python tools/res_time.py 300x157 157x300 96x96 64x64 200x100 400x400 1024x1024 2>&1 | tail -7
for tool in memcheck synccheck racecheck; do
  for c in deep resident; do
    ( time timeout 200 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_run.py $c 4 ) > gpurun_out/r02p/${tool}_${c}.txt 2>&1
    echo "$tool $c rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02p/${tool}_${c}.txt | tail -1)" | tee -a gpurun_out/r02p/summary.txt
  done
  ( time PCD_RES_NO_DEEP=1 timeout 200 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_run.py resident 4 ) > gpurun_out/r02p/${tool}_phase.txt 2>&1
  echo "$tool phase rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02p/${tool}_phase.txt | tail -1)" | tee -a gpurun_out/r02p/summary.txt
done
timeout 600 python bench.py > gpurun_out/r02p/bench_n1.json 2> gpurun_out/r02p/bench_n1.err
tail -c 400 gpurun_out/r02p/bench_n1.err
cut -c1-200 gpurun_out/r02p/bench_n1.json
