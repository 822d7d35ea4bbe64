#!/bin/bash
# compute-sanitizer over small workloads of every kernel family; every case also checks its result
mkdir -p gpurun_out/r02s
cd /root/repo
for tool in memcheck synccheck racecheck; do
  for c in design streaming tiled peer resident; do
    ( time timeout 150 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_run.py $c 4 ) > gpurun_out/r02s/${tool}_${c}.txt 2>&1
    echo "$tool $c rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02s/${tool}_${c}.txt | tail -1)" | tee -a gpurun_out/r02s/summary.txt
  done
done
