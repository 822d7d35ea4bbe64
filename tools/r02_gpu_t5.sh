#!/bin/bash
cd /root/repo
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for n in 0 29 37 45 52 59; do
  echo "chunks=$n: $(PCD_WAVE_CHUNKS=$n timeout 100 python tools/wave_time.py 2048x2048 --sweeps 512 2>&1 | tail -1)"
done
for n in 0 16 24 28 32; do
  echo "chunks=$n: $(PCD_WAVE_CHUNKS=$n timeout 100 python tools/wave_time.py 4096x4096 --sweeps 256 2>&1 | tail -1)"
done
