#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r02w
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02w/launches_bench_noslab.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-slab > gpurun_out/r02w/bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02w/launches_bench_noslab.csv')))
hdr=None; acc=collections.Counter(); cnt=collections.Counter()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        try:
            name=r[hdr.index('Kernel Name')].split('(')[0][:60]; v=float(r[hdr.index('Metric Value')]); acc[name]+=v; cnt[name]+=1
        except: pass
tot=sum(acc.values())
for k,v in acc.most_common(8): print(f"{k:60s} {cnt[k]:5d} {v/1e6:10.3f} ms {100*v/tot:6.2f}%")
PY
