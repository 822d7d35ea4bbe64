#!/bin/bash
cd /root/repo
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for s in "1024 592" "512 1024" "1024 1036" "300 157"; do set -- $s; timeout 60 python tools/gpu_probe.py resident $1 $2 2000 2>&1 | tail -1; done
