#!/bin/bash
mkdir -p gpurun_out/r02c
cd /root/repo
timeout 120 python -X faulthandler -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c/smoke_plain.log 2>&1; echo "smoke plain rc=$?"; tail -3 gpurun_out/r02c/smoke_plain.log | cut -c1-200
MALLOC_CHECK_=3 timeout 120 python -X faulthandler -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c/smoke_mc.log 2>&1; echo "smoke mc rc=$?"; tail -3 gpurun_out/r02c/smoke_mc.log | cut -c1-200
timeout 120 python -X faulthandler tools/smoke_bisect.py ABCD > gpurun_out/r02c/bisect_plain.log 2>&1; echo "ABCD plain rc=$?"; tail -3 gpurun_out/r02c/bisect_plain.log | cut -c1-200
ASAN=$(gcc -print-file-name=libasan.so)
if [ -f "$ASAN" ]; then
  LD_PRELOAD=$ASAN ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:halt_on_error=1 timeout 300 python -X faulthandler -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c/smoke_asan.log 2>&1; echo "smoke asan rc=$?"; grep -n "ERROR\|#[0-9] " gpurun_out/r02c/smoke_asan.log | head -40 | cut -c1-220
fi
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -q 2>&1 | tail -5
