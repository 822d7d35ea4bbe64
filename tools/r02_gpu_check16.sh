#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r02x
timeout 600 python -m pytest tests/test_gpu_dct.py -m gpu -q 2>&1 | tail -3
timeout 200 python bench.py --backend dct --no-cpu --no-slab > gpurun_out/r02x/dct_tma.json 2>/dev/null
PCD_DCT_NO_TMA=1 timeout 200 python bench.py --backend dct --no-cpu --no-slab > gpurun_out/r02x/dct_notma.json 2>/dev/null
python - <<'PY'
import json
for f in ('dct_tma','dct_notma'):
    d=json.load(open(f'gpurun_out/r02x/{f}.json')); print(f, round(d['value'],1), 'it/s', d['roofline']['kernel_ms_per_launch'], 'ms per GEMM', d['ms_by_step'][:4])
PY
