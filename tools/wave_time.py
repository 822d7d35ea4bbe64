#!/usr/bin/env python
"""Times the large-grid K-SOR path (wavefront kernel) on one GPU for a list of WxH grids: us per sweep, fixed sweep count.
    python tools/wave_time.py 8192x1024 8192x2048 2048x2048 [--sweeps 512]"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import poisson_caustic_design_b200 as P

sweeps = 512
shapes = []
args = sys.argv[1:]
while args:
    a = args.pop(0)
    if a == "--sweeps":
        sweeps = int(args.pop(0))
    else:
        w, h = a.split("x")
        shapes.append((int(w), int(h)))
for (W, H) in shapes:
    rng = np.random.RandomState(1)
    D = rng.standard_normal((H, W)) * 1e-3
    D -= D.mean()
    s = P.Solver(W, H, 0, P.SOLVER_TILED)
    s.upload(D, np.zeros_like(D))
    s.run(64, 0.0)
    best = None
    for _ in range(3):
        info = s.run(sweeps, 0.0)
        best = info["kernel_ms"] if best is None else min(best, info["kernel_ms"])
    s.close()
    us = best * 1e3 / sweeps
    print(f"{W}x{H}: {us:.2f} us/sweep, {24.0 * W * H / us / 1e3:.0f} GB/s algorithmic, {info['launches']} launches", flush=True)
