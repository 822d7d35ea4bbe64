#!/usr/bin/env python
"""Times the on-chip (resident) K-SOR kernel on one GPU for a list of WxH grids, both variants: us per sweep at a fixed
sweep count.  `deep` = one neighbour exchange per sweep, `phase` = one per colour phase (PCD_RES_NO_DEEP=1).
    python tools/res_time.py 1024x1024 1000x1000 400x400 [--sweeps 4000]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poisson_caustic_design_b200 as P

sweeps = 4000
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
    s = P.Solver(W, H, 0, P.SOLVER_RESIDENT)
    out = {}
    fields = {}
    for name in ("deep", "phase"):
        if name == "phase":
            os.environ["PCD_RES_NO_DEEP"] = "1"
        else:
            os.environ.pop("PCD_RES_NO_DEEP", None)
        s.upload(D, np.zeros_like(D))
        s.run(64, 0.0)
        best = None
        for _ in range(3):
            s.upload(D, np.zeros_like(D))
            info = s.run(sweeps, 0.0)
            best = info["kernel_ms"] if best is None else min(best, info["kernel_ms"])
        out[name] = (best * 1e3 / sweeps, s.resident_exchange)
        fields[name] = s.download()
    s.close()
    same = np.array_equal(fields["deep"], fields["phase"])
    print(f"{W}x{H}: deep {out['deep'][0]:.3f} us/sweep (kernel {out['deep'][1]}), phase {out['phase'][0]:.3f} us/sweep "
          f"(kernel {out['phase'][1]}), same bits: {same}", flush=True)
