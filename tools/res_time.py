"""Scratch: time the resident K-SOR kernel.  python tools/res_time.py [W] [H] [sweeps]"""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poisson_caustic_design_b200 as P
W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = int(sys.argv[2]) if len(sys.argv) > 2 else W
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
rng = np.random.RandomState(0)
D = rng.standard_normal((H, W)); D -= D.mean()
s = P.Solver(W, H, 0, P.SOLVER_RESIDENT)
s.upload(D, np.zeros_like(D))
s.run(500, 0.0)
s.upload(D, np.zeros_like(D))
info = s.run(n, 0.0)
us = info["kernel_ms"] * 1e3 / info["sweeps"]
print(f"resident {W}x{H}: {us:.3f} us/sweep, {24.0 * W * H / us / 1e3:.0f} GB/s algorithmic")
