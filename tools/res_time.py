import numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import poisson_caustic_design_b200 as P
rng = np.random.RandomState(0)
W = H = 1024
D = rng.standard_normal((H, W)); D -= D.mean()
s = P.Solver(W, H, 0, P.SOLVER_RESIDENT)
s.upload(D, np.zeros_like(D))
s.run(500, 0.0)
s.upload(D, np.zeros_like(D))
info = s.run(4000, 0.0)
print("resident 1024^2:", info["kernel_ms"] * 1e3 / info["sweeps"], "us/sweep", "nowait" if os.environ.get("PCD_RES_NOWAIT") else "")
