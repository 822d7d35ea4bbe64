"""Scratch probe for a GPU box: quick correctness + timing of the K-SOR kernel families.
    python tools/gpu_probe.py [streaming|resident] [W] [H] [sweeps]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poisson_caustic_design_b200 as P

path = sys.argv[1] if len(sys.argv) > 1 else "resident"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
H = int(sys.argv[3]) if len(sys.argv) > 3 else W
sweeps = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
pid = {"streaming": P.SOLVER_STREAMING, "resident": P.SOLVER_RESIDENT, "auto": P.SOLVER_AUTO, "tiled": P.SOLVER_TILED}[path]
rng = np.random.RandomState(0)
D = rng.standard_normal((H, W)); D -= D.mean()
s = P.Solver(W, H, 0, pid)
print("path", s.path, flush=True)
s.upload(D, np.zeros_like(D))
for rep in range(3):
    s.upload(None, np.zeros_like(D))
    info = s.run(sweeps, 0.0)
    us = info["device_ms"] * 1e3 / max(info["sweeps"], 1)
    print(f"{path} {W}x{H}: {info['sweeps']} sweeps in {info['device_ms']:.3f} ms -> {us:.3f} us/sweep, "
          f"{24.0 * W * H / (us * 1e-6) / 1e9:.1f} GB/s algorithmic, launches {info['launches']}", flush=True)
if W * H <= 512 * 512:
    from oracle import oracle as O
    want = O.OracleLib().poisson_rb(D, np.zeros_like(D), sweeps, 0.0)[0]
    got = s.download()
    print("bit-exact vs RB oracle:", np.array_equal(got, want), "max diff", np.abs(got - want).max())
