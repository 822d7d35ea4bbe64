"""GPU soak: 10^6 sweeps at 1024 x 1024 through the two kernels whose CTAs talk to each other without
a grid-wide barrier -- the on-chip kernel (flag-in-data halo messages through L2 / DSMEM, warps drifting up to 15
phases apart, stop announced ahead) and the persistent wavefront kernel (neighbour-only sequence words between
passes) -- against the per-colour streaming kernels, which have neither.  Every 10^5 sweeps the three fields are
compared bit for bit and the right-hand side is replaced, so the fields keep moving.  ~40 s on a B200."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_soak_million_sweeps_1024(pcd):
    n, seg, nseg = 1024, 100_000, 10
    rng = np.random.RandomState(2026)
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    solvers = {name: pcd.Solver(n, n, 0, path) for name, path in
               (("streaming", pcd.SOLVER_STREAMING), ("resident", pcd.SOLVER_RESIDENT), ("tiled", pcd.SOLVER_TILED))}
    phi0 = rng.standard_normal((n, n))
    for s in solvers.values():
        s.upload(None, phi0)
    total = 0
    for k in range(nseg):
        D = (np.cos(np.pi * (xx + 0.5) / n * (k + 1)) * np.cos(np.pi * (yy + 0.5) / n * (2 * k + 1)) + 0.05 * rng.standard_normal((n, n))) * 1e-3
        D -= D.mean()
        infos = {}
        for name, s in solvers.items():
            s.upload(D, None)                 # new right-hand side, the field continues
            infos[name] = s.run(seg, 0.0)     # tol 0: never stops early
            assert infos[name]["sweeps"] == seg and infos[name]["path"] == name
        total += seg
        ref = solvers["streaming"].download()
        for name in ("resident", "tiled"):
            got = solvers[name].download()
            assert np.array_equal(got, ref), (name, total, np.abs(got - ref).max())
            assert infos[name]["last_max_update"] == infos["streaming"]["last_max_update"], (name, total)
    for s in solvers.values():
        s.close()
    assert total == 1_000_000
