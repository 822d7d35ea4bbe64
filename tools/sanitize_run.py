#!/usr/bin/env python
"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck), VERDICT r01 #9:

    compute-sanitizer --tool racecheck python tools/sanitize_run.py resident

Cases: `resident` / `tiled` / `streaming` (K-SOR on 96x96 and 300x157, a few sweeps, checked against the oracle),
`deep` (the resident kernel with one exchange per sweep: 64x444 and 40x740, three and five rows per CTA),
`peer` (three slabs on one GPU driving the fused ghost-row exchange), `design` (init + one transport iteration + one
height iteration on a 24x24 mesh).  Every case checks its result, so a sanitizer run that perturbs timing still has
to produce the right bits."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import poisson_caustic_design_b200 as P  # noqa: E402
from oracle import oracle as O  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "resident"
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
port = O.OracleLib()
rng = np.random.RandomState(0)

if case in ("resident", "tiled", "streaming"):
    path = {"resident": P.SOLVER_RESIDENT, "tiled": P.SOLVER_TILED, "streaming": P.SOLVER_STREAMING}[case]
    for (h, w) in ((96, 96), (157, 300)):
        D = rng.standard_normal((h, w))
        D -= D.mean()
        phi0 = rng.standard_normal((h, w))
        s = P.Solver(w, h, 0, path)
        s.upload(D, phi0)
        info = s.run(sweeps, 0.0)
        got = s.download()
        s.close()
        assert np.array_equal(got, port.poisson_rb(D, phi0, sweeps, 0.0)[0]), (case, h, w)
        print(case, (h, w), "ok", info["path"], info["sweeps"], flush=True)
elif case == "deep":
    for (h, w) in ((444, 64), (740, 40)):
        D = rng.standard_normal((h, w))
        D -= D.mean()
        phi0 = rng.standard_normal((h, w))
        s = P.Solver(w, h, 0, P.SOLVER_RESIDENT)
        s.upload(D, phi0)
        info = s.run(sweeps, 0.0)
        got = s.download()
        assert s.resident_exchange == 2, "the deep-halo kernel did not run"
        s.close()
        assert np.array_equal(got, port.poisson_rb(D, phi0, sweeps, 0.0)[0]), (case, h, w)
        print(case, (h, w), "ok", info["path"], info["sweeps"], flush=True)
elif case == "peer":
    from poisson_caustic_design_b200 import slab
    H, W, G = 157, 300, 3
    D = rng.standard_normal((H, W))
    D -= D.mean()
    phi0 = rng.standard_normal((H, W))
    engines = []
    for g in range(G):
        row0, rows = slab.partition(H, G, g)
        e = slab.CudaSlabEngine(W, H, row0, rows, 0)
        e.upload(slab.with_ghosts(D, row0, rows, e.GH), slab.with_ghosts(phi0, row0, rows, e.GH))
        engines.append(e)
    info = slab.solve_local_peer(engines, sweeps, 0.0, 16)
    got = np.concatenate([e.download() for e in engines], axis=0)
    for e in engines:
        e.close()
    assert np.array_equal(got, port.poisson_rb(D, phi0, sweeps, 0.0)[0])
    print("peer ok", info, flush=True)
elif case == "design":
    yy, xx = np.mgrid[0:96, 0:96].astype(np.float64)
    img = 0.1 + np.exp(-((xx - 40) ** 2 + (yy - 55) ** 2) / 300.0) + (np.hypot(xx - 70, yy - 30) < 12)
    s, resized = O.prepare_image(img, 24, 0.5, 1.5, 0.1)
    cd = P.from_setup(s)
    cd.initialize_solvers(resized)
    od = port.design(s, solver_mode=0)
    od.initialize_solvers(resized)
    a, b = cd.perform_transport_iteration(), od.transport_iteration()
    assert abs(a - b) < 1e-6 * abs(b)
    cd.perform_height_map_iteration(0)
    cd.close()
    od.close()
    print("design ok", a, flush=True)
else:
    raise SystemExit("unknown case " + case)
