"""Debug helper: runs the sections of __graft_entry__.smoke() named on the command line (e.g. `ABCD`) in order in
ONE process, under MALLOC_CHECK_, to localise a host-heap corruption.  A resident solve, B tiled solve, C design next
to the oracle design, D streaming solve, O oracle-only solve."""
import gc
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np

import poisson_caustic_design_b200 as P
from oracle import oracle as O

port = O.OracleLib()
rng = np.random.RandomState(0)


def churn(tag):
    gc.collect()
    xs = [np.empty(n) for n in (10, 1000, 9216, 96 * 96, 57600, 100000) for _ in range(20)]
    del xs
    gc.collect()
    print(tag, "ok", flush=True)


D = rng.standard_normal((96, 96))
D -= D.mean()
z = np.zeros_like(D)
for sec in sys.argv[1]:
    if sec == "A":
        sv = P.Solver(96, 96, 0, P.SOLVER_RESIDENT); sv.upload(D, z); info = sv.run(100000, 1e-7); got = sv.download(); sv.close()
        want = port.poisson_rb(D, z, 100000, 1e-7, extra_sweeps=info["sweeps"] - info["converged_at"])[0]
        assert np.array_equal(got, want)
    elif sec == "B":
        Dw = rng.standard_normal((96, 600)); Dw -= Dw.mean()
        sv = P.Solver(600, 96, 0, P.SOLVER_TILED); sv.upload(Dw, np.zeros_like(Dw)); info = sv.run(41, 0.0); got = sv.download(); sv.close()
        assert np.array_equal(got, port.poisson_rb(Dw, np.zeros_like(Dw), 41, 0.0)[0])
    elif sec in "Cc":
        yy, xx = np.mgrid[0:96, 0:96].astype(np.float64)
        img = 0.1 + np.exp(-((xx - 40) ** 2 + (yy - 55) ** 2) / 300.0) + (np.hypot(xx - 70, yy - 30) < 12)
        s, resized = O.prepare_image(img, 24, 0.5, 1.5, 0.1)
        cd = P.from_setup(s); cd.initialize_solvers(resized)
        od = port.design(s, solver_mode=0); od.initialize_solvers(resized)
        for it in range(2):
            a, b = cd.perform_transport_iteration(), od.transport_iteration()
        disp = np.abs(od.get("target_x") - od.get("source_x")).max()
        assert np.abs(cd.get("target_x") - od.get("target_x")).max() < 1e-6 * disp
        if sec == "C":
            cd.perform_height_map_iteration(0); od.height_iteration(0)
            zz, zr = cd.get("source_z"), od.get("source_z")
        cd.close(); od.close()
    elif sec == "D":
        sv = P.Solver(96, 96, 0, P.SOLVER_STREAMING); sv.upload(D, z); info = sv.run(60, 0.0); got = sv.download(); sv.close()
        assert np.array_equal(got, port.poisson_rb(D, z, 60, 0.0)[0])
    elif sec == "O":
        port.poisson_rb(D, z, 60, 0.0)
    churn(sec)
print("all ok", flush=True)
