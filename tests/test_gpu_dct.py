"""GPU: the opt-in direct Poisson backend (PCD_SOLVER_DCT) through the C ABI.

Parity contract of an OPT-IN backend (it returns the converged discrete field, the reference a field truncated at
max|delta| < tol, so bit parity is not on offer):
* against an independent direct solve on the CPU (scipy DCT-II): grad(phi) rel L-inf <= 1e-10, residual of the
  discrete equation <= 1e-9 of max|D|;
* against K-SOR run to tol 1e-9: grad(phi) rel L-inf <= 1e-7;
* whole designs (C1, C2, C3): transport-iteration count equal to the reference's, step sizes within 1e-4
  absolute, final vertices within 1e-3 of the max displacement, heights within 5e-4 of their range (the distance
  between the converged field and the reference's own stopping point at max|delta| < 1e-8 grows with the grid
  width: 7e-5 of range at W = 400, 3.2e-4 at W = 1024 -- tests/test_oracle_golden.py::test_height_truncation_evidence);
* NaN holes fall back to the masked sweeps (bit-exact against the oracle)."""
import os

import numpy as np
import pytest

from conftest import GOLD

pytestmark = pytest.mark.gpu


def grad(a):
    p = np.pad(a, 1, mode="edge")
    return (p[1:-1, 2:] - p[1:-1, :-2]) / 2, (p[2:, 1:-1] - p[:-2, 1:-1]) / 2


def direct_cpu(D):
    import scipy.fft as sf
    H, W = D.shape
    lam = (2 - 2 * np.cos(np.pi * np.arange(W) / W))[None, :] + (2 - 2 * np.cos(np.pi * np.arange(H) / H))[:, None]
    lam[0, 0] = 1.0
    spec = -sf.dctn(D, type=2, norm="ortho") / lam
    spec[0, 0] = 0.0
    return sf.idctn(spec, type=2, norm="ortho")


def solve(pcd, D, path, tol=1e-9, phi0=None):
    h, w = D.shape
    s = pcd.Solver(w, h, 0, path)
    s.upload(D, np.zeros_like(D) if phi0 is None else phi0)
    info = s.run(100000, tol)
    out = s.download()
    s.close()
    return out, info


@pytest.mark.parametrize("shape", [(256, 256), (400, 400), (512, 1024), (300, 157), (157, 301), (64, 2), (1, 9)])
def test_dct_backend_solves_the_discrete_problem(pcd, shape):
    H, W = shape
    rng = np.random.RandomState(H + W)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    D = np.cos(np.pi * (xx + 0.5) / W * 3) * np.cos(np.pi * (yy + 0.5) / H * 2) + 0.1 * rng.standard_normal((H, W))
    D -= D.mean()
    got, info = solve(pcd, D, pcd.SOLVER_DCT, phi0=rng.standard_normal((H, W)))   # the warm start is ignored
    assert info["path"] == "dct" and info["sweeps"] == 0 and info["converged_at"] == 1
    p = np.pad(got, 1, mode="edge")
    lap = p[1:-1, :-2] + p[1:-1, 2:] + p[:-2, 1:-1] + p[2:, 1:-1] - 4 * got
    assert np.abs(lap - D).max() <= 1e-9 * np.abs(D).max()
    assert abs(got.mean()) <= 1e-12 * np.abs(got).max()
    want = direct_cpu(D)
    gx, gy = grad(got)
    wx, wy = grad(want)
    scale = max(np.abs(wx).max(), np.abs(wy).max())
    assert max(np.abs(gx - wx).max(), np.abs(gy - wy).max()) <= 1e-10 * scale


@pytest.mark.parametrize("shape", [(256, 256), (400, 400), (512, 1024)])
def test_dct_backend_vs_converged_sor(pcd, shape):
    H, W = shape
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    D = np.exp(-((xx - 0.3 * W) ** 2 + (yy - 0.6 * H) ** 2) / (0.02 * W * H)) + (np.hypot(xx - 0.7 * W, yy - 0.3 * H) < 0.1 * W)
    D -= D.mean()
    a, _ = solve(pcd, D, pcd.SOLVER_DCT)
    b, ib = solve(pcd, D, pcd.SOLVER_AUTO, tol=1e-9)
    assert ib["converged_at"] > 0
    ax, ay = grad(a)
    bx, by = grad(b)
    scale = max(np.abs(bx).max(), np.abs(by).max())
    assert max(np.abs(ax - bx).max(), np.abs(ay - by).max()) <= 1e-7 * scale


def test_dct_backend_nan_holes_fall_back_to_masked_sweeps(pcd, port, golden):
    D = golden("solver")["nan_D"]
    z = np.zeros_like(D)
    h, w = D.shape
    s = pcd.Solver(w, h, 0, pcd.SOLVER_DCT)
    s.upload(D, z)
    info = s.run(200, 0.0)
    got = s.download()
    s.close()
    assert info["path"] == "streaming" and info["sweeps"] == 200
    assert np.array_equal(got, port.poisson_rb(D, z, 200, 0.0)[0], equal_nan=True)


CONFIGS = {"c1": ("siggraph", 100), "c2": ("lena", 256), "c3": ("hello", 256)}


@pytest.mark.parametrize("key", ["c1", "c2", "c3"])
def test_designs_with_the_dct_backend_match_the_reference(pcd, oracle_mod, golden, key):
    g = golden(f"full_{key}")
    image, res_w = CONFIGS[key]
    O = oracle_mod
    gray = O.rgba_to_gray(golden("images")[image])
    s, resized = O.prepare_image(gray, res_w, O.f32(0.5), O.f32(1.5), O.f32(0.1))
    conv = O.f32(0.01)
    cd = pcd.from_setup(s, solver_path=pcd.SOLVER_DCT)
    cd.initialize_solvers(resized)
    sub = int(g["vertex_sub"][0])

    def vget(name):
        return np.ascontiguousarray(cd.get(name).reshape(s.mesh_ny, s.mesh_nx)[::sub, ::sub]).ravel()

    sx0, sy0 = vget("source_x"), vget("source_y")
    steps = cd.run_transport(50, conv)
    assert cd.last_solve_info()["path"] == "dct"
    ref_steps = g["steps"]
    assert len(steps) == len(ref_steps), (steps, ref_steps)            # transport-iteration count equal
    assert np.abs(np.array(steps) - ref_steps).max() < 1e-4
    disp = max(np.abs(g["target_x"] - sx0).max(), np.abs(g["target_y"] - sy0).max())
    d = max(np.abs(vget("target_x") - g["target_x"]).max(), np.abs(vget("target_y") - g["target_y"]).max())
    assert d <= 1e-3 * disp, (key, d, disp)
    for hi in range(3):
        cd.perform_height_map_iteration(hi)
    z, zr = vget("source_z"), g["source_z"]
    rng = zr.max() - zr.min()
    print(f"{key}: vertices {d / disp:.2e} of max displacement, heights {np.abs(z - zr).max() / rng:.2e} of range")
    assert np.abs(z - zr).max() <= 5e-4 * rng, (key, np.abs(z - zr).max(), rng)
    cd.close()
