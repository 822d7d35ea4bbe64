"""GPU: K-SOR through the C ABI against the oracle.

* bit-exact against the red-black restatement run for the same number of sweeps (both kernel families)
* within the stated tolerance of the reference's lexicographic solver (golden vectors from oracle/_ref):
  gradient rel L-inf <= 1e-5 of max|grad| at tol 1e-7 (SURVEY 8c-i observed 2e-9 at production sizes),
  mean-removed phi abs <= 1e-4
* NaN holes, non-square grids (omega follows W), warm start, max_iterations cap, determinism
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PATHS = ["streaming", "resident", "tiled"]


def path_id(pcd, name):
    return {"streaming": pcd.SOLVER_STREAMING, "resident": pcd.SOLVER_RESIDENT, "tiled": pcd.SOLVER_TILED}[name]


def run_gpu(pcd, D, phi0, max_it, tol, path, lag=0):
    h, w = D.shape
    s = pcd.Solver(w, h, 0, path_id(pcd, path))
    assert s.path == path
    if lag:
        s.set_check_lag(lag)
    s.upload(D, phi0)
    info = s.run(max_it, tol)
    out = s.download()
    s.close()
    return out, info


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("shape", [(48, 48), (24, 64), (50, 20), (2, 3), (1, 9), (150, 300), (301, 157)])
def test_fixed_sweeps_bit_exact_vs_red_black_oracle(pcd, port, path, shape):
    h, w = shape
    rng = np.random.RandomState(h * 1000 + w)
    D = rng.standard_normal((h, w))
    D -= D.mean()
    phi0 = rng.standard_normal((h, w))
    for k in (1, 2, 9):
        got, info = run_gpu(pcd, D, phi0, k, 0.0, path)
        want, n, conv, last = port.poisson_rb(D, phi0, k, 0.0)
        assert info["sweeps"] == k and info["converged_at"] == 0
        assert np.array_equal(got, want), (path, shape, k, np.abs(got - want).max())
        assert info["last_max_update"] == last


@pytest.mark.parametrize("path", PATHS)
def test_convergence_rule_and_lag(pcd, port, golden, path):
    g = golden("solver")
    for name in ("sq48", "rect64x24", "rect20x50"):
        D = g[f"{name}_D"]
        z = np.zeros_like(D)
        got, info = run_gpu(pcd, D, z, 100000, 1e-7, path)
        _, n_exact, conv_exact, _ = port.poisson_rb(D, z, 100000, 1e-7)
        assert info["converged_at"] == conv_exact == n_exact           # same sweep satisfies the test
        extra = info["sweeps"] - info["converged_at"]
        assert 0 <= extra <= 128        # large-grid path: blocks of 64 sweeps, the stop is acted on one block late
        want, n, conv, _ = port.poisson_rb(D, z, 100000, 1e-7, extra_sweeps=extra)
        assert n == info["sweeps"] and np.array_equal(got, want)
        assert info["last_max_update"] < 1e-7
        # against the reference's own (lexicographic) answer
        lex = g[f"{name}_phi_conv"]
        gl, gg = port.gradient(lex), port.gradient(got)
        scale = max(np.abs(gl[0]).max(), np.abs(gl[1]).max())
        assert max(np.abs(gl[0] - gg[0]).max(), np.abs(gl[1] - gg[1]).max()) / scale < 1e-5
        assert np.abs((lex - lex.mean()) - (got - got.mean())).max() < 1e-4


@pytest.mark.parametrize("path", PATHS)
def test_warm_start_and_cap(pcd, port, golden, path):
    g = golden("solver")
    D = g["sq48_D"]
    conv = g["sq48_phi_conv"]
    got, info = run_gpu(pcd, D, conv, 100000, 1e-9, path)             # phi is in/out: continue from a solution
    extra = info["sweeps"] - info["converged_at"]
    want, _, _, _ = port.poisson_rb(D, conv, 100000, 1e-9, extra_sweeps=extra)
    assert np.array_equal(got, want) and info["converged_at"] > 0
    got, info = run_gpu(pcd, D, np.zeros_like(D), 13, 1e-30, path)   # cap binds
    assert info["sweeps"] == 13 and info["converged_at"] == 0
    assert np.array_equal(got, port.poisson_rb(D, np.zeros_like(D), 13, 1e-30)[0])
    got, info = run_gpu(pcd, D, conv, 0, 1e-7, path)                  # max_iterations = 0: untouched
    assert info["sweeps"] == 0 and np.array_equal(got, conv)


@pytest.mark.parametrize("path", PATHS)
def test_nan_holes(pcd, port, golden, path):
    g = golden("solver")
    D = g["nan_D"]
    z = np.zeros_like(D)
    for k in (5, 200):
        got, info = run_gpu(pcd, D, z, k, 0.0, path)
        want = port.poisson_rb(D, z, k, 0.0)[0]
        assert np.array_equal(got, want, equal_nan=True)
        assert np.isnan(got).sum() == np.isnan(D).sum()               # holes stay holes, nothing leaks
    # same rule as the reference: compare with its lexicographic result after many sweeps
    got, _ = run_gpu(pcd, D, z, 3000, 0.0, path)
    lex = port.poisson_lex(D, z, 3000, 0.0)[0]
    m = np.isfinite(lex)
    assert np.array_equal(m, np.isfinite(got))
    assert np.abs((lex[m] - lex[m].mean()) - (got[m] - got[m].mean())).max() < 1e-6


def test_paths_agree_bit_for_bit_and_are_deterministic(pcd):
    rng = np.random.RandomState(0)
    D = rng.standard_normal((256, 256))
    D -= D.mean()
    z = np.zeros_like(D)
    a, ia = run_gpu(pcd, D, z, 300, 0.0, "streaming")
    b, ib = run_gpu(pcd, D, z, 300, 0.0, "resident")
    c, _ = run_gpu(pcd, D, z, 300, 0.0, "resident")
    assert np.array_equal(a, b) and np.array_equal(b, c)
    assert ia["last_max_update"] == ib["last_max_update"]


def test_drop_in_signature(pcd, port):
    """poisson_solver(D, phi, width, height, max_iterations, tol, max_threads), src/solver.h:8."""
    rng = np.random.RandomState(1)
    D = rng.standard_normal((40, 56))
    D -= D.mean()
    phi = np.zeros_like(D)
    info = pcd.poisson_solver(D, phi, 56, 40, 100000, 1e-7, 4)
    assert info["converged_at"] > 0
    want = port.poisson_rb(D, np.zeros_like(D), 100000, 1e-7, extra_sweeps=info["sweeps"] - info["converged_at"])[0]
    assert np.array_equal(phi, want)


def test_full_size_properties_1024(pcd):
    """BASELINE.json configs[3] size: properties that need no CPU solve -- the converged field
    satisfies the discrete equation (residual <= tol*cnt/omega), result independent of the kernel family."""
    n = 1024
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float64)
    D = np.cos(np.pi * (xx + 0.5) / n * 3) * np.cos(np.pi * (yy + 0.5) / n * 2) * 1e-3
    D -= D.mean()
    s = pcd.Solver(n, n, 0, pcd.SOLVER_AUTO)
    assert s.path == "resident"
    s.upload(D, np.zeros_like(D))
    info = s.run(100000, 1e-7)
    phi = s.download()
    s.close()
    assert 0 < info["converged_at"] < 100000
    p = np.pad(phi, 1, mode="edge")
    lap = p[1:-1, :-2] + p[1:-1, 2:] + p[:-2, 1:-1] + p[2:, 1:-1] - 4 * phi      # Neumann by dropped neighbours
    omega = 2.0 / (1.0 + 3.14159265 / n)
    # delta = omega/cnt * residual at update time; later updates of the same sweep move it by O(cnt*delta), and the
    # max|delta| of SOR near omega = 2 is not monotone: the solver runs a few sweeps past the one that met the rule
    # (info["sweeps"] - info["converged_at"]), during which the residual wanders within a small multiple of it
    assert np.abs(lap - D).max() <= 1e-7 * 4 / omega * 40
    s2 = pcd.Solver(n, n, 0, pcd.SOLVER_STREAMING)
    s2.upload(D, np.zeros_like(D))
    s2.run(info["sweeps"], 0.0)
    assert np.array_equal(s2.download(), phi)
    s2.close()


def test_full_size_properties_8192(pcd):
    """BASELINE.json configs[4] size (no CPU solve possible): the wavefront kernel equals the per-colour kernels bit
    for bit after the same number of sweeps (odd count: exercises the single-sweep pass), and the solve is exactly
    linear under scaling by a power of two (every operation of the update scales exactly)."""
    n = 8192
    rng = np.random.RandomState(5)
    D = rng.standard_normal((n, n))
    D -= D.mean()
    phi0 = rng.standard_normal((n, n))
    s = pcd.Solver(n, n, 0, pcd.SOLVER_AUTO)
    assert s.path == "tiled"
    s.upload(D, phi0)
    info = s.run(7, 0.0)
    a = s.download()
    s.upload(4.0 * D, 4.0 * phi0)
    info4 = s.run(7, 0.0)
    a4 = s.download()
    s.close()
    assert info["sweeps"] == 7 and info4["last_max_update"] == 4.0 * info["last_max_update"]
    assert np.array_equal(a4, 4.0 * a)
    s2 = pcd.Solver(n, n, 0, pcd.SOLVER_STREAMING)
    s2.upload(D, phi0)
    info2 = s2.run(7, 0.0)
    assert np.array_equal(s2.download(), a)
    assert info2["last_max_update"] == info["last_max_update"]
    s2.close()


@pytest.mark.parametrize("shape", [(1024, 1024), (1000, 1000), (777, 1024), (1024, 600), (333, 1000), (1024, 1023), (149, 1024),
                                   (296, 64), (1036, 998)])
def test_resident_kernel_at_production_sizes_bit_exact(pcd, shape):
    """The on-chip kernel (CTA pairs over DSMEM, uneven row split, warp-level barriers, stop announced ahead) against the
    per-colour kernels at sizes where all 148 CTAs and all 16 warps take part: same bits after 1, 2 and 37 sweeps and
    after a run to convergence (same converged_at, same field for the same number of sweeps)."""
    H, W = shape
    rng = np.random.RandomState(H + W)
    D = rng.standard_normal((H, W)) * 1e-3
    D -= D.mean()
    phi0 = rng.standard_normal((H, W))
    r = pcd.Solver(W, H, 0, pcd.SOLVER_RESIDENT)
    s = pcd.Solver(W, H, 0, pcd.SOLVER_STREAMING)
    for n in (1, 2, 37):
        r.upload(D, phi0)
        s.upload(D, phi0)
        ir, is_ = r.run(n, 0.0), s.run(n, 0.0)
        assert ir["sweeps"] == n and np.array_equal(r.download(), s.download()), n
        assert ir["last_max_update"] == is_["last_max_update"], n
    z = np.zeros_like(D)
    r.upload(D, z)
    ir = r.run(100000, 1e-6)
    assert 0 < ir["converged_at"] <= ir["sweeps"] <= ir["converged_at"] + 64
    s.upload(D, z)
    is_ = s.run(ir["sweeps"], 0.0)
    assert np.array_equal(r.download(), s.download())
    s.upload(D, z)
    assert s.run(100000, 1e-6)["converged_at"] == ir["converged_at"]
    r.close(); s.close()


@pytest.mark.parametrize("sweeps", [1, 37])
def test_resident_1024_bit_exact_vs_oracle(pcd, port, sweeps):
    """BASELINE.json configs[3] size, pinned to the ORACLE (not to another CUDA kernel): the on-chip kernel with all
    148 CTAs / CTA pairs / warp-level barriers against pcdo_poisson_rb after 1 and 37 sweeps."""
    n = 1024
    rng = np.random.RandomState(1024 + sweeps)
    D = rng.standard_normal((n, n))
    D -= D.mean()
    phi0 = rng.standard_normal((n, n))
    got, info = run_gpu(pcd, D, phi0, sweeps, 0.0, "resident")
    want, k, conv, last = port.poisson_rb(D, phi0, sweeps, 0.0)
    assert info["sweeps"] == sweeps == k and info["path"] == "resident"
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert info["last_max_update"] == last


def test_wavefront_8192_bit_exact_vs_oracle(pcd, port):
    """BASELINE.json configs[4] size, pinned to the ORACLE: the wavefront kernel (TS=2 passes, odd counts exercise the
    single-sweep pass) against pcdo_poisson_rb after 3 and 7 sweeps; ~10 s of CPU."""
    n = 8192
    rng = np.random.RandomState(8192)
    D = rng.standard_normal((n, n))
    D -= D.mean()
    phi0 = rng.standard_normal((n, n))
    s = pcd.Solver(n, n, 0, pcd.SOLVER_AUTO)
    assert s.path == "tiled"
    for sweeps in (3, 7):
        s.upload(D, phi0)
        info = s.run(sweeps, 0.0)
        got = s.download()
        want, k, conv, last = port.poisson_rb(D, phi0, sweeps, 0.0)
        assert info["sweeps"] == sweeps == k
        assert np.array_equal(got, want), (sweeps, np.abs(got - want).max())
        assert info["last_max_update"] == last
    s.close()


def test_check_lag_is_clamped(pcd, port):
    """a check_lag beyond the pinned mirror (4096 entries) must not overflow it."""
    rng = np.random.RandomState(2)
    D = rng.standard_normal((64, 64))
    D -= D.mean()
    for path in ("resident", "streaming", "tiled"):
        got, info = run_gpu(pcd, D, np.zeros_like(D), 5000, 0.0, path, lag=1 << 20)
        assert info["sweeps"] == 5000
        assert np.array_equal(got, port.poisson_rb(D, np.zeros_like(D), 5000, 0.0)[0])
    s = pcd.Solver(64, 64, 0, pcd.SOLVER_RESIDENT)
    s.set_check_lag(-7)          # negative: the default
    s.upload(D, np.zeros_like(D))
    assert s.run(100000, 1e-7)["converged_at"] > 0
    s.close()


@pytest.mark.parametrize("shape", [(400, 400), (444, 64), (298, 36), (297, 66), (600, 202), (740, 6), (1036, 1024), (1024, 1024),
                                   (512, 130), (900, 1000), (1100, 600), (1332, 64), (1300, 1024), (157, 300), (96, 96), (5, 8),
                                   (720, 1280), (400, 1100), (1024, 1332), (64, 1200), (4, 1030)])
def test_resident_deep_halo_kernel_bit_exact(pcd, port, shape, monkeypatch):
    """The resident kernel with ONE neighbour exchange per sweep (redundant update of the colour-0 cells of the rows just
    outside a slab, double-buffered messages, lane shuffles + edge-lane polls for the halo row's left/right cells):
    pinned to the oracle after 1, 2, 3, 10 and 41 sweeps, to the exchange-per-phase kernel after a converged run, on
    shapes with slabs of 2..9 rows, one to 512 column pairs, partly filled warps and single-lane warps, and on grids
    wider than 1024 columns, which are solved transposed (summation order switched so the bits stay the reference's)."""
    H, W = shape
    rng = np.random.RandomState(7 * H + W)
    D = rng.standard_normal((H, W))
    D -= D.mean()
    phi0 = rng.standard_normal((H, W))
    s = pcd.Solver(W, H, 0, pcd.SOLVER_RESIDENT)
    for n in (1, 2, 3, 10, 41):
        s.upload(D, phi0)
        info = s.run(n, 0.0)
        assert s.resident_exchange == 2 and info["launches"] == (1 if W <= 1024 else 4)   # + three transposes
        want, k, conv, last = port.poisson_rb(D, phi0, n, 0.0)
        got = s.download()
        assert np.array_equal(got, want), (shape, n, np.abs(got - want).max())
        assert info["sweeps"] == n and info["last_max_update"] == last
    z = np.zeros_like(D)
    D2 = D * 1e-3
    s.upload(D2, z)
    deep = s.run(100000, 1e-6)
    f_deep = s.download()
    assert s.resident_exchange == 2
    if deep["converged_at"] == 0:   # 740 x 6: omega follows the WIDTH (src/solver.cpp:71), the long axis needs > 10^5 sweeps
        assert shape == (740, 6) and deep["sweeps"] == 100000
    else:
        assert 0 < deep["converged_at"] <= deep["sweeps"] <= deep["converged_at"] + 64
    monkeypatch.setenv("PCD_RES_NO_DEEP", "1")
    s.upload(D2, z)
    classic = s.run(100000, 1e-6)
    # strips of 8-9 rows (H > 1036) and transposed solves (W > 1024) exist only in the deep-halo kernel: without it the
    # solve runs on the wavefront path
    phase_fits = H <= 1036 and W <= 1024
    assert s.resident_exchange == (1 if phase_fits else 0)
    assert classic["converged_at"] == deep["converged_at"]
    if phase_fits:
        assert classic["sweeps"] == deep["sweeps"]
        assert np.array_equal(s.download(), f_deep)
        assert classic["last_max_update"] == deep["last_max_update"]
    s.close()


def test_resident_tall_strips_hand_nan_holes_to_the_large_grid_path(pcd, port):
    """1100 rows = 8 rows per CTA: on chip only with the deep-halo kernel, which does not take NaN holes -- the solve then
    runs on the masked large-grid path, same bits as the oracle."""
    H, W = 1100, 256
    rng = np.random.RandomState(11)
    D = rng.standard_normal((H, W))
    phi0 = rng.standard_normal((H, W))
    s = pcd.Solver(W, H, 0, pcd.SOLVER_AUTO)
    s.upload(D, phi0)
    info = s.run(9, 0.0)
    assert s.resident_exchange == 2 and info["path"] == "resident"
    assert np.array_equal(s.download(), port.poisson_rb(D, phi0, 9, 0.0)[0])
    D[500:520, 100:130] = np.nan
    s.upload(D, phi0)
    info = s.run(9, 0.0)
    assert s.resident_exchange == 0 and info["path"] == "streaming" and info["sweeps"] == 9
    assert np.array_equal(s.download(), port.poisson_rb(D, phi0, 9, 0.0)[0], equal_nan=True)
    s.close()


def test_resident_deep_halo_kernel_hands_nan_holes_back(pcd, port):
    """NaN holes need the masked neighbour rule: the deep-halo kernel finds them while loading, voids its launch without
    touching the field, and the exchange-per-phase kernel runs instead."""
    H, W = 800, 400   # six rows per CTA: the deep-halo kernel is the default choice
    rng = np.random.RandomState(5)
    D = rng.standard_normal((H, W))
    D[100:140, 200:260] = np.nan
    D[799, 399] = np.nan
    phi0 = rng.standard_normal((H, W))
    s = pcd.Solver(W, H, 0, pcd.SOLVER_RESIDENT)
    s.upload(D, phi0)
    info = s.run(25, 0.0)
    assert s.resident_exchange == 1 and info["launches"] == 2 and info["sweeps"] == 25
    want = port.poisson_rb(D, phi0, 25, 0.0)[0]
    assert np.array_equal(s.download(), want, equal_nan=True)
    D[np.isnan(D)] = 0.5
    s.upload(D, phi0)
    info = s.run(25, 0.0)
    assert s.resident_exchange == 2 and info["launches"] == 1
    assert np.array_equal(s.download(), port.poisson_rb(D, phi0, 25, 0.0)[0])
    s.close()
    # the default choice: the deep-halo kernel wherever it is valid (even width, at least three rows; short grids get
    # fewer CTAs of three rows each), the exchange-per-phase kernel for odd widths and grids of one or two rows
    for (H2, W2, want_kernel) in ((400, 400, 2), (296, 64, 2), (157, 300, 2), (2, 64, 1), (400, 301, 1)):
        s = pcd.Solver(W2, H2, 0, pcd.SOLVER_RESIDENT)
        D = rng.standard_normal((H2, W2))
        s.upload(D, np.zeros_like(D))
        s.run(5, 0.0)
        assert s.resident_exchange == want_kernel, (H2, W2)
        assert np.array_equal(s.download(), port.poisson_rb(D, np.zeros_like(D), 5, 0.0)[0])
        s.close()


_LAUNCH_MODE_SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import poisson_caustic_design_b200 as P
from oracle import oracle as O
port = O.OracleLib()
for (H, W) in ((600, 202), (1024, 1024)):
    rng = np.random.RandomState(H + W)
    D = rng.standard_normal((H, W)); D -= D.mean()
    phi0 = rng.standard_normal((H, W))
    s = P.Solver(W, H, 0, P.SOLVER_RESIDENT)
    s.upload(D, phi0)
    info = s.run(23, 0.0)
    assert s.resident_exchange == int(sys.argv[2]), s.resident_exchange
    assert np.array_equal(s.download(), port.poisson_rb(D, phi0, 23, 0.0)[0]), (H, W)
    s.close()
print("ok")
"""


@pytest.mark.parametrize("env,kernel", [({"PCD_RES_NO_PAIRS": "1"}, 2), ({"PCD_RES_NO_PAIRS": "1", "PCD_RES_NO_DEEP": "1"}, 1),
                                        ({"PCD_RES_CLUSTER": "4"}, 2)])
def test_resident_kernels_other_launch_modes(pcd, env, kernel):
    """The cooperative launch without CTA pairs (every link through L2) and clusters of four are read from the environment
    once per process, so they run in a child process: both resident kernels against the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _LAUNCH_MODE_SCRIPT, root, str(kernel)], env={**os.environ, **env},
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
