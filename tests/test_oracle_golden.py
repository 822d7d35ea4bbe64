"""CPU: the plain-C restatement (oracle/pcd_oracle.c) against the golden vectors that
tests/golden/make_golden.py generated from the UNMODIFIED reference (oracle/_ref).  Bit-exact."""
import numpy as np
import pytest

from conftest import setup_from_params

STAGE_CASES = ["sq16", "sq17", "rect24x12", "rect24x8"]


def test_solver_lexicographic_known_answers(port, golden):
    g = golden("solver")
    for name in ("sq48", "rect64x24", "rect20x50", "tiny3x2", "row1x9"):
        D = g[f"{name}_D"]
        for k in (1, 2, 7):
            phi, n, _ = port.poisson_lex(D, np.zeros_like(D), k, 0.0)
            assert n == k
            assert np.array_equal(phi, g[f"{name}_phi_k{k}"]), (name, k)
        phi, n, last = port.poisson_lex(D, np.zeros_like(D), 100000, 1e-7)
        assert np.array_equal(phi, g[f"{name}_phi_conv"]), name
        assert last < 1e-7 and n < 100000
        warm, n2, _ = port.poisson_lex(D, phi, 100000, 1e-9)       # phi is in/out (src/solver.cpp:85-90)
        assert np.array_equal(warm, g[f"{name}_phi_warm"]), name


def test_solver_nan_holes_and_cap(port, golden):
    g = golden("solver")
    D = g["nan_D"]
    for k in (5, 200):
        phi, _, _ = port.poisson_lex(D, np.zeros_like(D), k, 0.0)
        assert np.array_equal(phi, g[f"nan_phi_k{k}"], equal_nan=True)
    phi, n, last = port.poisson_lex(g["sq48_D"], np.zeros((48, 48)), 13, 1e-30)
    assert n == 13 and last > 1e-30
    assert np.array_equal(phi, g["cap_phi"])


def test_red_black_agrees_with_lexicographic(port, golden):
    """Same omega / stopping rule, different ordering: gradients agree to ~1e-7 relative."""
    g = golden("solver")
    D = g["sq48_D"]
    lex, n_lex, _ = port.poisson_lex(D, np.zeros_like(D), 100000, 1e-7)
    rb, n_rb, conv, _ = port.poisson_rb(D, np.zeros_like(D), 100000, 1e-7)
    assert conv == n_rb and abs(n_rb - n_lex) <= 0.15 * n_lex
    gl, gr = port.gradient(lex), port.gradient(rb)
    scale = max(np.abs(gl[0]).max(), np.abs(gl[1]).max())
    assert np.abs(gl[0] - gr[0]).max() / scale < 1e-5
    assert np.abs((lex - lex.mean()) - (rb - rb.mean())).max() < 1e-4
    # extra sweeps after convergence are executed verbatim
    rb2, n2, conv2, _ = port.poisson_rb(D, np.zeros_like(D), 100000, 1e-7, extra_sweeps=5)
    assert conv2 == conv and n2 == conv + 5


@pytest.mark.parametrize("case", STAGE_CASES)
def test_pipeline_stages_bit_exact(port, golden, oracle_mod, case):
    g = golden(f"stages_{case}")
    s = setup_from_params(oracle_mod, g["params"])
    d = port.design(s, solver_mode=0)
    d.initialize_solvers(g["image"])
    assert np.array_equal(d.get("pixels"), g["pixels"])
    assert np.array_equal(d.get("target_areas"), g["target_areas"])
    it = 0
    while f"it{it}_step" in g:
        step = d.transport_iteration()
        assert step == g[f"it{it}_step"][0]
        for f in ("errors", "raster", "phi", "vertex_gradient_x", "vertex_gradient_y", "target_x", "target_y"):
            assert np.array_equal(d.get(f), g[f"it{it}_{f}"]), (case, it, f)
        it += 1
    ix, iy = d.inverted_transport_map()
    assert np.array_equal(ix, g["inverted_x"]) and np.array_equal(iy, g["inverted_y"])
    for hi in range(3):
        d.height_iteration(hi)
        for f in ("normals_x", "normals_y", "norm_x", "norm_y", "divergence", "h", "source_z"):
            assert np.array_equal(d.get(f), g[f"h{hi}_{f}"]), (case, hi, f)
    d.close()


def test_c1_first_iterations_match_reference_run(port, golden, oracle_mod):
    """BASELINE.json configs[0] (siggraph, res_w=100): first two transport iterations of the
    restatement against the reference's own run (13 iterations; BASELINE.md anchors)."""
    g = golden("full_c1")
    img = oracle_mod.rgba_to_gray(golden("images")["siggraph"])
    s, resized = oracle_mod.prepare_image(img, 100, oracle_mod.f32(0.5), oracle_mod.f32(1.5), oracle_mod.f32(0.1))
    assert [s.mesh_nx, s.mesh_ny, s.res_x, s.res_y] == [int(v) for v in g["params"][:4]]
    d = port.design(s)
    d.initialize_solvers(resized)
    assert np.array_equal(d.get("target_areas"), g["target_areas"])
    assert abs(d.get("target_areas").sum() - 0.25) < 1e-12
    steps = [d.transport_iteration() for _ in range(2)]
    assert steps[0] == g["steps"][0] and steps[1] == g["steps"][1]
    assert d.last_sweeps == 4098                                   # SURVEY 8c anchor (second solve)
    assert len(g["steps"]) == 13 and abs(g["steps"][0] - 0.065981381) < 1e-9
    assert np.array_equal(d.get("target_x"), g["target_x_it0"]) is False  # moved on to iteration 1
    d.close()
