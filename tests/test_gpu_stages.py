"""GPU: every stage kernel fed the reference's own inputs for that stage (golden dumps from oracle/_ref),
so per-kernel error is isolated from trajectory drift.  Tolerance: rel L-inf <= 1e-12 for
stages that only reassociate fp64 sums; the rasteriser is compared where the mesh has no folds."""
import numpy as np
import pytest

from conftest import rel_linf, setup_from_params

pytestmark = pytest.mark.gpu
CASES = ["sq16", "sq17", "rect24x12", "rect24x8"]
TOL = 1e-12


def make(pcd, oracle_mod, g):
    s = setup_from_params(oracle_mod, g["params"])
    cd = pcd.from_setup(s)
    cd.initialize_solvers(g["image"])
    return s, cd


@pytest.mark.parametrize("case", CASES)
def test_init_pixels_mesh_target_areas(pcd, oracle_mod, golden, case):
    g = golden(f"stages_{case}")
    s, cd = make(pcd, oracle_mod, g)
    assert np.array_equal(cd.get("pixels"), g["pixels"])                      # same expression, no sums
    assert np.array_equal(cd.get("target_x"), g["it0_pre_target_x"])
    assert np.array_equal(cd.get("target_y"), g["it0_pre_target_y"])
    assert np.array_equal(cd.get("source_x"), g["it0_pre_target_x"])
    ta = cd.get("target_areas")
    assert rel_linf(ta, g["target_areas"]) < TOL                              # K-TAREA (Sutherland-Hodgman)
    assert abs(ta.sum() - s.width * s.height) < 1e-12
    assert not cd.get("phi").any() and not cd.get("h").any()
    cd.close()


@pytest.mark.parametrize("case", CASES)
def test_transport_stages_from_reference_state(pcd, oracle_mod, port, golden, case):
    g = golden(f"stages_{case}")
    s, cd = make(pcd, oracle_mod, g)
    cd.set("target_areas", g["target_areas"])
    it = 0
    while f"it{it}_step" in g:
        cd.set("target_x", g[f"it{it}_pre_target_x"])
        cd.set("target_y", g[f"it{it}_pre_target_y"])
        cd.stage_errors()                                                     # K-AREA + K-ERR
        assert rel_linf(cd.get("errors"), g[f"it{it}_errors"]) < TOL
        cd.set("errors", g[f"it{it}_errors"])
        cd.stage_raster()                                                     # K-RAST
        cd.stage_subtract_average()                                           # K-MEAN
        scale = np.abs(g[f"it{it}_raster"]).max()
        assert np.abs(cd.get("raster") - g[f"it{it}_raster"]).max() / scale < 1e-11, (case, it)
        # K-SOR on the reference's raster, warm-started like the reference
        cd.set("raster", g[f"it{it}_raster"])
        cd.set("phi", g[f"it{it}_pre_phi"])
        info = cd.stage_solve_transport()
        want = port.poisson_rb(g[f"it{it}_raster"], g[f"it{it}_pre_phi"], 100000, 1e-7,
                               extra_sweeps=info["sweeps"] - info["converged_at"])[0]
        assert np.array_equal(cd.get("phi"), want)
        # K-STEP from the reference's phi
        cd.set("phi", g[f"it{it}_phi"])
        step = cd.stage_step()
        assert np.array_equal(cd.get("vertex_gradient_x"), g[f"it{it}_vertex_gradient_x"])
        assert np.array_equal(cd.get("vertex_gradient_y"), g[f"it{it}_vertex_gradient_y"])
        assert np.array_equal(cd.get("target_x"), g[f"it{it}_target_x"])
        assert np.array_equal(cd.get("target_y"), g[f"it{it}_target_y"])
        assert step == g[f"it{it}_step"][0]
        if f"it{it}_gradient_x" in g:
            assert np.array_equal(cd.get("gradient_x"), g[f"it{it}_gradient_x"])
            assert np.array_equal(cd.get("gradient_y"), g[f"it{it}_gradient_y"])
        it += 1
    cd.close()


@pytest.mark.parametrize("case", CASES)
def test_height_stages_from_reference_state(pcd, oracle_mod, golden, case):
    g = golden(f"stages_{case}")
    s, cd = make(pcd, oracle_mod, g)
    last = max(int(k[2:k.index("_")]) for k in g if k.startswith("it") and k.endswith("_step"))
    cd.set("target_x", g[f"it{last}_target_x"])
    cd.set("target_y", g[f"it{last}_target_y"])
    ix, iy = cd.inverted_transport_map()                                      # K-INV
    assert np.abs(ix - g["inverted_x"]).max() < 1e-12 and np.abs(iy - g["inverted_y"]).max() < 1e-12
    for hi in range(3):
        cd.set("h", g[f"h{hi}_pre_h"])
        cd.set("source_z", g[f"h{hi}_pre_source_z"])
        cd.perform_height_map_iteration(hi)
        for f, tol in (("normals_x", 1e-11), ("normals_y", 1e-11), ("norm_x", 1e-11), ("norm_y", 1e-11)):
            assert rel_linf(cd.get(f), g[f"h{hi}_{f}"]) < tol, (case, hi, f)
        assert np.abs(cd.get("divergence") - g[f"h{hi}_divergence"]).max() < 1e-11 * np.abs(g[f"h{hi}_norm_x"]).max()
        # heights: red-black vs the reference's lexicographic ordering at tol 1e-8 -> compare z (min-shifted)
        z, zr = cd.get("source_z"), g[f"h{hi}_source_z"]
        assert np.abs(z - zr).max() <= 2e-5 * max(zr.max() - zr.min(), 1e-30) + 1e-6, (case, hi, np.abs(z - zr).max())
    cd.close()


def test_raster_miss_is_an_error(pcd, oracle_mod, golden):
    """The reference prints 'interpolation miss!' and exit(0)s (src/mesh.cpp:276-281); here: status code."""
    g = golden("stages_sq16")
    s, cd = make(pcd, oracle_mod, g)
    x = g["it0_pre_target_x"].copy()
    x[:] = x * 0.5                                                            # mesh no longer covers the domain
    cd.set("target_x", x)
    cd.stage_errors()
    with pytest.raises(pcd.PcdError) as e:
        cd.stage_raster()
    assert e.value.status == pcd.PCD_ERR_RASTER_MISS
    cd.close()


def test_call_order_errors(pcd):
    cd = pcd.CausticDesign()
    cd.set_mesh_resolution(8, 8)
    cd.set_domain_resolution(32, 32)
    cd.set_mesh_size(1.0, 1.0)
    with pytest.raises(pcd.PcdError):
        cd.initialize_solvers(np.zeros((16, 16)))                             # wrong image shape
    bad = pcd.CausticDesign()
    with pytest.raises(pcd.PcdError):
        bad.initialize_solvers(np.zeros((0, 0)))
