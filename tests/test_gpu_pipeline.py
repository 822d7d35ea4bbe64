"""GPU: the whole path (init -> transport loop -> 3 height iterations) through the C ABI against the
reference's own end-to-end runs (tests/golden/full_c*.npz from oracle/_ref, threads=1).

Stated tolerances (BASELINE.json north_star: relative L-inf on heights and vertex displacement, iteration
count +-1):
* transport-iteration count equal (+-1 allowed), per-iteration step sizes within 1e-5 absolute;
* vertex positions: rel L-inf <= 1e-6 of the max displacement while the mesh has no folds (iteration 0
  everywhere, iteration 5 for C1/C2; C3's black background folds the mesh earlier), <= 1e-3 at the end
  (rasteriser tie-break under folds: lowest triangle index here, BVH traversal order in the reference;
  SURVEY App. B measured 2.4-3.4e-4 from the tie-break alone);
* heights (source z): rel L-inf <= 5e-4 of their range.  This is the truncation error of the reference's own
  stopping rule (max|delta| < 1e-8), not kernel error: the CPU oracle itself, switched from lexicographic to
  red-black ordering with everything else identical, differs from the reference by 1.57e-4 of the range on C1
  (4.614e-6 absolute; the CUDA path: 4.618e-6) while vertices agree to 1.6e-9 ("Parity")."""
import os

import numpy as np
import pytest

from conftest import GOLD

pytestmark = pytest.mark.gpu

CONFIGS = {
    "c1": ("siggraph", 100),   # BASELINE.json configs[0]
    "c2": ("lena", 256),       # configs[1]
    "c3": ("hello", 256),      # configs[2], non-square
}


@pytest.mark.parametrize("key", ["c1", "c2", "c3"])
def test_end_to_end_matches_reference(pcd, oracle_mod, golden, key):
    path = os.path.join(GOLD, f"full_{key}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated yet")
    g = golden(f"full_{key}")
    image, res_w = CONFIGS[key]
    O = oracle_mod
    gray = O.rgba_to_gray(golden("images")[image])
    s, resized = O.prepare_image(gray, res_w, O.f32(0.5), O.f32(1.5), O.f32(0.1))
    conv = O.f32(0.01)
    cd = pcd.from_setup(s)
    cd.initialize_solvers(resized)
    sub = int(g["vertex_sub"][0])          # big fixtures keep every sub-th vertex in x and y

    def vget(name):
        return np.ascontiguousarray(cd.get(name).reshape(s.mesh_ny, s.mesh_nx)[::sub, ::sub]).ravel()

    def vsub(a):
        return np.ascontiguousarray(np.asarray(a).reshape(s.mesh_ny, s.mesh_nx)[::sub, ::sub]).ravel()

    assert np.abs(vget("target_areas") - g["target_areas"]).max() < 1e-12 * g["target_areas"].max()
    assert abs(cd.get("target_areas").sum() - g["target_areas_sum"][0]) < 1e-12
    sx0, sy0 = vget("source_x"), vget("source_y")
    steps = []
    for itr in range(50):
        step = cd.perform_transport_iteration()
        steps.append(step)
        if itr in (0, 5) and f"target_x_it{itr}" in g:
            d = max(np.abs(vget("target_x") - g[f"target_x_it{itr}"]).max(),
                    np.abs(vget("target_y") - g[f"target_y_it{itr}"]).max())
            disp = max(np.abs(g[f"target_x_it{itr}"] - sx0).max(), np.abs(g[f"target_y_it{itr}"] - sy0).max())
            tol = 1e-3 if (key == "c3" and itr == 5) else 1e-6
            assert d <= tol * disp, (key, itr, d, disp)
        if step < conv:
            break
    ref_steps = g["steps"]
    assert abs(len(steps) - len(ref_steps)) <= 1, (steps, ref_steps)
    n = min(len(steps), len(ref_steps))
    assert np.abs(np.array(steps[:n]) - ref_steps[:n]).max() < 1e-5
    if len(steps) == len(ref_steps):
        tx, ty = vget("target_x"), vget("target_y")
        disp = max(np.abs(g["target_x"] - sx0).max(), np.abs(g["target_y"] - sy0).max())
        d = max(np.abs(tx - g["target_x"]).max(), np.abs(ty - g["target_y"]).max())
        assert d <= 1e-3 * disp, (key, d, disp)
        ix, iy = cd.inverted_transport_map()
        assert max(np.abs(vsub(ix) - g["inverted_x"]).max(), np.abs(vsub(iy) - g["inverted_y"]).max()) <= 5e-3 * s.width
    for hi in range(3):
        cd.perform_height_map_iteration(hi)
    if len(steps) == len(ref_steps):
        z, zr = vget("source_z"), g["source_z"]
        rng = zr.max() - zr.min()
        print(f"{key}: heights {np.abs(z - zr).max() / rng:.3e} of range, vertices {d / disp:.3e} of max displacement")
        assert np.abs(z - zr).max() <= 5e-4 * rng, (key, np.abs(z - zr).max(), rng)
        h = cd.get("h")
        hs, hr = h[::8, ::8], g["h_sub8"]
        assert np.abs((hs - hs.mean()) - (hr - hr.mean())).max() <= 1e-3 * (g["h_range"][1] - g["h_range"][0])
    cd.close()


def test_c4_first_iterations_match_reference(pcd, golden):
    """BASELINE.json configs[3] -- the workload bench.py's metric is quoted on -- pinned to the reference itself:
    tests/golden/c4_first_iterations.npz holds the first two transport iterations of oracle/_ref (threads=1) on
    synth_density(1024, 1024, 1024).  Fold-free at this point.  Iteration 0 sees identical inputs: errors / raster agree to
    1e-11 (fp64 reassociation); from then on the vertices carry the difference between the two sweep orderings' stopping
    points (1e-7 of the fields).  Steps rel <= 1e-6, vertices <= 1e-6 of the max displacement, grad(phi) rel L-inf <= 1e-6."""
    path = os.path.join(GOLD, "c4_first_iterations.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated yet")
    import hashlib
    from poisson_caustic_design_b200 import synth
    g = golden("c4_first_iterations")
    img = synth.synth_density(1024, 1024, 1024)
    assert hashlib.md5(img.tobytes()).hexdigest().encode() == g["image_md5"].tobytes()
    st = synth.Setup(256, 1024, 1024, mesh_width=1.0, focal_l=1.5, thickness=0.2)
    cd = pcd.from_setup(st)
    cd.initialize_solvers(img)
    ny, nx = st.mesh_ny, st.mesh_nx

    def vsub(name):
        return np.ascontiguousarray(cd.get(name).reshape(ny, nx)[::2, ::2]).ravel()

    ta = cd.get("target_areas")
    assert abs(ta.sum() - g["target_areas_sum"][0]) < 1e-12
    assert np.abs(vsub("target_areas") - g["target_areas_sub2"]).max() < 1e-12 * g["target_areas_sub2"].max()
    sx, sy = g["source_x_sub2"], g["source_y_sub2"]
    for it in range(len(g["steps"])):
        step = cd.perform_transport_iteration()
        info = cd.last_solve_info()
        assert info["path"] == "resident"
        assert abs(step - g["steps"][it]) <= 1e-6 * g["steps"][it], (it, step, g["steps"][it])
        ftol = 1e-11 if it == 0 else 1e-7
        assert np.abs(vsub("errors") - g[f"it{it}_errors_sub2"]).max() <= ftol * np.abs(g[f"it{it}_errors_sub2"]).max()
        ras = cd.get("raster")[::8, ::8]
        assert np.abs(ras - g[f"it{it}_raster_sub8"]).max() <= ftol * np.abs(g[f"it{it}_raster_sub8"]).max()
        gx, gy = cd.get("gradient_x")[::8, ::8], cd.get("gradient_y")[::8, ::8]
        gmax = g[f"it{it}_grad_absmax"][0]
        assert max(np.abs(gx - g[f"it{it}_gx_sub8"]).max(), np.abs(gy - g[f"it{it}_gy_sub8"]).max()) <= 1e-6 * gmax
        tx, ty = vsub("target_x"), vsub("target_y")
        disp = max(np.abs(g[f"it{it}_target_x_sub2"] - sx).max(), np.abs(g[f"it{it}_target_y_sub2"] - sy).max())
        d = max(np.abs(tx - g[f"it{it}_target_x_sub2"]).max(), np.abs(ty - g[f"it{it}_target_y_sub2"]).max())
        assert d <= 1e-6 * disp, (it, d, disp)
    cd.close()


def test_folded_mesh_end_to_end_vs_oracle(pcd, port, oracle_mod):
    """ADVICE r01 (K-RAST tie-break): the high-contrast image of tests/test_oracle_pin.py whose mesh FOLDS (the CPU
    restatement is bit-identical to the reference on it, folds included).  Under folds the rasteriser here picks the
    lowest triangle index among the triangles covering a sample, the reference the first hit in BVH order
    (src/bvh.cpp:197-247), so the fields may differ there; this bounds the end-to-end effect: same step sizes to
    1e-3 relative, vertices within 1e-3 of the max displacement after six iterations, heights within 5e-4 of range."""
    rng = np.random.RandomState(3)
    img = np.zeros((80, 80))
    img[20:60, 30:50] = 1.0
    img += 0.02 * rng.rand(80, 80)
    s, resized = oracle_mod.prepare_image(img, 20, 0.5, 1.5, 0.1)
    cd = pcd.from_setup(s)
    cd.initialize_solvers(resized)
    od = port.design(s, solver_mode=0)
    od.initialize_solvers(resized)
    folded = False
    for it in range(6):
        a, b = cd.perform_transport_iteration(), od.transport_iteration()
        assert abs(a - b) <= 1e-3 * b, (it, a, b)
        # signed area of the deformed triangles turns negative where the mesh folds
        tx, ty = od.get("target_x").reshape(s.mesh_ny, s.mesh_nx), od.get("target_y").reshape(s.mesh_ny, s.mesh_nx)
        ax, ay = tx[:-1, 1:] - tx[:-1, :-1], ty[:-1, 1:] - ty[:-1, :-1]
        bx, by = tx[1:, :-1] - tx[:-1, :-1], ty[1:, :-1] - ty[:-1, :-1]
        folded = folded or bool(((ax * by - ay * bx) <= 0).any())
    disp = max(np.abs(od.get("target_x") - od.get("source_x")).max(), np.abs(od.get("target_y") - od.get("source_y")).max())
    d = max(np.abs(cd.get("target_x") - od.get("target_x")).max(), np.abs(cd.get("target_y") - od.get("target_y")).max())
    assert d <= 1e-3 * disp, (d, disp)
    for hi in range(3):
        cd.perform_height_map_iteration(hi)
        od.height_iteration(hi)
    z, zr = cd.get("source_z"), od.get("source_z")
    assert np.abs(z - zr).max() <= 5e-4 * (zr.max() - zr.min())
    print(f"folded={folded} vertex diff {d / disp:.3e} of max displacement, z diff {np.abs(z - zr).max() / (zr.max() - zr.min()):.3e} of range")
    cd.close()
    od.close()


def test_height_tolerance_knob(pcd, oracle_mod, golden):
    """pcd_set_tolerances: the height solves stop at 1e-9 by default and at the reference's 1e-8 on request.  On C1's
    first height problem red-black sweeps meet 1e-8 after 1 075 sweeps and 1e-9 after 1 572
    (tests/test_oracle_golden.py::test_height_truncation_evidence); the design on the device must reproduce both."""
    O = oracle_mod
    gray = O.rgba_to_gray(golden("images")["siggraph"])
    s, resized = O.prepare_image(gray, 100, O.f32(0.5), O.f32(1.5), O.f32(0.1))
    got = {}
    for tol in (0.0, 1e-8):
        cd = pcd.from_setup(s)
        cd.initialize_solvers(resized)
        if tol:
            cd.set_tolerances(0.0, tol)
        cd.run_transport(50, O.f32(0.01))
        cd.perform_height_map_iteration(0)
        got[tol] = cd.last_solve_info()["converged_at"]
        cd.close()
    assert abs(got[0.0] - 1572) <= 3 and abs(got[1e-8] - 1075) <= 3, got


def test_c5_stages_match_reference(pcd, golden):
    """BASELINE.json configs[4] (synthetic 8192x8192 density, mesh 2048x2048): the non-solver stages of the first
    transport iteration at FULL size against the reference itself (tests/golden/c5_stages.npz: oracle/_ref with its
    solver switched off by nthreads = 0) -- target areas over 4 M vertices, dual-cell areas / errors, the rasteriser on
    67 M samples, mean removal.  The mesh is still the regular lattice, so everything is reassociation-level."""
    path = os.path.join(GOLD, "c5_stages.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated yet")
    import hashlib
    from poisson_caustic_design_b200 import synth
    g = golden("c5_stages")
    img = synth.synth_density(8192, 8192, 8192)
    assert hashlib.md5(img.tobytes()).hexdigest().encode() == g["image_md5"].tobytes()
    st = synth.Setup(2048, 8192, 8192, mesh_width=1.0, focal_l=1.5, thickness=0.2)
    cd = pcd.from_setup(st)
    cd.initialize_solvers(img)
    del img
    ny, nx = st.mesh_ny, st.mesh_nx
    ta = cd.get("target_areas")
    assert abs(ta.sum() - g["target_areas_sum"][0]) < 1e-11
    tas = np.ascontiguousarray(ta.reshape(ny, nx)[::16, ::16]).ravel()
    assert np.abs(tas - g["target_areas_sub16"]).max() <= 1e-12 * g["target_areas_max"][0] * 10
    cd.stage_errors()
    cd.stage_raster()
    cd.stage_subtract_average()
    err = np.ascontiguousarray(cd.get("errors").reshape(ny, nx)[::16, ::16]).ravel()
    assert np.abs(err - g["errors_sub16"]).max() <= 1e-10 * g["errors_absmax"][0]
    ras = cd.get("raster")
    assert np.abs(ras[::64, ::64] - g["raster_sub64"]).max() <= 1e-9 * g["raster_absmax"][0]
    assert np.abs(ras[[0, 1, 4095, 8190, 8191], ::8] - g["raster_rows"]).max() <= 1e-9 * g["raster_absmax"][0]
    cd.close()
