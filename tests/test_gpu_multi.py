"""GPU: one Poisson problem on several GPUs of ONE process (include/pcd.h: pcd_multi_*, csrc/multi_gpu.cu; the C++
host's `--gpus N`) -- torch-free row slabs with the ghost-row exchange inside the persistent pass kernel.  With one
GPU the same code runs G slabs on that GPU (they share a stream and advance pass by pass); with >= 2 GPUs the slabs
sit on different devices and talk through peer access.  Bit-identical to the single-GPU solve in both cases."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
PKG = os.path.join(ROOT, "poisson_caustic_design_b200")
CLI = os.path.join(PKG, "caustic_design")


def single_gpu(pcd, D, phi0, sweeps, tol=0.0, path=None):
    h, w = D.shape
    s = pcd.Solver(w, h, 0, pcd.SOLVER_STREAMING if path is None else path)
    s.upload(D, phi0)
    info = s.run(sweeps, tol)
    out = s.download()
    s.close()
    return out, info


def multi_solve(pcd, D, phi0, devices, max_it, tol, check_every=None):
    import torch
    dev = torch.device("cuda", devices[0])
    Dd, pd = torch.from_numpy(D).to(dev), torch.from_numpy(phi0).to(dev)
    h, w = D.shape
    m = pcd.MultiGpuSolver(w, h, devices)
    if check_every:
        m.set_check_every(check_every)
    torch.cuda.synchronize()
    info = m.solve(Dd.data_ptr(), pd.data_ptr(), max_it, tol)
    out = pd.cpu().numpy()
    m.close()
    return out, info


def device_lists(pcd):
    lists = [[0, 0], [0, 0, 0]]
    n = pcd.device_count()
    if n >= 2:
        lists.append([0, 1])
    if n >= 4:
        lists.append([0, 1, 2, 3])
    return lists


@pytest.mark.parametrize("shape,sweeps", [((203, 150), 41), ((640, 1100), 37), ((900, 520), 24)])
def test_multi_gpu_solver_bit_identical(pcd, port, shape, sweeps):
    rng = np.random.RandomState(sum(shape) + sweeps)
    H, W = shape
    D = rng.standard_normal((H, W))
    D -= D.mean()
    phi0 = rng.standard_normal((H, W))
    want, winfo = single_gpu(pcd, D, phi0, sweeps)
    for devices in device_lists(pcd):
        got, info = multi_solve(pcd, D, phi0, devices, sweeps, 0.0, 16)
        assert info["sweeps"] == sweeps and np.array_equal(got, want), devices
        assert info["last_max_update"] == winfo["last_max_update"]
    assert np.array_equal(want, port.poisson_rb(D, phi0, sweeps, 0.0)[0])


def test_multi_gpu_stopping_rule_nan_and_thin_slabs(pcd, port, golden):
    rng = np.random.RandomState(3)
    H, W = 96, 128
    D = rng.standard_normal((H, W))
    D -= D.mean()
    z = np.zeros_like(D)
    _, _, conv_exact, _ = port.poisson_rb(D, z, 100000, 1e-7)
    for devices in device_lists(pcd):
        got, info = multi_solve(pcd, D, z, devices, 100000, 1e-7, 32)
        assert info["converged_at"] == conv_exact
        want = port.poisson_rb(D, z, 100000, 1e-7, extra_sweeps=info["sweeps"] - conv_exact)[0]
        assert np.array_equal(got, want), devices
    # NaN holes and slabs thinner than two ghost depths run on devices[0] alone -- same answer as the plain solver
    Dn = golden("solver")["nan_D"]
    got, info = multi_solve(pcd, Dn, np.zeros_like(Dn), [0, 0], 200, 0.0)
    assert np.array_equal(got, port.poisson_rb(Dn, np.zeros_like(Dn), 200, 0.0)[0], equal_nan=True)
    Dt = rng.standard_normal((30, 64))
    Dt -= Dt.mean()
    got, info = multi_solve(pcd, Dt, np.zeros_like(Dt), [0, 0, 0, 0], 50, 0.0)
    assert np.array_equal(got, port.poisson_rb(Dt, np.zeros_like(Dt), 50, 0.0)[0])


def test_design_on_several_slabs_matches_one_gpu(pcd):
    """pcd_multi_attach: the context's transport and height iterations with their Poisson solves on row slabs."""
    from poisson_caustic_design_b200 import synth
    res_w, aspect = 320, 4.0      # domain 1280 x 320, wavefront path on one GPU (forced: it would run on chip, transposed): same sweep schedule on slabs
    W = 4 * res_w
    H = int(W / aspect)
    image = synth.synth_density(W, H, 7)

    def run(devices):
        cd = pcd.from_setup(synth.Setup(res_w, W, H), 0, solver_path=pcd.SOLVER_TILED)
        cd.initialize_solvers(image)
        m = None
        if devices:
            m = pcd.MultiGpuSolver(W, H, devices)
            m.attach(cd)
        steps = [cd.perform_transport_iteration() for _ in range(2)]
        info = cd.last_solve_info()
        cd.perform_height_map_iteration(0)
        out = {k: cd.get(k) for k in ("phi", "h", "target_x", "target_y", "source_z")}
        if m is not None:
            m.close()       # before the context
        cd.close()
        return steps, info, out

    steps_ref, info_ref, ref = run(None)
    for devices in device_lists(pcd):
        steps, info, got = run(devices)
        assert steps == steps_ref, devices
        assert info["sweeps"] == info_ref["sweeps"] and info["converged_at"] == info_ref["converged_at"]
        for k in ref:
            assert np.array_equal(got[k], ref[k]), (devices, k)


def test_cli_gpus_flag(pcd, golden, tmp_path):
    """`caustic_design --gpus 2` (skipped below two GPUs): same step sizes and the same OBJ as one GPU."""
    if pcd.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import re
    from PIL import Image
    Image.fromarray(golden("images")["hello"], "RGB").save(tmp_path / "hello.png")
    outs = []
    for gpus in (1, 2):
        out_dir = str(tmp_path / f"g{gpus}") + "/"
        os.makedirs(out_dir)
        # one GPU on the wavefront path (the slabs' sweep schedule; 1280 x 640 would otherwise run on chip, transposed)
        cmd = [CLI, f"--input_png={tmp_path}/hello.png", "--res_w=320", "--mesh_width=0.5", f"--output={out_dir}", f"--gpus={gpus}",
               "--solver_path=tiled"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(([float(m) for m in re.findall(r"Transport step size = ([0-9.]+)", r.stdout)],
                     open(os.path.join(out_dir, "output.obj"), "rb").read()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
