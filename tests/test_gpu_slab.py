"""GPU: the row-slab solver (csrc/sor_slab.cu + slab.py).  On one GPU: G slabs emulated in one process must be
bit-identical to the single-GPU solve ("multi-GPU without 8 GPUs").  With >= 2 GPUs: a real
2-process NCCL run (torchrun) against the same answer."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def single_gpu(pcd, D, phi0, sweeps):
    h, w = D.shape
    s = pcd.Solver(w, h, 0, pcd.SOLVER_STREAMING)
    s.upload(D, phi0)
    info = s.run(sweeps, 0.0)
    out = s.download()
    s.close()
    return out, info


@pytest.mark.parametrize("holes", [False, True])
@pytest.mark.parametrize("G", [1, 2, 5])
def test_slabs_on_one_gpu_bit_identical(pcd, port, G, holes):
    from poisson_caustic_design_b200 import slab
    rng = np.random.RandomState(G)
    H, W = 203, 150
    D = rng.standard_normal((H, W))
    if holes:
        D[50:60, 70:90] = np.nan                      # NaN holes -> colour mode (masked kernels)
    D[np.isfinite(D)] -= D[np.isfinite(D)].mean()
    phi0 = rng.standard_normal((H, W))
    engines = []
    for g in range(G):
        row0, rows = slab.partition(H, G, g)
        e = slab.CudaSlabEngine(W, H, row0, rows, 0)
        e.upload(slab.with_ghosts(D, row0, rows, e.GH), slab.with_ghosts(phi0, row0, rows, e.GH))
        engines.append(e)
    info = slab.solve_local(engines, 41, 0.0, 16)
    got = np.concatenate([e.download() for e in engines], axis=0)
    for e in engines:
        e.close()
    want, winfo = single_gpu(pcd, D, phi0, 41)
    assert info["sweeps"] == 41 and info["mode"] == ("colour" if holes else "wavefront")
    assert np.array_equal(got, want, equal_nan=True)
    assert np.array_equal(got, port.poisson_rb(D, phi0, 41, 0.0)[0], equal_nan=True)
    assert info["last_max_update"] == winfo["last_max_update"]


def test_slab_stopping_rule_matches_single_gpu(pcd, port):
    from poisson_caustic_design_b200 import slab
    rng = np.random.RandomState(3)
    H, W = 96, 128
    D = rng.standard_normal((H, W))
    D -= D.mean()
    z = np.zeros_like(D)
    engines = []
    for g in range(3):
        row0, rows = slab.partition(H, 3, g)
        e = slab.CudaSlabEngine(W, H, row0, rows, 0)
        e.upload(slab.with_ghosts(D, row0, rows, e.GH), slab.with_ghosts(z, row0, rows, e.GH))
        engines.append(e)
    info = slab.solve_local(engines, 100000, 1e-7, 32)
    got = np.concatenate([e.download() for e in engines], axis=0)
    for e in engines:
        e.close()
    _, _, conv_exact, _ = port.poisson_rb(D, z, 100000, 1e-7)
    assert info["converged_at"] == conv_exact
    want = port.poisson_rb(D, z, 100000, 1e-7, extra_sweeps=info["sweeps"] - conv_exact)[0]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape,G,sweeps", [((203, 150), 2, 41), ((203, 150), 5, 41), ((640, 1100), 2, 37), ((900, 520), 3, 24)])
def test_fused_exchange_on_one_gpu_bit_identical(pcd, port, shape, G, sweeps):
    """The pass kernels that push ghost rows into the neighbouring slab and wait on its flag (pcd_slab_peer_*), with
    all G slabs on one device: same field, same per-sweep maxima as the single-GPU solve."""
    from poisson_caustic_design_b200 import slab
    rng = np.random.RandomState(G + sweeps)
    H, W = shape
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
    # a second solve on the same slabs (sequence numbers keep counting)
    for g, e in enumerate(engines):
        row0, rows = slab.partition(H, G, g)
        e.upload(None, slab.with_ghosts(phi0, row0, rows, e.GH))
    info2 = slab.solve_local_peer(engines, 5, 0.0, 16)
    got2 = np.concatenate([e.download() for e in engines], axis=0)
    for e in engines:
        e.close()
    want, winfo = single_gpu(pcd, D, phi0, sweeps)
    assert info["sweeps"] == sweeps and info["mode"] == "peer"
    assert np.array_equal(got, want)
    assert info["last_max_update"] == winfo["last_max_update"]
    assert np.array_equal(got2, single_gpu(pcd, D, phi0, 5)[0]) and info2["sweeps"] == 5


def test_fused_exchange_production_geometry(pcd):
    """8192-wide slabs of 1024 rows -- the geometry of the 8192^2 problem on eight GPUs, where the launcher shortens the
    first chunk of a slab with an upper neighbour (it also feeds that neighbour's ghost rows) and gives the bottom rows a
    32-row chunk of their own: two such slabs on one GPU against the single-GPU per-colour kernels."""
    from poisson_caustic_design_b200 import slab
    rng = np.random.RandomState(8)
    H, W, G, sweeps = 2048, 8192, 2, 9
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
    want, winfo = single_gpu(pcd, D, phi0, sweeps)
    assert info["sweeps"] == sweeps and np.array_equal(got, want)
    assert info["last_max_update"] == winfo["last_max_update"]


@pytest.mark.parametrize("mode", ["peer", "wavefront", "colour"])
def test_two_gpu_nccl_run(pcd, tmp_path, mode):
    if pcd.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(ROOT, "tools", "slab_run.py")
    out = tmp_path / "out.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", script, "--W", "1100", "--H", "640", "--sweeps", "61", "--out", str(out), "--mode", mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f'"mode": "{mode}"' in r.stdout
    z = np.load(out)
    want, _ = single_gpu(pcd, z["D"], z["phi0"], 61)
    assert np.array_equal(z["phi"], want)


def test_two_gpu_soak(pcd, tmp_path):
    """10^5 sweeps of the fused peer protocol on two real GPUs (3 125 persistent launches of 16 passes each, per-strip
    NVLink flags, one-block-late stopping rule with the maxima copied out on a side stream) against one GPU."""
    if pcd.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(ROOT, "tools", "slab_run.py")
    out = tmp_path / "soak.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", script, "--W", "1536", "--H", "1100", "--sweeps", "100000", "--check_every", "32", "--out", str(out),
           "--mode", "peer"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    z = np.load(out)
    want, _ = single_gpu(pcd, z["D"], z["phi0"], 100000)
    assert np.array_equal(z["phi"], want)


def _design(pcd, res_w, aspect, seed, device=0):
    from poisson_caustic_design_b200 import synth
    W = 4 * res_w
    H = int(W / aspect)
    image = synth.synth_density(W, H, seed)
    # the wavefront path on purpose (its sweep schedule is the slab driver's): 1280 x 320 would otherwise run on chip, transposed
    cd = pcd.from_setup(synth.Setup(res_w, W, H), device, solver_path=pcd.SOLVER_TILED)
    cd.initialize_solvers(image)
    return cd


def test_solve_hook_on_one_gpu_matches_the_builtin_solver(pcd):
    """The slab driver installed as the context's Poisson solver (pcd_set_solve_hook), world size 1: same sweep
    schedule as the built-in large-grid solver, so steps and fields are bit-identical."""
    from poisson_caustic_design_b200 import slab
    ref = _design(pcd, 320, 4.0, 7)          # domain 1280 x 320 on the wavefront path
    steps_ref = [ref.perform_transport_iteration() for _ in range(2)]
    info_ref = ref.last_solve_info()
    ref.perform_height_map_iteration(0)
    got = _design(pcd, 320, 4.0, 7)
    hook = slab.SlabSolveHook(got, None, 0, 1, 0)
    steps = [got.perform_transport_iteration() for _ in range(2)]
    info = got.last_solve_info()
    got.perform_height_map_iteration(0)
    assert hook.error is None and len(hook.solves) == 3
    assert info_ref["path"] == "tiled" and info["path"] == "tiled"
    assert info["sweeps"] == info_ref["sweeps"] and info["converged_at"] == info_ref["converged_at"]
    assert steps == steps_ref
    for name in ("phi", "h", "target_x", "target_y", "source_z"):
        assert np.array_equal(got.get(name), ref.get(name)), name
    hook.close()
    # the hook is gone: the context solves on its own again
    assert got.perform_transport_iteration() == ref.perform_transport_iteration()
    got.close(); ref.close()


def test_two_gpu_design_matches_one_gpu(pcd, tmp_path):
    if pcd.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(ROOT, "tools", "dist_design_run.py")
    out = tmp_path / "design.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", script, "--res_w", "320", "--aspect", "4", "--iters", "2", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    z = np.load(out)
    ref = _design(pcd, 320, 4.0, 7)
    steps_ref = [ref.perform_transport_iteration() for _ in range(2)]
    ref.perform_height_map_iteration(0)
    assert list(z["steps"]) == steps_ref
    for name, key in (("phi", "phi"), ("h", "h"), ("target_x", "tx"), ("target_y", "ty"), ("source_z", "sz")):
        assert np.array_equal(z[key], ref.get(name)), name
    ref.close()
