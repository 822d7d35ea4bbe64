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
