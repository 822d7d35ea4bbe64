"""CPU (gloo, world_size 2 and 3): the multi-GPU host logic of poisson_caustic_design_b200.slab -- row
partition, one-row halo exchange per colour phase, all-reduced stopping rule -- with a numpy engine standing in
for the CUDA slab.  The G-slab result must be bit-identical to the single-domain red-black oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, D, phi0, max_it, tol, check_every, out_dir, mode):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from poisson_caustic_design_b200 import slab
    from slab_numpy_engine import NumpySlabEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W = D.shape
    row0, rows = slab.partition(H, world, rank)
    eng = NumpySlabEngine(W, H, row0, rows)
    eng.upload(slab.with_ghosts(D, row0, rows, eng.GH), slab.with_ghosts(phi0, row0, rows, eng.GH))
    info = slab.solve(eng, dist, rank, world, max_it, tol, check_every, mode)
    assert info["mode"] == ("colour" if (mode == "colour" or np.isnan(D).any()) else "wavefront")
    np.save(os.path.join(out_dir, f"phi_{rank}.npy"), eng.download())
    np.save(os.path.join(out_dir, f"info_{rank}.npy"), np.array([info["sweeps"], info["converged_at"], info["last_max_update"]]))
    dist.destroy_process_group()


def run_world(world, D, phi0, max_it, tol, check_every, tmp_path, port, mode="auto"):
    mp.spawn(_worker, args=(world, port, D, phi0, max_it, tol, check_every, str(tmp_path), mode), nprocs=world, join=True)
    phi = np.concatenate([np.load(tmp_path / f"phi_{r}.npy") for r in range(world)], axis=0)
    infos = [np.load(tmp_path / f"info_{r}.npy") for r in range(world)]
    for i in infos[1:]:
        assert np.array_equal(i, infos[0])           # every rank took the same decision
    return phi, infos[0]


def test_partition_covers_the_grid():
    from poisson_caustic_design_b200 import slab
    for H in (1, 7, 64, 1000, 8192):
        for world in (1, 2, 3, 8):
            parts = [slab.partition(H, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(p[1] for p in parts) == H
            for a, b in zip(parts, parts[1:]):
                assert a[0] + a[1] == b[0]
            assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1
    g = slab.with_ghosts(np.arange(12.0).reshape(4, 3), 0, 2)
    assert g.shape == (4, 3) and not g[0].any() and np.array_equal(g[1:], np.arange(9.0).reshape(3, 3))
    g = slab.with_ghosts(np.arange(12.0).reshape(4, 3), 2, 2, 5)
    assert g.shape == (12, 3) and not g[:3].any() and not g[7:].any() and np.array_equal(g[3:7], np.arange(12.0).reshape(4, 3))


@pytest.mark.parametrize("mode", ["wavefront", "colour"])
@pytest.mark.parametrize("world", [2, 3])
def test_gloo_slabs_without_holes(port, world, mode, tmp_path):
    """Both ways to advance (TS sweeps per pass + GH-row exchange; one colour phase + one-row exchange)."""
    rng = np.random.RandomState(10 * world)
    H, W = 41, 30                                     # uneven partition, slabs >= GH rows
    D = rng.standard_normal((H, W))
    D -= D.mean()
    phi0 = rng.standard_normal((H, W))
    phi, info = run_world(world, D, phi0, 23, 0.0, 8, tmp_path, 29700 + 10 * world + (mode == "colour"), mode)
    want, n, conv, last = port.poisson_rb(D, phi0, 23, 0.0)     # 23: the last pass is a partial one
    assert int(info[0]) == 23 and int(info[1]) == 0
    assert np.array_equal(phi, want)
    assert info[2] == last


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_slabs_bit_identical_to_single_domain(port, world, tmp_path):
    rng = np.random.RandomState(world)
    H, W = 37, 26                                     # uneven partition
    D = rng.standard_normal((H, W))
    D[10:13, 5:9] = np.nan                            # a hole straddling slab interiors
    D[18, 20] = np.nan
    D[np.isfinite(D)] -= D[np.isfinite(D)].mean()
    phi0 = rng.standard_normal((H, W))
    phi, info = run_world(world, D, phi0, 25, 0.0, 8, tmp_path, 29600 + world)
    want, n, conv, last = port.poisson_rb(D, phi0, 25, 0.0)
    assert int(info[0]) == 25 and int(info[1]) == 0
    assert np.array_equal(phi, want, equal_nan=True)
    assert info[2] == last


def test_gloo_slabs_stopping_rule(port, tmp_path):
    rng = np.random.RandomState(9)
    H, W = 24, 32
    D = rng.standard_normal((H, W))
    D -= D.mean()
    z = np.zeros_like(D)
    phi, info = run_world(2, D, z, 100000, 1e-7, 16, tmp_path, 29650)
    _, n_exact, conv_exact, _ = port.poisson_rb(D, z, 100000, 1e-7)
    assert int(info[1]) == conv_exact                 # same sweep satisfies max|delta| < tol
    assert int(info[0]) % 16 == 0 and 16 <= int(info[0]) - conv_exact < 32    # the stop is acted on one block late
    want = port.poisson_rb(D, z, 100000, 1e-7, extra_sweeps=int(info[0]) - conv_exact)[0]
    assert np.array_equal(phi, want)
    assert info[2] < 1e-7


def _peer_worker(rank, world, port, D, phi0, max_it, tol, check_every, out_dir, failing_rank):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from poisson_caustic_design_b200 import slab
    from slab_numpy_engine import NumpyPeerEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W = D.shape
    row0, rows = slab.partition(H, world, rank)
    eng = NumpyPeerEngine(W, H, row0, rows, dist, rank, world, fail_connect=(rank == failing_rank))
    eng.upload(slab.with_ghosts(D, row0, rows, eng.GH), slab.with_ghosts(phi0, row0, rows, eng.GH))
    info = slab.solve(eng, dist, rank, world, max_it, tol, check_every)
    if failing_rank < 0:
        assert info["mode"] == "peer" and eng.runs == -(-info["sweeps"] // check_every)
        assert set(eng.neighbours) == {s for s, nb in ((0, rank - 1), (1, rank + 1)) if 0 <= nb < world}
    else:
        assert info["mode"] == "wavefront" and eng.runs == 0        # every rank fell back, also the ones that attached fine
    np.save(os.path.join(out_dir, f"phi_{rank}.npy"), eng.download())
    np.save(os.path.join(out_dir, f"info_{rank}.npy"), np.array([info["sweeps"], info["converged_at"], info["last_max_update"]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("failing_rank", [-1, 1])
def test_gloo_peer_mode_host_logic(port, failing_rank, tmp_path):
    """The fused-exchange path of slab.solve (handles all-gathered, neighbours attached, one peer_run per block of
    sweeps, the only collective = the all-reduce of the maxima), and the agreement to fall back when ONE rank
    cannot attach its neighbour."""
    world = 3
    rng = np.random.RandomState(77)
    H, W = 47, 22                                      # slabs of 15/16 rows >= 2*GH
    D = rng.standard_normal((H, W))
    D -= D.mean()
    phi0 = rng.standard_normal((H, W))
    mp.spawn(_peer_worker, args=(world, 29680 + failing_rank, D, phi0, 21, 0.0, 8, str(tmp_path), failing_rank), nprocs=world, join=True)
    phi = np.concatenate([np.load(tmp_path / f"phi_{r}.npy") for r in range(world)], axis=0)
    infos = [np.load(tmp_path / f"info_{r}.npy") for r in range(world)]
    for i in infos[1:]:
        assert np.array_equal(i, infos[0])
    want, n, conv, last = port.poisson_rb(D, phi0, 21, 0.0)
    assert int(infos[0][0]) == 21 and np.array_equal(phi, want) and infos[0][2] == last


def _stall_worker(rank, world, port, D, phi0, out_dir, stalled_rank):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from poisson_caustic_design_b200 import slab
    from slab_numpy_engine import NumpyPeerEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W = D.shape
    row0, rows = slab.partition(H, world, rank)
    eng = NumpyPeerEngine(W, H, row0, rows, dist, rank, world, stall_after_runs=2 if rank == stalled_rank else 0)
    eng.upload(slab.with_ghosts(D, row0, rows, eng.GH), slab.with_ghosts(phi0, row0, rows, eng.GH))
    raised = 0
    try:
        slab.solve(eng, dist, rank, world, 100, 0.0, 8)
    except RuntimeError as e:
        raised = 1
        assert "gave up waiting" in str(e)
    # every rank left the solve after the SAME block (the one in which the stalled rank reported), so this
    # collective -- the hook's broadcasts in the product -- lines up again on all ranks
    t = torch.tensor([eng.runs], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, f"stall_{rank}.npy"), np.array([raised, eng.runs, int(t.item())]))
    dist.destroy_process_group()


def test_gloo_peer_mode_stalled_neighbour_stops_every_rank_together(tmp_path):
    """ADVICE r01 (slab.py timeout flag): the error word of a rank whose pass waited in vain rides in the block's
    all-reduce, so all ranks raise in the same block instead of one rank raising while the others block in a
    later collective."""
    world = 3
    rng = np.random.RandomState(5)
    H, W = 47, 22
    D = rng.standard_normal((H, W))
    D -= D.mean()
    mp.spawn(_stall_worker, args=(world, 29691, D, np.zeros_like(D), str(tmp_path), 1), nprocs=world, join=True)
    res = [np.load(tmp_path / f"stall_{r}.npy") for r in range(world)]
    for r in res:
        assert r[0] == 1 and r[1] == 3 and r[2] == 3      # all raised together: block 3 was already queued when block 2's word arrived
