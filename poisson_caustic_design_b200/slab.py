"""Row-slab Poisson solve across GPUs: one process per GPU, torch.distributed for the plumbing.

The W x H grid is cut into `world` contiguous row slabs.  Per red-black colour phase every rank updates its
rows (CUDA: ``pcd_slab_sweep_colour``), then swaps one boundary row with each neighbour (NCCL send/recv over
NVLink; gloo in the CPU tests); every ``check_every`` sweeps the per-sweep maxima are all-reduced (MAX) and
every rank takes the same stop decision.  Updates of one colour are order-independent and max is exact, so
the result is bit-identical to the single-GPU solve run for the same number of sweeps.

``engine`` abstracts the local slab (CudaSlabEngine below; the CPU tests plug a numpy engine in to exercise
this host logic without a GPU).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _check, _p, lib


def partition(H: int, world: int, rank: int):
    """Rows [row0, row0+rows) owned by `rank` (difference between ranks <= 1 row)."""
    row0 = rank * H // world
    return row0, (rank + 1) * H // world - row0


def with_ghosts(a: np.ndarray, row0: int, rows: int) -> np.ndarray:
    """Rows row0-1 .. row0+rows of a global [H, W] array; ghost rows outside the grid are zero."""
    H, W = a.shape
    out = np.zeros((rows + 2, W), dtype=np.float64)
    lo, hi = max(row0 - 1, 0), min(row0 + rows + 1, H)
    out[lo - (row0 - 1): hi - (row0 - 1)] = a[lo:hi]
    return out


class _DevView:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class CudaSlabEngine:
    """Local slab on one GPU (C ABI: pcd_slab_*), running on torch's current stream."""

    def __init__(self, W: int, H: int, row0: int, rows: int, device: int):
        import torch
        self.W, self.H, self.row0, self.rows, self.device = W, H, row0, rows, device
        torch.cuda.set_device(device)
        self._h = C.c_void_p()
        stream = torch.cuda.current_stream(device).cuda_stream
        _check(lib().pcd_slab_create(W, H, row0, rows, device, C.c_void_p(stream), C.byref(self._h)))
        phi, D, mx = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().pcd_slab_device_ptrs(self._h, C.byref(phi), C.byref(D), C.byref(mx)))
        dev = torch.device("cuda", device)
        self.phi = torch.as_tensor(_DevView(phi.value, (rows + 2, W), "<f8"), device=dev)
        self._max = torch.as_tensor(_DevView(mx.value, (4096,), "<f8"), device=dev)  # bit patterns of doubles >= 0

    def close(self):
        if self._h:
            lib().pcd_slab_destroy(self._h)
            self._h = C.c_void_p()

    def upload(self, D_g: np.ndarray, phi_g: np.ndarray):
        _check(lib().pcd_slab_upload(self._h, _p(np.ascontiguousarray(D_g)), _p(np.ascontiguousarray(phi_g))))

    def sweep_colour(self, colour: int, slot: int):
        _check(lib().pcd_slab_sweep_colour(self._h, colour, slot))

    def clear_max(self, n: int):
        _check(lib().pcd_slab_clear_max(self._h, n))

    def max_tensor(self, n: int):
        return self._max[:n]

    def download(self) -> np.ndarray:
        out = np.empty((self.rows, self.W), dtype=np.float64)
        _check(lib().pcd_slab_download(self._h, _p(out)))
        return out


def _exchange(engine, dist, rank: int, world: int):
    """Swap boundary rows with the neighbouring ranks: owned row 1 -> upper neighbour's lower ghost,
    owned row `rows` -> lower neighbour's upper ghost."""
    ops = []
    phi, rows = engine.phi, engine.rows
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, phi[1], rank - 1))
        ops.append(dist.P2POp(dist.irecv, phi[0], rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, phi[rows], rank + 1))
        ops.append(dist.P2POp(dist.irecv, phi[rows + 1], rank + 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def solve(engine, dist, rank: int, world: int, max_iterations: int, tol: float, check_every: int = 64):
    """Distributed red-black SOR.  Returns {"sweeps", "converged_at", "last_max_update"} (same on all ranks)."""
    import torch
    done, conv, last = 0, 0, 0.0
    check_every = max(1, min(check_every, 4096))
    while done < max_iterations and not conv:
        k = min(check_every, max_iterations - done)
        engine.clear_max(k)
        for j in range(k):
            for colour in (0, 1):
                engine.sweep_colour(colour, j)
                if world > 1:
                    _exchange(engine, dist, rank, world)
        m = engine.max_tensor(k).clone()
        if world > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
        m_host = m.cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m)
        below = np.nonzero(m_host < tol)[0]
        if below.size:
            conv = done + int(below[0]) + 1
            last = float(m_host[below[0]])
        else:
            last = float(m_host[-1])
        done += k
    return {"sweeps": done, "converged_at": conv, "last_max_update": last}


def solve_local(engines, max_iterations: int, tol: float, check_every: int = 64):
    """Single-process emulation of G slabs (all engines in this process, e.g. G slabs on ONE GPU): same phase
    structure as `solve`, halo rows copied directly between the engines."""
    import torch
    G = len(engines)
    done, conv, last = 0, 0, 0.0
    check_every = max(1, min(check_every, 4096))
    while done < max_iterations and not conv:
        k = min(check_every, max_iterations - done)
        for e in engines:
            e.clear_max(k)
        for j in range(k):
            for colour in (0, 1):
                for e in engines:
                    e.sweep_colour(colour, j)
                for g in range(G - 1):
                    up, dn = engines[g], engines[g + 1]
                    dn.phi[0].copy_(up.phi[up.rows])
                    up.phi[up.rows + 1].copy_(dn.phi[1])
        ms = [e.max_tensor(k) for e in engines]
        m = ms[0].clone()
        for other in ms[1:]:
            m = torch.maximum(m, other.to(m.device)) if isinstance(m, torch.Tensor) else np.maximum(m, other)
        m_host = m.cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m)
        below = np.nonzero(m_host < tol)[0]
        if below.size:
            conv = done + int(below[0]) + 1
            last = float(m_host[below[0]])
        else:
            last = float(m_host[-1])
        done += k
    return {"sweeps": done, "converged_at": conv, "last_max_update": last}
