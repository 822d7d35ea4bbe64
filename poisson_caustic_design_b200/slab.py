"""Row-slab Poisson solve across GPUs: one process per GPU, torch.distributed for the plumbing.

The W x H grid is cut into `world` contiguous row slabs, each held with GH ghost rows above and below.
Three ways to advance, all bit-identical to the single-GPU solve run for the same number of sweeps (updates of
one colour are order-independent, max is exact):

* peer mode (default on GPUs when the neighbours' memory can be attached through CUDA IPC): the ghost-row exchange
  is fused into the wavefront pass kernel -- plain stores into the neighbour's field over NVLink plus a sequence
  flag the neighbour's next pass waits on (csrc/sor_tiled.cu, PEER variant); a block of sweeps is back-to-back
  kernel launches issued from C, no host-side collective on the data path;
* wavefront mode: ``engine.pass_`` runs up to TS full red-black sweeps over the slab in one kernel (temporal
  blocking); afterwards every rank sends its GH boundary rows to each neighbour (NCCL send/recv over NVLink; gloo
  in the CPU tests) -- ONE exchange per TS sweeps;
* colour mode (D has NaN holes, or slabs thinner than GH rows): one kernel per colour phase, one ghost row
  exchanged per phase.

Every ``check_every`` sweeps the per-sweep maxima are all-reduced (MAX) and every rank takes the same stop
decision.  ``engine`` abstracts the local slab (CudaSlabEngine below; the CPU tests plug a numpy engine in to
exercise this host logic without a GPU).  ``SlabSolveHook`` installs the distributed solve as the Poisson solver
of a whole caustic design (include/pcd.h: pcd_set_solve_hook).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _check, _p, lib


def partition(H: int, world: int, rank: int):
    """Rows [row0, row0+rows) owned by `rank` (difference between ranks <= 1 row)."""
    row0 = rank * H // world
    return row0, (rank + 1) * H // world - row0


def with_ghosts(a: np.ndarray, row0: int, rows: int, gh: int = 1) -> np.ndarray:
    """Rows row0-gh .. row0+rows+gh-1 of a global [H, W] array; ghost rows outside the grid are zero."""
    H, W = a.shape
    out = np.zeros((rows + 2 * gh, W), dtype=np.float64)
    lo, hi = max(row0 - gh, 0), min(row0 + rows + gh, H)
    out[lo - (row0 - gh): hi - (row0 - gh)] = a[lo:hi]
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
        self.GH = lib().pcd_slab_ghost_rows()
        self.TS = lib().pcd_slab_sweeps_per_pass()
        torch.cuda.set_device(device)
        self._h = C.c_void_p()
        stream = torch.cuda.current_stream(device).cuda_stream
        _check(lib().pcd_slab_create(W, H, row0, rows, device, C.c_void_p(stream), C.byref(self._h)))
        p0, p1, mx = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().pcd_slab_device_ptrs(self._h, C.byref(p0), C.byref(p1), C.byref(mx)))
        dev = torch.device("cuda", device)
        shape = (rows + 2 * self.GH, W)
        self._phi = [torch.as_tensor(_DevView(p0.value, shape, "<f8"), device=dev),
                     torch.as_tensor(_DevView(p1.value, shape, "<f8"), device=dev)]
        self._max = torch.as_tensor(_DevView(mx.value, (4096,), "<f8"), device=dev)  # bit patterns of doubles >= 0

    def close(self):
        if self._h:
            lib().pcd_slab_destroy(self._h)
            self._h = C.c_void_p()

    @property
    def phi(self):
        """[rows + 2*GH, W] view of the buffer that currently holds the field."""
        return self._phi[lib().pcd_slab_current(self._h)]

    @property
    def has_nan(self) -> bool:
        return bool(lib().pcd_slab_has_nan(self._h))

    def upload(self, D_g: np.ndarray, phi_g: np.ndarray):
        _check(lib().pcd_slab_upload(self._h, None if D_g is None else _p(np.ascontiguousarray(D_g)),
                                     None if phi_g is None else _p(np.ascontiguousarray(phi_g))))

    def sweep_colour(self, colour: int, slot: int):
        _check(lib().pcd_slab_sweep_colour(self._h, colour, slot))

    def pass_(self, nsweeps: int, slot: int):
        _check(lib().pcd_slab_pass(self._h, nsweeps, slot))

    def pass_part(self, nsweeps: int, slot: int, row_begin: int, row_count: int, stream=None):
        """The pass over owned rows [row_begin, row_begin+row_count) only (global indices); buffers are not switched."""
        cs = C.c_void_p(stream.cuda_stream) if stream is not None else None
        _check(lib().pcd_slab_pass_part(self._h, nsweeps, slot, row_begin, row_count, cs))

    def flip(self):
        _check(lib().pcd_slab_flip(self._h))

    def bands(self, band: int = 64):
        """(edge bands, interior) row ranges of a pass: the bands produce everything the neighbours need."""
        r0, n = self.row0, self.rows
        b = min(band, n // 2)
        if b < self.GH or n - 2 * b < 1:
            return [(r0, n)], None
        return [(r0, b), (r0 + n - b, b)], (r0 + b, n - 2 * b)

    # ---- ghost-row exchange fused into the pass kernel (peer memory over NVLink) ----
    def peer_export(self) -> np.ndarray:
        h = np.zeros(lib().pcd_slab_peer_handle_bytes(), dtype=np.uint8)
        _check(lib().pcd_slab_peer_export(self._h, h.ctypes.data_as(C.c_void_p)))
        return h

    def peer_connect_ipc(self, side: int, handles: np.ndarray, peer_row0: int, peer_rows: int):
        h = np.ascontiguousarray(handles, dtype=np.uint8)
        _check(lib().pcd_slab_peer_connect_ipc(self._h, side, h.ctypes.data_as(C.c_void_p), peer_row0, peer_rows))

    def peer_connect_local(self, side: int, other: "CudaSlabEngine"):
        _check(lib().pcd_slab_peer_connect_local(self._h, side, other._h))

    def peer_run(self, nsweeps: int, slot: int):
        _check(lib().pcd_slab_peer_run(self._h, nsweeps, slot))

    def peer_timed_out(self) -> bool:
        t = C.c_int(0)
        _check(lib().pcd_slab_peer_status(self._h, C.byref(t)))
        return bool(t.value)

    def peer_error_into(self, slot):
        """The slab's error word (1.0 = a pass ran into its time limit waiting for a neighbour) into the one-element
        device tensor `slot`, asynchronously on the slab's stream."""
        _check(lib().pcd_slab_peer_error_to(self._h, C.c_void_p(slot.data_ptr())))

    def set_sm_reserve(self, n: int):
        _check(lib().pcd_slab_set_sm_reserve(self._h, n))

    def load_device(self, D_dev: int, phi_dev: int):
        """D and phi (ghost rows included) from full W x H device arrays on this GPU."""
        _check(lib().pcd_slab_load_device(self._h, C.c_void_p(D_dev), C.c_void_p(phi_dev)))

    def store_device(self, phi_dev: int):
        _check(lib().pcd_slab_store_device(self._h, C.c_void_p(phi_dev)))

    def clear_max(self, n: int, first: int = 0):
        _check(lib().pcd_slab_clear_max_range(self._h, first, n))

    def max_tensor(self, n: int, first: int = 0):
        return self._max[first:first + n]

    def download(self) -> np.ndarray:
        out = np.empty((self.rows, self.W), dtype=np.float64)
        _check(lib().pcd_slab_download(self._h, _p(out)))
        return out


def _exchange(engine, dist, rank: int, world: int, depth: int):
    """Refresh `depth` ghost rows on each side: my top `depth` owned rows -> upper neighbour's lower ghosts,
    my bottom `depth` owned rows -> lower neighbour's upper ghosts."""
    ops = []
    phi, rows, GH = engine.phi, engine.rows, engine.GH
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, phi[GH:GH + depth], rank - 1))
        ops.append(dist.P2POp(dist.irecv, phi[GH - depth:GH], rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, phi[GH + rows - depth:GH + rows], rank + 1))
        ops.append(dist.P2POp(dist.irecv, phi[GH + rows:GH + rows + depth], rank + 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def _decide(m_host: np.ndarray, tol: float, done: int):
    below = np.nonzero(m_host < tol)[0]
    if below.size:
        return done + int(below[0]) + 1, float(m_host[below[0]])
    return 0, float(m_host[-1])


def use_wavefront(engine, dist, world: int, min_rows: int) -> bool:
    """Same answer on every rank: no NaN anywhere and every slab at least GH rows thick."""
    import torch
    ok = int((not engine.has_nan) and hasattr(engine, "pass_") and min_rows >= engine.GH)
    if world > 1:
        t = torch.tensor([ok], dtype=torch.int32, device=engine.phi.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = int(t.item())
    return bool(ok)


def connect_peers(engine, dist, rank: int, world: int) -> bool:
    """Attach the neighbouring ranks' slabs for the fused exchange (CUDA IPC handles travel through an all-gather).
    Same answer on every rank; False leaves the host-driven exchange in charge."""
    import torch
    if getattr(engine, "_peers_connected", False):
        return True
    if not hasattr(engine, "peer_run"):   # engine type is the same on every rank
        return False
    H = engine.H
    min_rows = min(partition(H, world, r)[1] for r in range(world))
    ok = int(hasattr(engine, "peer_run") and world > 1 and min_rows >= 2 * engine.GH)
    dev = engine.phi.device
    mine = torch.from_numpy(engine.peer_export()).to(dev) if ok else torch.zeros(128, dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if ok:
        try:
            for side, nb in ((0, rank - 1), (1, rank + 1)):
                if 0 <= nb < world:
                    r0, n = partition(H, world, nb)
                    engine.peer_connect_ipc(side, gathered[nb].cpu().numpy(), r0, n)
        except RuntimeError:
            ok = 0
    t = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    engine._peers_connected = bool(t.item())
    return engine._peers_connected


def solve(engine, dist, rank: int, world: int, max_iterations: int, tol: float, check_every: int = 64, mode: str = "auto"):
    """Distributed red-black SOR.  Returns {"sweeps", "converged_at", "last_max_update", "mode"} (same on all ranks).

    mode: "auto" (fused peer exchange when the neighbours' memory can be attached, else wavefront passes with a
    host-driven exchange, else colour phases), "peer", "wavefront", "colour"."""
    import torch
    H = engine.H
    min_rows = min(partition(H, world, r)[1] for r in range(world))
    wave = use_wavefront(engine, dist, world, min_rows) if mode in ("auto", "peer") else (mode == "wavefront")
    peer = False
    if wave and world > 1 and mode in ("auto", "peer"):
        peer = connect_peers(engine, dist, rank, world)
        if mode == "peer" and not peer:
            raise RuntimeError("slab.solve(mode='peer'): the neighbouring slabs could not be attached")
    if peer:
        return _solve_peer(engine, dist, world, max_iterations, tol, check_every)
    TS = engine.TS if wave else 1
    check_every = max(TS, min(check_every, 4096))
    check_every = (check_every + TS - 1) // TS * TS
    # exchange/compute overlap (CUDA engines, world > 1): the bands next to the slab edges run first, their ghost-row
    # exchange goes to a side stream while the interior rows of the same pass are still being updated
    overlap = wave and world > 1 and hasattr(engine, "pass_part") and engine.bands()[1] is not None
    if overlap:
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        ev_bands, ev_xchg = torch.cuda.Event(), torch.cuda.Event()
        edge_bands, interior = engine.bands()
        engine.set_sm_reserve(8)   # room for the NCCL send/recv kernels next to the interior pass (this slab only)
    done, conv, last = 0, 0, 0.0
    try:
        # same schedule as the fused path and the single-GPU large-grid solver: the stop is acted on ONE BLOCK LATE
        while done < max_iterations:
            was_conv = bool(conv)
            k = min(check_every, max_iterations - done)
            engine.clear_max(k)
            if wave and overlap:
                j = 0
                while j < k:
                    ns = min(TS, k - j)
                    for (rb, rc) in edge_bands:
                        engine.pass_part(ns, j, rb, rc)
                    ev_bands.record(main)
                    engine.pass_part(ns, j, interior[0], interior[1])
                    engine.flip()
                    with torch.cuda.stream(side):
                        side.wait_event(ev_bands)
                        _exchange(engine, dist, rank, world, engine.GH)   # on the buffer the pass just filled
                        ev_xchg.record(side)
                    main.wait_event(ev_xchg)
                    j += ns
            elif wave:
                j = 0
                while j < k:
                    ns = min(TS, k - j)
                    engine.pass_(ns, j)
                    if world > 1:
                        _exchange(engine, dist, rank, world, engine.GH)
                    j += ns
            else:
                for j in range(k):
                    for colour in (0, 1):
                        engine.sweep_colour(colour, j)
                        if world > 1:
                            _exchange(engine, dist, rank, world, 1)
            m = engine.max_tensor(k).clone()
            if world > 1:
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
            m_host = m.cpu().numpy() if isinstance(m, torch.Tensor) else np.asarray(m)
            if not conv:
                conv, last = _decide(m_host, tol, done)
            done += k
            if was_conv:
                break
    finally:
        if overlap:
            engine.set_sm_reserve(0)   # also when the loop raised: later solves on this slab get every SM back
    return {"sweeps": done, "converged_at": conv, "last_max_update": last, "mode": "wavefront" if wave else "colour"}


def _solve_peer(engine, dist, world: int, max_iterations: int, tol: float, check_every: int):
    """Fused path: a block of check_every sweeps is ONE persistent kernel launch (the ghost rows travel inside it); the
    only collective is the all-reduce of the block's per-sweep maxima.  The stopping rule is evaluated ONE BLOCK LATE:
    block b+1 is already queued when the maxima of block b reach the host (copied out on a side stream), so the GPUs
    never drain between blocks; a solve that meets the rule in block b therefore executes block b+1 as well -- exactly
    what the single-GPU large-grid solver does (run_tiled, csrc/sor_kernels.cu), so both end on the same bits."""
    import torch
    check_every = max(1, min(check_every, 2048))
    half = 2048                                    # the slot ring holds two blocks: even blocks low half, odd high
    launched, done, conv, last = 0, 0, 0, 0.0
    side = getattr(engine, "_side_stream", None)
    cuda = hasattr(engine, "_h")
    if cuda and side is None:
        side = engine._side_stream = torch.cuda.Stream()
        engine._stage = [None, None]
        engine._host = [torch.empty(half + 1, dtype=torch.float64, pin_memory=True) for _ in range(2)]
    pending = []                                   # blocks in flight: (k, first sweep, host getter)
    blk = 0

    def launch():
        nonlocal launched, blk
        k = min(check_every, max_iterations - launched)
        off = (blk & 1) * half
        engine.clear_max(k, off)
        engine.peer_run(k, off)
        # the block's maxima plus, in the last slot, this rank's error word (a pass that waited ~3 s in vain for a
        # neighbour): one all-reduce (MAX) tells EVERY rank about a stalled neighbour in the same block, so all ranks
        # stop together instead of launching further passes on invalid ghost rows or blocking in a later collective
        src = engine.max_tensor(k, off)
        m = torch.empty(k + 1, dtype=src.dtype, device=src.device)
        m[:k].copy_(src)
        engine.peer_error_into(m[k:])
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        if cuda:
            ev = torch.cuda.Event()
            ev.record()
            host = engine._host[blk & 1]
            with torch.cuda.stream(side):
                side.wait_event(ev)
                host[:k + 1].copy_(m, non_blocking=True)
                got = torch.cuda.Event()
                got.record(side)
            m.record_stream(side)

            def fetch(host=host, got=got, k=k):
                got.synchronize()
                return host[:k + 1].numpy().copy()
        else:
            def fetch(m=m):
                return m.numpy().copy()
        pending.append((k, launched, fetch))
        launched += k
        blk += 1

    while True:
        if launched < max_iterations and not conv and len(pending) < 2:
            launch()
            continue
        if not pending:
            break
        k, first, fetch = pending.pop(0)
        m_host = fetch()
        if m_host[k] != 0.0:
            raise RuntimeError("slab solve: a pass gave up waiting for a neighbouring rank's ghost rows "
                               f"(reported by at least one rank in sweeps {first + 1}..{first + k}; every rank stops here)")
        done = first + k
        if not conv:
            conv, last = _decide(m_host[:k], tol, first)
    return {"sweeps": done, "converged_at": conv, "last_max_update": last, "mode": "peer"}


def solve_local_peer(engines, max_iterations: int, tol: float, check_every: int = 64):
    """G slabs in THIS process on one GPU driving the fused-exchange kernels (attached with peer_connect_local):
    passes are issued slab by slab so that every wait finds its flag already raised or raised by an earlier launch."""
    import torch
    TS = engines[0].TS
    if not getattr(engines[0], "_local_connected", False):
        for g in range(len(engines) - 1):
            engines[g].peer_connect_local(1, engines[g + 1])
            engines[g + 1].peer_connect_local(0, engines[g])
        engines[0]._local_connected = True
    check_every = max(1, min(check_every, 4096))
    done, conv, last = 0, 0, 0.0
    while done < max_iterations:          # the stop is acted on one block late (see _solve_peer)
        was_conv = bool(conv)
        k = min(check_every, max_iterations - done)
        for e in engines:
            e.clear_max(k)
        for j in range(0, k, TS):
            for e in engines:
                e.peer_run(min(TS, k - j), j)
        m = engines[0].max_tensor(k).clone()
        for e in engines[1:]:
            m = torch.maximum(m, e.max_tensor(k))
        if not conv:
            conv, last = _decide(m.cpu().numpy(), tol, done)
        done += k
        if was_conv:
            break
    if any(e.peer_timed_out() for e in engines):
        raise RuntimeError("slab solve: a pass gave up waiting for a neighbouring slab's ghost rows")
    return {"sweeps": done, "converged_at": conv, "last_max_update": last, "mode": "peer"}


def solve_local(engines, max_iterations: int, tol: float, check_every: int = 64, mode: str = "auto"):
    """Single-process emulation of G slabs (all engines in this process, e.g. G slabs on ONE GPU): same phase
    structure as `solve`, ghost rows copied directly between the engines."""
    import torch
    G = len(engines)
    e0 = engines[0]
    min_rows = min(e.rows for e in engines)
    wave = (mode == "wavefront") or (mode == "auto" and all((not e.has_nan) and hasattr(e, "pass_") for e in engines)
                                     and min_rows >= e0.GH)
    TS = e0.TS if wave else 1
    check_every = max(TS, min(check_every, 4096))
    check_every = (check_every + TS - 1) // TS * TS

    def exchange(depth):
        for g in range(G - 1):
            up, dn = engines[g], engines[g + 1]
            GH = up.GH
            dn.phi[GH - depth:GH].copy_(up.phi[GH + up.rows - depth:GH + up.rows])
            up.phi[GH + up.rows:GH + up.rows + depth].copy_(dn.phi[GH:GH + depth])

    done, conv, last = 0, 0, 0.0
    while done < max_iterations:          # the stop is acted on one block late (see _solve_peer)
        was_conv = bool(conv)
        k = min(check_every, max_iterations - done)
        for e in engines:
            e.clear_max(k)
        if wave:
            j = 0
            while j < k:
                ns = min(TS, k - j)
                for e in engines:
                    if hasattr(e, "pass_part") and e.bands()[1] is not None:   # same banded pass as the overlapped path
                        eb, inner = e.bands()
                        for (rb, rc) in eb:
                            e.pass_part(ns, j, rb, rc)
                        e.pass_part(ns, j, inner[0], inner[1])
                        e.flip()
                    else:
                        e.pass_(ns, j)
                exchange(e0.GH)
                j += ns
        else:
            for j in range(k):
                for colour in (0, 1):
                    for e in engines:
                        e.sweep_colour(colour, j)
                    exchange(1)
        m = engines[0].max_tensor(k).clone()
        for e in engines[1:]:
            m = torch.maximum(m, e.max_tensor(k).to(m.device))
        if not conv:
            conv, last = _decide(m.cpu().numpy(), tol, done)
        done += k
        if was_conv:
            break
    return {"sweeps": done, "converged_at": conv, "last_max_update": last, "mode": "wavefront" if wave else "colour"}


SOLVE_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p)


class SlabSolveHook:
    """Multi-GPU caustic design: every rank runs the same (replicated, cheap) stages of the transport and height
    iterations on its own GPU, and the Poisson solves -- 98 % of the time at 8192 x 8192 -- are spread over all ranks
    as row slabs.  Installed as the context's solve hook (include/pcd.h: pcd_set_solve_hook), so
    ``perform_transport_iteration``, ``run_transport`` and ``perform_height_map_iteration`` use it unchanged.  With the
    default ``check_every`` the sweep schedule equals the single-GPU large-grid solver's, so every rank ends up with
    the bits a single GPU would have produced."""

    def __init__(self, design, dist, rank: int, world: int, device: int, check_every: int = 64, mode: str = "auto"):
        from . import pcd_solve_info
        self._info_t = pcd_solve_info
        self.design, self.dist, self.rank, self.world, self.device = design, dist, rank, world, device
        self.check_every, self.mode = check_every, mode
        self.W, self.H = design._cfg.res_x, design._cfg.res_y
        self.row0, self.rows = partition(self.H, world, rank)
        self.engine = CudaSlabEngine(self.W, self.H, self.row0, self.rows, device)
        self.error = None
        self.solves = []
        self._cb = SOLVE_HOOK(self._solve)   # must outlive the installation
        _check(lib().pcd_set_solve_hook(design._h, C.cast(self._cb, C.c_void_p), None))

    def close(self):
        if self.design is not None and self.design._h:
            lib().pcd_set_solve_hook(self.design._h, None, None)
        self.engine.close()
        self.design = None

    def _solve(self, user, D_dev, phi_dev, W, H, max_it, tol, info_ptr):
        import torch
        try:
            eng = self.engine
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.load_device(D_dev, phi_dev)    # also clears the slab's error word: a failed solve does not poison the next
            e0.record()
            r = solve(eng, self.dist, self.rank, self.world, max_it, tol, self.check_every, self.mode)
            e1.record()
            eng.store_device(phi_dev)
            if self.world > 1:   # every rank continues with the whole field
                # (a stalled neighbour makes solve() raise on EVERY rank in the same block -- the error word rides in
                # the all-reduce of the maxima -- so no rank reaches these broadcasts alone)
                phi = torch.as_tensor(_DevView(phi_dev, (H, W), "<f8"), device=torch.device("cuda", self.device))
                for src in range(self.world):
                    r0, n = partition(H, self.world, src)
                    self.dist.broadcast(phi[r0:r0 + n], src=src)
            torch.cuda.current_stream().synchronize()
            ms = e0.elapsed_time(e1)
            info = C.cast(info_ptr, C.POINTER(self._info_t)).contents
            info.sweeps, info.converged_at, info.last_max_update = r["sweeps"], r["converged_at"], r["last_max_update"]
            info.device_ms = info.kernel_ms = ms
            info.launches = (r["sweeps"] + eng.TS - 1) // eng.TS
            info.path = 3   # PCD_SOLVER_TILED: the wavefront kernel family
            self.solves.append(dict(r, ms=ms))
            return 0
        except Exception as ex:   # never let an exception cross the C frame
            self.error = ex
            return 3
