"""Test double for poisson_caustic_design_b200.slab: a numpy implementation of the local-slab engine interface
(same red-black colour update, same ghost-row layout, same two ways to advance) so that the distributed HOST
logic -- partition, ghost-row exchange, all-reduced stopping rule -- can be exercised on CPU with gloo.
Not part of the product."""
import numpy as np
import torch


class NumpySlabEngine:
    GH = 5   # ghost rows each side (= pcd_slab_ghost_rows())
    TS = 2   # sweeps per wavefront pass (= pcd_slab_sweeps_per_pass())

    def __init__(self, W, H, row0, rows):
        self.W, self.H, self.row0, self.rows = W, H, row0, rows
        self.LR = rows + 2 * self.GH
        self.phi = torch.zeros((self.LR, W), dtype=torch.float64)   # torch CPU tensor: gloo sends views of it
        self.D = np.zeros((self.LR, W))
        self._max = torch.zeros(4096, dtype=torch.float64)
        self.has_nan = False
        omega = 2.0 / (1.0 + 3.14159265 / W)
        with np.errstate(divide="ignore"):
            self.w = omega / np.arange(5, dtype=np.float64)

    def upload(self, D_g, phi_g):
        self.D = np.array(D_g, dtype=np.float64)
        self.phi.copy_(torch.from_numpy(np.array(phi_g, dtype=np.float64)))
        W, LR = self.W, self.LR
        gy = self.row0 - self.GH + np.arange(LR)
        self.gy = gy
        nan = np.isnan(self.D)
        m = np.zeros((LR, W, 4), dtype=bool)
        m[:, 1:, 0] = ~nan[:, :-1]
        m[1:, :, 1] = ~nan[:-1, :] & (gy[1:, None] != 0)
        m[:, :-1, 2] = ~nan[:, 1:]
        m[:-1, :, 3] = ~nan[1:, :] & (gy[:-1, None] != self.H - 1)
        self.m = m
        own = slice(self.GH, self.GH + self.rows)
        self.has_nan = bool(nan[max(self.GH - 1, 0):self.GH + self.rows + 1].any())

    def _update(self, colour, slot, lo, hi):
        """Colour update of local rows [lo, hi); the max only counts owned rows."""
        phi = self.phi.numpy()
        W, GH, rows = self.W, self.GH, self.rows
        r = np.arange(lo, hi)[:, None]
        x = np.arange(W)[None, :]
        gy = self.gy[lo:hi][:, None]
        active = (((x + gy + colour) & 1) == 0) & (gy >= 0) & (gy < self.H)
        p = np.pad(phi, ((0, 0), (1, 1)))
        own = phi[lo:hi]
        m = self.m[lo:hi]
        s = np.zeros_like(own)
        s = s + np.where(m[..., 0], p[lo:hi, :-2], 0.0)
        s = s + np.where(m[..., 1], phi[lo - 1:hi - 1], 0.0)
        s = s + np.where(m[..., 2], p[lo:hi, 2:], 0.0)
        s = s + np.where(m[..., 3], phi[lo + 1:hi + 1], 0.0)
        cnt = m.sum(axis=-1)
        with np.errstate(invalid="ignore", over="ignore"):
            delta = self.w[cnt] * (s - cnt.astype(np.float64) * own - self.D[lo:hi])
        ad = np.abs(delta)
        owned = (r >= GH) & (r < GH + rows)
        a = ad[active & owned & (ad > 0)]
        if a.size:
            self._max[slot] = max(float(self._max[slot]), float(a.max()))
        own[active] = (own + delta)[active]

    def sweep_colour(self, colour, slot):
        self._update(colour, slot, self.GH, self.GH + self.rows)

    def pass_(self, nsweeps, slot):
        # all local rows except the outermost ring: staleness creeps in one row per phase and never reaches the
        # owned rows (GH = 2*TS + 1)
        assert 1 <= nsweeps <= self.TS
        for t in range(nsweeps):
            for colour in (0, 1):
                self._update(colour, slot + t, 1, self.LR - 1)

    def clear_max(self, n, first=0):
        self._max[first:first + n] = 0.0

    def max_tensor(self, n, first=0):
        return self._max[first:first + n]

    def download(self):
        return self.phi.numpy()[self.GH:self.GH + self.rows].copy()


class NumpyPeerEngine(NumpySlabEngine):
    """Adds the fused-exchange interface (peer_export / peer_connect_ipc / peer_run / peer_timed_out): `peer_run`
    stands in for the pass kernels that push their ghost rows into the neighbour themselves -- here the push is a
    gloo send/recv issued from inside the engine, so slab.solve's peer path (handle all-gather, agreement, block
    structure, status check) runs on CPU."""

    def __init__(self, W, H, row0, rows, dist, rank, world, fail_connect=False, stall_after_runs=0):
        super().__init__(W, H, row0, rows)
        self.dist, self.rank, self.world, self.fail_connect = dist, rank, world, fail_connect
        self.stall_after_runs = stall_after_runs   # > 0: this rank's passes report a stalled neighbour from that run on
        self.neighbours = {}
        self.runs = 0

    def peer_export(self):
        h = np.zeros(128, dtype=np.uint8)
        h[:4] = np.frombuffer(np.int32(self.row0).tobytes(), dtype=np.uint8)   # something rank-specific to check
        return h

    def peer_connect_ipc(self, side, handles, peer_row0, peer_rows):
        if self.fail_connect:
            raise RuntimeError("cannot attach the neighbour")
        assert int(np.frombuffer(np.asarray(handles, dtype=np.uint8)[:4].tobytes(), dtype=np.int32)[0]) == peer_row0
        assert peer_row0 == (self.row0 - peer_rows if side == 0 else self.row0 + self.rows)
        self.neighbours[side] = (peer_row0, peer_rows)

    def peer_run(self, nsweeps, slot):
        from poisson_caustic_design_b200 import slab
        self.runs += 1
        for j in range(0, nsweeps, self.TS):
            self.pass_(min(self.TS, nsweeps - j), slot + j)
            slab._exchange(self, self.dist, self.rank, self.world, self.GH)

    def peer_timed_out(self):
        return bool(self.stall_after_runs and self.runs >= self.stall_after_runs)

    def peer_error_into(self, slot):
        slot[0] = 1.0 if self.peer_timed_out() else 0.0
