"""Test double for poisson_caustic_design_b200.slab: a numpy implementation of the local-slab engine interface
(same red-black colour update, same ghost-row layout) so that the distributed HOST logic -- partition, halo
exchange, all-reduced stopping rule -- can be exercised on CPU with gloo.  Not part of the product."""
import numpy as np
import torch


class NumpySlabEngine:
    def __init__(self, W, H, row0, rows):
        self.W, self.H, self.row0, self.rows = W, H, row0, rows
        self.phi = torch.zeros((rows + 2, W), dtype=torch.float64)   # torch CPU tensor: gloo sends views of it
        self.D = np.zeros((rows + 2, W))
        self._max = torch.zeros(4096, dtype=torch.float64)
        omega = 2.0 / (1.0 + 3.14159265 / W)
        with np.errstate(divide="ignore"):
            self.w = omega / np.arange(5, dtype=np.float64)

    def upload(self, D_g, phi_g):
        self.D = np.array(D_g, dtype=np.float64)
        self.phi.copy_(torch.from_numpy(np.array(phi_g, dtype=np.float64)))
        W, rows = self.W, self.rows
        gy = self.row0 + np.arange(rows + 2) - 1
        nan = np.isnan(self.D)
        m = np.zeros((rows + 2, W, 4), dtype=bool)
        m[:, 1:, 0] = ~nan[:, :-1]
        m[1:, :, 1] = ~nan[:-1, :] & (gy[1:, None] != 0)
        m[:, :-1, 2] = ~nan[:, 1:]
        m[:-1, :, 3] = ~nan[1:, :] & (gy[:-1, None] != self.H - 1)
        self.m = m

    def sweep_colour(self, colour, slot):
        phi = self.phi.numpy()
        W, rows = self.W, self.rows
        r = np.arange(1, rows + 1)[:, None]
        x = np.arange(W)[None, :]
        gy = self.row0 + r - 1
        active = ((x + gy + colour) & 1) == 0
        p = np.pad(phi, ((0, 0), (1, 1)))
        own = phi[1:rows + 1]
        m = self.m[1:rows + 1]
        s = np.zeros_like(own)
        s = s + np.where(m[..., 0], p[1:rows + 1, :-2], 0.0)
        s = s + np.where(m[..., 1], phi[0:rows], 0.0)
        s = s + np.where(m[..., 2], p[1:rows + 1, 2:], 0.0)
        s = s + np.where(m[..., 3], phi[2:rows + 2], 0.0)
        cnt = m.sum(axis=-1)
        with np.errstate(invalid="ignore", over="ignore"):
            delta = self.w[cnt] * (s - cnt.astype(np.float64) * own - self.D[1:rows + 1])
        ad = np.abs(delta)
        a = ad[active & (ad > 0)]
        if a.size:
            self._max[slot] = max(float(self._max[slot]), float(a.max()))
        own[active] = (own + delta)[active]

    def clear_max(self, n):
        self._max[:n] = 0.0

    def max_tensor(self, n):
        return self._max[:n]

    def download(self):
        return self.phi.numpy()[1:self.rows + 1].copy()
