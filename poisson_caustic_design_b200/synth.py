"""Deterministic synthetic density images (SURVEY 8d / BASELINE.json configs[3..4]).

``synth_density(W, H, seed)``: fp64 image at DOMAIN resolution (so main.cpp's nearest resize is the
identity when res_w = W/4): background 0.02, 12 hard discs (value 1), 6 Gaussian blobs, 4 axis-aligned
bars (0.6), clipped to [0,1] and quantised to 8 bit like a PNG.  The random stream is splitmix64 on
``seed`` so that a C++ host produces the same image bit for bit.
"""
from __future__ import annotations

import numpy as np

_M = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _M

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
        return z ^ (z >> 31)

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        return lo + (hi - lo) * ((self.next() >> 11) * (1.0 / (1 << 53)))


def synth_density(W: int, H: int, seed: int, background: float = 0.02) -> np.ndarray:
    rng = SplitMix64(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    img = np.full((H, W), background, dtype=np.float64)
    for _ in range(12):                                   # hard discs
        cx, cy, r = rng.uniform() * W, rng.uniform() * H, rng.uniform(0.03, 0.12) * W
        img[(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = 1.0
    for _ in range(6):                                    # Gaussian blobs
        cx, cy = rng.uniform() * W, rng.uniform() * H
        s, a = rng.uniform(0.02, 0.08) * W, rng.uniform(0.3, 1.0)
        img += a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2.0 * s * s))
    for _ in range(4):                                    # axis-aligned bars
        x0, y0 = rng.uniform() * W, rng.uniform() * H
        if rng.uniform() < 0.5:
            w, h = rng.uniform(0.2, 0.5) * W, rng.uniform(0.01, 0.03) * H
        else:
            w, h = rng.uniform(0.01, 0.03) * W, rng.uniform(0.2, 0.5) * H
        img[(np.abs(xx - x0) < w / 2) & (np.abs(yy - y0) < h / 2)] = 0.6
    img = np.clip(img, 0.0, 1.0)
    return np.round(img * 255.0) / 255.0


class Setup:
    """main.cpp:216-231 for an image that is already at domain resolution (4*res_w wide)."""

    def __init__(self, res_w: int, img_w: int, img_h: int, mesh_width: float = 1.0, focal_l: float = 1.5,
                 thickness: float = 0.2):
        import math
        aspect = float(img_w) / float(img_h)
        self.mesh_nx = res_w
        self.mesh_ny = int(res_w / aspect)
        self.res_x = 4 * res_w
        self.res_y = int(4 * res_w / aspect)
        self.width = float(np.float32(mesh_width))          # CLI floats, main.cpp:147-151
        self.height = math.floor(res_w / aspect) * (self.width / res_w)
        self.focal_l = float(np.float32(focal_l))
        self.thickness = float(np.float32(thickness))
