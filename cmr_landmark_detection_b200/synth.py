"""Seeded synthetic inputs shaped like the reference generator's output contract
(src/data/Generators.py:228, :376-398): x float32 [B,H,W,1] in [0,1] (per-slice min-max, Generators.py:379),
y float32 [B,H,W,2] Gaussian heat maps: binary disk -> scipy gaussian_filter(sigma) per channel -> ONE joint
min-max over (H,W,C) (Generators.py:385-391, Preprocess.py:491)."""
from __future__ import annotations

import numpy as np
from scipy.ndimage import gaussian_filter


def make_batch(B: int, H: int, W: int, seed: int = 42, sigma: float = 2.0, channels: int = 2):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    x = np.zeros((B, H, W, 1), np.float32)
    y = np.zeros((B, H, W, channels), np.float32)
    margin = max(2, min(H, W) // 8)
    for b in range(B):
        img = np.zeros((H, W), np.float32)
        for _ in range(int(rng.integers(6, 11))):
            cy, cx = rng.uniform(0, H), rng.uniform(0, W)
            sy, sx = rng.uniform(H / 32, H / 5), rng.uniform(W / 32, W / 5)
            img += rng.uniform(0.3, 1.0) * np.exp(-((yy - cy) ** 2 / (2 * sy * sy) + (xx - cx) ** 2 / (2 * sx * sx)))
        img += 0.05 * rng.random((H, W), dtype=np.float32)
        img = (img - img.min()) / (img.max() - img.min() + 1e-12)
        x[b, :, :, 0] = img
        ys = np.sort(rng.uniform(margin, H - margin, size=channels))      # anterior above inferior
        for c in range(channels):
            cx = rng.uniform(margin, W - margin)
            disk = ((yy - ys[c]) ** 2 + (xx - cx) ** 2 <= 9.0).astype(np.float32)
            y[b, :, :, c] = gaussian_filter(disk, sigma)
        y[b] = (y[b] - y[b].min()) / (y[b].max() - y[b].min() + 1e-12)
    return x, y


def make_volume_heat(Z: int, H: int, W: int, seed: int = 7, ragged: bool = True) -> np.ndarray:
    """A [Z,H,W,2] fp32 volume of plausible predicted heat maps (some slices miss a landmark)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    heat = np.zeros((Z, H, W, 2), np.float32)
    for z in range(Z):
        for c in range(2):
            cy, cx = rng.uniform(0.15 * H, 0.85 * H), rng.uniform(0.15 * W, 0.85 * W)
            s = rng.uniform(1.5, 4.0)
            amp = rng.uniform(0.7, 0.99)
            if ragged and rng.random() < 0.2:
                amp = rng.uniform(0.1, 0.45)
            heat[z, :, :, c] = amp * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    heat += 0.02 * rng.random(heat.shape, dtype=np.float32)
    return heat
