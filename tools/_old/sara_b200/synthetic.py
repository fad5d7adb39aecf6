"""Seeded synthetic frames; host-side numpy only.

`grad(w, h)`  : I(x, y) = (x + y) / (w + h - 2)          (config C2, ~0 keypoints)
`tex(w, h, seed, shift)` : 0.5 + sum of w*h/400 isotropic Gaussian blobs
    (centre ~U(image), sigma ~logU[1.5, 16] px, amplitude ~U[-0.35, 0.35]) plus
    N(0, 0.01) pixel noise, clipped to [0, 1].  `shift=(dx, dy)` translates the
    blob centres analytically, which gives an exact SfM-like sequence.
"""
from __future__ import annotations

import numpy as np


def grad(w: int = 1920, h: int = 1080) -> np.ndarray:
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    return ((x + y) / float(w + h - 2)).astype(np.float32)


def tex(w: int, h: int, seed: int = 1234, shift=(0.0, 0.0), noise_seed=None) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = max(1, (w * h) // 400)
    cx = rng.uniform(0, w, n) + shift[0]
    cy = rng.uniform(0, h, n) + shift[1]
    sg = np.exp(rng.uniform(np.log(1.5), np.log(16.0), n))
    am = rng.uniform(-0.35, 0.35, n)
    img = np.full((h, w), 0.5, np.float64)
    for i in range(n):
        r = int(np.ceil(4.0 * sg[i]))
        x0, x1 = int(np.floor(cx[i])) - r, int(np.floor(cx[i])) + r + 1
        y0, y1 = int(np.floor(cy[i])) - r, int(np.floor(cy[i])) + r + 1
        x0c, x1c, y0c, y1c = max(x0, 0), min(x1, w), max(y0, 0), min(y1, h)
        if x0c >= x1c or y0c >= y1c:
            continue
        xs = np.arange(x0c, x1c, dtype=np.float64) - cx[i]
        ys = np.arange(y0c, y1c, dtype=np.float64) - cy[i]
        inv = 1.0 / (2.0 * sg[i] * sg[i])
        img[y0c:y1c, x0c:x1c] += am[i] * np.outer(np.exp(-ys * ys * inv), np.exp(-xs * xs * inv))
    nrng = np.random.default_rng(seed + 7919 if noise_seed is None else noise_seed)
    img += nrng.normal(0.0, 0.01, (h, w))
    return np.clip(img, 0.0, 1.0).astype(np.float32)


def sequence_frame(w: int, h: int, i: int, seed: int = 1000) -> np.ndarray:
    """Frame i of the translated-scene sequence (configs C4/C5)."""
    return tex(w, h, seed=seed, shift=(3.0 * i, 1.5 * i), noise_seed=seed + 100003 * (i + 1))
