"""Seeded synthetic inputs.  Used by tests, smoke() and bench.py."""
from __future__ import annotations

import numpy as np


def frames_noise(n, h, w, seed=1234):
    """Set A: i.i.d. uniform bytes, [n,h,w,3] uint8 (BGR)."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def frames_structured(n, h, w, seed=5678):
    """Set B: gradients + 8-32 filled rectangles per frame."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(n):
        a, b, c = rng.uniform(0.2, 1.0, 3)
        base = np.stack([(xx * a * 255 / max(w - 1, 1)), (yy * b * 255 / max(h - 1, 1)),
                         ((xx + yy) * c * 255 / max(w + h - 2, 1))], -1)
        img = base.astype(np.uint8)
        for _ in range(int(rng.integers(8, 33))):
            x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
            rw, rh = int(rng.integers(4, max(w // 3, 5))), int(rng.integers(4, max(h // 3, 5)))
            img[y0:y0 + rh, x0:x0 + rw] = rng.integers(0, 256, 3, dtype=np.uint8)
        out[i] = img
    return out


def frames_const(n, h, w, value=128):
    """Set C: the reference's warm-up frame (onnx_engine.cpp:925)."""
    return np.full((n, h, w, 3), value, np.uint8)


def stress_head(n, nc, A, seed=42, img=640):
    """cfg5 decode/NMS stress tensor [n, 4+nc, A] fp32: scores ~ Beta(0.5, 8),
    cx,cy ~ U(0,img), w,h ~ LogNormal(ln 60, 0.6) clipped to [4, img]."""
    rng = np.random.default_rng(seed)
    raw = np.empty((n, 4 + nc, A), np.float32)
    raw[:, 0:2] = rng.uniform(0, img, size=(n, 2, A)).astype(np.float32)
    raw[:, 2:4] = np.clip(rng.lognormal(np.log(60.0), 0.6, size=(n, 2, A)), 4, img).astype(np.float32)
    raw[:, 4:] = rng.beta(0.5, 8.0, size=(n, nc, A)).astype(np.float32)
    return raw


def stress_head_adversarial(n, nc, A, seed=43, img=640, clusters=64):
    """Heavily overlapping clusters with exact score ties (tie-break and IoU==thr paths)."""
    rng = np.random.default_rng(seed)
    raw = np.zeros((n, 4 + nc, A), np.float32)
    cx = rng.uniform(40, img - 40, size=(n, clusters)).astype(np.float32)
    cy = rng.uniform(40, img - 40, size=(n, clusters)).astype(np.float32)
    which = rng.integers(0, clusters, size=(n, A))
    jitter = rng.integers(-3, 4, size=(n, 2, A)).astype(np.float32)          # integer jitter -> exact duplicates
    raw[:, 0] = np.take_along_axis(cx, which, 1) + jitter[:, 0]
    raw[:, 1] = np.take_along_axis(cy, which, 1) + jitter[:, 1]
    raw[:, 2] = 48.0 + 8.0 * rng.integers(0, 3, size=(n, A))
    raw[:, 3] = 64.0 + 8.0 * rng.integers(0, 3, size=(n, A))
    cls = rng.integers(0, min(nc, 4), size=(n, A))
    score = (rng.integers(1, 9, size=(n, A)) / 8.0).astype(np.float32) * 0.9   # 8 distinct values -> many ties
    np.put_along_axis(raw[:, 4:], cls[:, None, :], score[:, None, :], 1)
    return raw
