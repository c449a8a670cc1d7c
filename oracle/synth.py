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


# ---- version-independent generators for the golden fixtures (tests/golden/): pure integer / exact float32 arithmetic,
# ---- so the same inputs are rebuilt anywhere without depending on a numpy RNG stream
def _mix64(n, seed):
    x = (np.arange(n, dtype=np.uint64) + np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)) * np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def golden_bytes(shape, seed):
    """uint8 array of `shape` from a splitmix-style integer hash."""
    n = int(np.prod(shape))
    return (_mix64(n, seed) >> np.uint64(24)).astype(np.uint8).reshape(shape)


def golden_unit(shape, seed):
    """float32 in [0, 1) with a 24-bit mantissa (exactly representable)."""
    n = int(np.prod(shape))
    return ((_mix64(n, seed) >> np.uint64(40)).astype(np.float32) / np.float32(16777216.0)).reshape(shape)


def golden_head(nc, A, seed, img=640, ties=False, clusters=0):
    """Raw head tensor [4+nc, A] fp32 for the golden post-processing cases.  Scores = u^6 (about 1/3 of the anchors
    have a class >= 0.01 at nc=80); `clusters` > 0 snaps the boxes onto that many centres with integer jitter (heavy
    overlap); `ties` quantises the scores to multiples of 1/16 (exact (class, confidence) ties)."""
    raw = np.empty((4 + nc, A), np.float32)
    u = golden_unit((4, A), seed)
    if clusters:
        c = (golden_unit((2, clusters), seed + 1) * np.float32(img - 80) + np.float32(40)).astype(np.float32)
        which = (_mix64(A, seed + 2) % np.uint64(clusters)).astype(np.int64)
        jit = ((_mix64(2 * A, seed + 3) % np.uint64(7)).astype(np.float32) - np.float32(3)).reshape(2, A)
        raw[0] = c[0][which] + jit[0]
        raw[1] = c[1][which] + jit[1]
        raw[2] = np.float32(48) + np.float32(8) * (_mix64(A, seed + 4) % np.uint64(3)).astype(np.float32)
        raw[3] = np.float32(64) + np.float32(8) * (_mix64(A, seed + 5) % np.uint64(3)).astype(np.float32)
    else:
        raw[0:2] = u[0:2] * np.float32(img)
        raw[2:4] = np.float32(4) + u[2:4] * u[2:4] * np.float32(200)
    s = golden_unit((nc, A), seed + 7)
    s2 = s * s
    s = s2 * s2 * s2
    if ties:
        s = np.floor(s * np.float32(16)) / np.float32(16)
    raw[4:] = s
    return raw
