"""ZLW1 weights container + the seeded synthetic-weight recipe.

TEST INFRASTRUCTURE (oracle side) — also used by bench.py to mint weights; the
product reads the same container in C++ (csrc/weights.cpp).

The reference loads an ultralytics YOLOv8 ONNX export (start.sh:122-125,
src/inference/onnx_engine.cpp:957-1062).  Neither the model nor `onnx` exist
offline, so weights are random-init per SURVEY.md §8d and stored BN-folded in
OIHW fp32 (the ONNX initialiser layout) under ultralytics' tensor names.

Container layout (little endian):
    "ZLW1" | u32 version=1 | u32 scale(0=n,1=s,2=m) | u32 nc | u32 n_tensors | u32 0
    per tensor: u32 name_len | name (padded with NULs to 4 B) | u32 ndim | u32 dims[ndim] | fp32 data
"""
from __future__ import annotations

import math
import struct
from collections import OrderedDict

import numpy as np

SCALES = {"n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768)}
SCALE_ID = {"n": 0, "s": 1, "m": 2}


def _ch(c, width, max_c):
    return int(math.ceil(min(c, max_c) * width / 8) * 8)


def _rep(n, depth):
    return max(round(n * depth), 1)


def conv_specs(scale: str, nc: int):
    """Every convolution of YOLOv8 in execution order.

    Returns a list of dicts: name, cin, cout, k, s, act (1 = SiLU, 0 = linear).
    """
    depth, width, max_c = SCALES[scale]
    c1, c2, c3, c4, c5 = (_ch(c, width, max_c) for c in (64, 128, 256, 512, 1024))
    specs = []

    def conv(name, cin, cout, k, s, act=1):
        specs.append(dict(name=name, cin=cin, cout=cout, k=k, s=s, act=act))

    def c2f(idx, cin, cout, n):
        c = cout // 2
        conv(f"model.{idx}.cv1.conv", cin, 2 * c, 1, 1)
        for j in range(n):
            conv(f"model.{idx}.m.{j}.cv1.conv", c, c, 3, 1)
            conv(f"model.{idx}.m.{j}.cv2.conv", c, c, 3, 1)
        conv(f"model.{idx}.cv2.conv", (2 + n) * c, cout, 1, 1)

    conv("model.0.conv", 3, c1, 3, 2)
    conv("model.1.conv", c1, c2, 3, 2)
    c2f(2, c2, c2, _rep(3, depth))
    conv("model.3.conv", c2, c3, 3, 2)
    c2f(4, c3, c3, _rep(6, depth))
    conv("model.5.conv", c3, c4, 3, 2)
    c2f(6, c4, c4, _rep(6, depth))
    conv("model.7.conv", c4, c5, 3, 2)
    c2f(8, c5, c5, _rep(3, depth))
    conv("model.9.cv1.conv", c5, c5 // 2, 1, 1)
    conv("model.9.cv2.conv", c5 * 2, c5, 1, 1)
    c2f(12, c5 + c4, c4, _rep(3, depth))
    c2f(15, c4 + c3, c3, _rep(3, depth))
    conv("model.16.conv", c3, c3, 3, 2)
    c2f(18, c3 + c4, c4, _rep(3, depth))
    conv("model.19.conv", c4, c4, 3, 2)
    c2f(21, c4 + c5, c5, _rep(3, depth))
    ch = (c3, c4, c5)
    cb = max(16, ch[0] // 4, 64)
    cc = max(ch[0], min(nc, 100))
    for l, cl in enumerate(ch):
        conv(f"model.22.cv2.{l}.0.conv", cl, cb, 3, 1)
        conv(f"model.22.cv2.{l}.1.conv", cb, cb, 3, 1)
        conv(f"model.22.cv2.{l}.2", cb, 64, 1, 1, act=0)
    for l, cl in enumerate(ch):
        conv(f"model.22.cv3.{l}.0.conv", cl, cc, 3, 1)
        conv(f"model.22.cv3.{l}.1.conv", cc, cc, 3, 1)
        conv(f"model.22.cv3.{l}.2", cc, nc, 1, 1, act=0)
    return specs


def conv_flops(scale: str, nc: int, h: int, w: int):
    """2*MAC over every conv for one h x w frame."""
    # spatial size of each conv output follows from the topology; recompute by walking strides
    from oracle.yolov8_ref import trace_shapes  # local import: torch-free callers use the table below
    total = 0
    for sp, (ho, wo) in zip(conv_specs(scale, nc), trace_shapes(scale, nc, h, w)):
        total += 2 * sp["cin"] * sp["cout"] * sp["k"] ** 2 * ho * wo
    return total


def make_weights(scale: str, nc: int, seed: int = 0, gain: float = 1.9, cls_bias: float = -2.0):
    """Seeded random-init folded weights, numpy PCG64 for reproducibility.

    conv ~ U(-b, b), b = gain*sqrt(3/fan_in)/sqrt(3) ... i.e. std = gain/sqrt(3*fan_in)*sqrt(3);
    BN gamma~U(0.5,1.5), beta~N(0,0.1), mean~N(0,0.1), var~U(0.5,1.5), eps=1e-3, folded:
        w' = w*gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps).
    Detect cv2.*.2 bias = 1.0; cv3.*.2 bias = cls_bias (calibrated so a few % of anchors pass 0.5).
    `gain` keeps activation variance roughly constant through the SiLU stack so
    the head sees a real signal instead of biases only.
    """
    rng = np.random.default_rng(seed)
    tensors = OrderedDict()
    for sp in conv_specs(scale, nc):
        cin, cout, k = sp["cin"], sp["cout"], sp["k"]
        fan_in = cin * k * k
        bound = gain / math.sqrt(fan_in)
        w = rng.uniform(-bound, bound, size=(cout, cin, k, k)).astype(np.float32)
        if sp["act"]:
            gamma = rng.uniform(0.5, 1.5, size=cout).astype(np.float32)
            beta = rng.normal(0.0, 0.1, size=cout).astype(np.float32)
            mean = rng.normal(0.0, 0.1, size=cout).astype(np.float32)
            var = rng.uniform(0.5, 1.5, size=cout).astype(np.float32)
            s = (gamma / np.sqrt(var + np.float32(1e-3))).astype(np.float32)
            w = (w * s[:, None, None, None]).astype(np.float32)
            b = (beta - mean * s).astype(np.float32)
        else:
            is_box = ".cv2." in sp["name"]
            b = np.full(cout, 1.0 if is_box else cls_bias, dtype=np.float32)
        tensors[sp["name"] + ".weight"] = np.ascontiguousarray(w)
        tensors[sp["name"] + ".bias"] = np.ascontiguousarray(b)
    return tensors


def dumps(tensors, scale: str, nc: int) -> bytes:
    out = [b"ZLW1", struct.pack("<5I", 1, SCALE_ID[scale], nc, len(tensors), 0)]
    for name, arr in tensors.items():
        nb = name.encode()
        pad = (-len(nb)) % 4
        arr = np.ascontiguousarray(arr, dtype="<f4")
        out.append(struct.pack("<I", len(nb)) + nb + b"\0" * pad)
        out.append(struct.pack("<I", arr.ndim) + struct.pack(f"<{arr.ndim}I", *arr.shape))
        out.append(arr.tobytes())
    return b"".join(out)


def loads(blob: bytes):
    assert blob[:4] == b"ZLW1", "bad magic"
    ver, scale_id, nc, n, _ = struct.unpack_from("<5I", blob, 4)
    assert ver == 1
    off = 24
    tensors = OrderedDict()
    for _ in range(n):
        (nl,) = struct.unpack_from("<I", blob, off)
        off += 4
        name = blob[off:off + nl].decode()
        off += nl + ((-nl) % 4)
        (nd,) = struct.unpack_from("<I", blob, off)
        off += 4
        dims = struct.unpack_from(f"<{nd}I", blob, off)
        off += 4 * nd
        cnt = int(np.prod(dims)) if nd else 1
        tensors[name] = np.frombuffer(blob, dtype="<f4", count=cnt, offset=off).reshape(dims).copy()
        off += 4 * cnt
    scale = {v: k for k, v in SCALE_ID.items()}[scale_id]
    return tensors, scale, nc


def save(path, tensors, scale, nc):
    with open(path, "wb") as f:
        f.write(dumps(tensors, scale, nc))


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())
