"""ctypes binding of oracle/_ref/libzl_ref{,_fast}.so — the reference's OWN preProcess / postProcess / applyNMS /
calculateIoU text (cut by line range from /root/reference, see oracle/ref/extract.py and oracle/ref/shim.cpp).
TEST INFRASTRUCTURE ONLY: it pins oracle/zl_oracle.c and mints tests/golden/; nothing in the product loads it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .oracle_c import DET_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"


def so_path(fast=False):
    return os.path.join(_HERE, "_ref", "libzl_ref_fast.so" if fast else "libzl_ref.so")


def available(fast=False):
    """Builds oracle/_ref when the reference tree is present (this container); on the GPU box the prebuilt files travel."""
    if os.path.exists(os.path.join(REF_ROOT, "src", "inference", "onnx_engine.cpp")):
        src = [os.path.join(_HERE, "ref", "shim.cpp"), os.path.join(_HERE, "ref", "extract.py")]
        so = so_path(fast)
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return os.path.exists(so_path(fast))


_libs = {}


def lib(fast=False):
    if fast not in _libs:
        if not available(fast):
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        L = C.CDLL(so_path(fast))
        L.zlr_preprocess.restype = C.c_int
        L.zlr_preprocess.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.zlr_postprocess.restype = C.c_int
        L.zlr_postprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.zlr_nms.restype = C.c_int
        L.zlr_nms.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.zlr_iou.restype = C.c_float
        L.zlr_iou.argtypes = [C.c_void_p, C.c_void_p]
        L.zlr_sizeof_detection.restype = C.c_int
        _libs[fast] = L
    return _libs[fast]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def preprocess(img_bytes, width, height, mw, mh, fast=False):
    img = np.ascontiguousarray(img_bytes, dtype=np.uint8).reshape(-1)
    out = np.zeros((3, mh, mw), np.float32)
    code = lib(fast).zlr_preprocess(_p(img), img.size, width, height, mw, mh, _p(out))
    return code, out


def postprocess(raw, img_w, img_h, conf_thr, iou_thr, fast=False):
    """raw: [4+nc, A] fp32 -> kept detections in the reference's output order."""
    raw = np.ascontiguousarray(raw, np.float32)
    nc, A = raw.shape[0] - 4, raw.shape[1]
    out = np.zeros(max(A, 1), DET_DTYPE)
    k = lib(fast).zlr_postprocess(_p(raw), nc, A, img_w, img_h, conf_thr, iou_thr, _p(out))
    if k < 0:
        raise RuntimeError(f"reference postProcess failed with ErrorCode {-k}")
    return out[:k].copy()


def nms(dets, iou_thr, fast=False):
    dets = np.ascontiguousarray(dets, DET_DTYPE)
    out = np.zeros(max(len(dets), 1), DET_DTYPE)
    k = lib(fast).zlr_nms(_p(dets), len(dets), iou_thr, _p(out))
    return out[:k].copy()


def iou(a, b, fast=False):
    fa, fb = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return float(lib(fast).zlr_iou(_p(fa), _p(fb)))


def canonical(dets):
    """std::sort (onnx_engine.cpp:846-851) leaves the order inside a (class, confidence) tie group unspecified:
    compare detection lists after completing the order with the box fields."""
    d = np.asarray(dets, DET_DTYPE)
    order = np.lexsort((d["h"], d["w"], d["y"], d["x"], -d["confidence"].astype(np.float64), d["class_id"]))
    return d[order]
