"""ctypes binding of oracle/zl_oracle.c (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libzl_oracle.so")


class Det(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("w", C.c_float), ("h", C.c_float),
                ("confidence", C.c_float), ("class_id", C.c_int32)]


DET_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("w", "<f4"), ("h", "<f4"),
                      ("confidence", "<f4"), ("class_id", "<i4")])


def build(force=False):
    src = os.path.join(_HERE, "zl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.zlo_preprocess.restype = C.c_int
        _lib.zlo_preprocess.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib.zlo_stretch_index.restype = None
        _lib.zlo_stretch_index.argtypes = [C.c_int, C.c_int, C.c_void_p]
        _lib.zlo_iou.restype = C.c_float
        _lib.zlo_iou.argtypes = [C.c_void_p, C.c_void_p]
        _lib.zlo_decode_filter.restype = C.c_int
        _lib.zlo_decode_filter.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        _lib.zlo_nms.restype = C.c_int
        _lib.zlo_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        _lib.zlo_postprocess.restype = C.c_int
        _lib.zlo_postprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        _lib.zlo_postprocess_batch.restype = C.c_long
        _lib.zlo_postprocess_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def preprocess(img_bytes, width, height, mw, mh):
    """img_bytes: uint8 array of any length.  Returns (code, [3,mh,mw] fp32)."""
    img = np.ascontiguousarray(img_bytes, dtype=np.uint8).reshape(-1)
    out = np.zeros((3, mh, mw), np.float32)
    code = lib().zlo_preprocess(_p(img), img.size, width, height, mw, mh, _p(out))
    return code, out


def stretch_index(src_dim, dst_dim):
    out = np.zeros(dst_dim, np.int32)
    lib().zlo_stretch_index(src_dim, dst_dim, _p(out))
    return out


def iou(a, b):
    """a, b: (x, y, w, h) centre-format tuples."""
    da = np.zeros(1, DET_DTYPE)
    db = np.zeros(1, DET_DTYPE)
    da[0] = (*map(np.float32, a), 0, 0)
    db[0] = (*map(np.float32, b), 0, 0)
    return float(lib().zlo_iou(_p(da), _p(db)))


def decode_filter(raw, img_w, img_h, conf_thr):
    """raw: [4+nc, A] fp32.  Returns (dets, anchors) in anchor order."""
    raw = np.ascontiguousarray(raw, np.float32)
    nc, A = raw.shape[0] - 4, raw.shape[1]
    out = np.zeros(max(A, 1), DET_DTYPE)
    anc = np.zeros(max(A, 1), np.int32)
    n = lib().zlo_decode_filter(_p(raw), nc, A, img_w, img_h, conf_thr, _p(out), _p(anc))
    return out[:n].copy(), anc[:n].copy()


def nms(dets, anchors, iou_thr):
    dets = np.ascontiguousarray(dets, DET_DTYPE)
    n = len(dets)
    anchors = np.ascontiguousarray(anchors if anchors is not None else np.arange(n), np.int32)
    out = np.zeros(max(n, 1), DET_DTYPE)
    anc = np.zeros(max(n, 1), np.int32)
    k = lib().zlo_nms(_p(dets), _p(anchors), n, iou_thr, _p(out), _p(anc))
    return out[:k].copy(), anc[:k].copy()


def postprocess(raw, img_w, img_h, conf_thr, iou_thr):
    """raw: [4+nc, A] fp32 -> (kept dets in sorted order, their anchor indices)."""
    raw = np.ascontiguousarray(raw, np.float32)
    nc, A = raw.shape[0] - 4, raw.shape[1]
    out = np.zeros(max(A, 1), DET_DTYPE)
    anc = np.zeros(max(A, 1), np.int32)
    k = lib().zlo_postprocess(_p(raw), nc, A, img_w, img_h, conf_thr, iou_thr, _p(out), _p(anc))
    return out[:k].copy(), anc[:k].copy()


def postprocess_batch(raw, img_w, img_h, conf_thr, iou_thr):
    """raw: [n, 4+nc, A].  Returns (dets concatenated, counts[n])."""
    raw = np.ascontiguousarray(raw, np.float32)
    n, nc, A = raw.shape[0], raw.shape[1] - 4, raw.shape[2]
    iw = np.ascontiguousarray(np.broadcast_to(np.asarray(img_w, np.int32), (n,)))
    ih = np.ascontiguousarray(np.broadcast_to(np.asarray(img_h, np.int32), (n,)))
    out = np.zeros(max(n * A, 1), DET_DTYPE)
    counts = np.zeros(n, np.int32)
    total = lib().zlo_postprocess_batch(_p(raw), n, nc, A, _p(iw), _p(ih), conf_thr, iou_thr, _p(out), _p(counts))
    return out[:total].copy(), counts
