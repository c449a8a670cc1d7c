"""ctypes binding of libzl_b200.so — used by tests/, bench.py and smoke().

This is a thin mirror of include/zl_b200.h: no compute happens here and there is
no fallback.  If the shared library is missing or no CUDA device exists the
calls fail loudly (ZlError).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("ZL_B200_LIB") or os.path.join(_PKG, "lib", "libzl_b200.so")     # override: A/B runs against another build
TEST_LIB_PATH = os.path.join(_PKG, "lib", "libzl_b200_test.so")     # engine objects + unit-test / measurement hooks (include/zl_b200_test.h)

OK, INVALID_ARGUMENT, NOT_INITIALIZED = 0, 2, 3
INFERENCE_ERROR, MODEL_NOT_FOUND, MODEL_LOAD_FAILED, INVALID_INPUT = 200, 201, 202, 203
SYSTEM_ERROR, INSUFFICIENT_RESOURCES = 300, 303
SCALE = {"n": 0, "s": 1, "m": 2}
FP32, BF16, FP16 = 0, 1, 2

DET_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("w", "<f4"), ("h", "<f4"),
                      ("confidence", "<f4"), ("class_id", "<i4")])


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("model_w", C.c_int32), ("model_h", C.c_int32),
                ("num_classes", C.c_int32), ("scale", C.c_int32), ("precision", C.c_int32),
                ("conf_threshold", C.c_float), ("iou_threshold", C.c_float),
                ("class_weights", C.POINTER(C.c_float)), ("max_batch", C.c_int32),
                ("max_frame_w", C.c_int32), ("max_frame_h", C.c_int32), ("preprocess_mode", C.c_int32),
                ("queue_depth", C.c_int32), ("num_lanes", C.c_int32), ("use_graph", C.c_int32),
                ("batch_window_us", C.c_int32), ("emit_wire", C.c_int32), ("cpu_core_id", C.c_int32), ("high_priority", C.c_int32), ("reserved", C.c_int32 * 5)]


class Stats(C.Structure):
    _fields_ = [("inference_count", C.c_uint64), ("inference_errors", C.c_uint64), ("dropped_frames", C.c_uint64),
                ("queue_size", C.c_uint64), ("queue_high_water_mark", C.c_uint64), ("batches", C.c_uint64),
                ("avg_inference_time_ms", C.c_double), ("p99_inference_time_ms", C.c_double),
                ("avg_preprocessing_time_ms", C.c_double), ("avg_postprocessing_time_ms", C.c_double),
                ("avg_device_time_ms", C.c_double), ("graph_captured", C.c_int32), ("device", C.c_int32),
                ("precision", C.c_int32), ("running", C.c_int32)]


class OpProfile(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("kind", C.c_int32), ("launches", C.c_int32), ("ms", C.c_float),
                ("flops", C.c_double), ("bytes", C.c_double)]


RESULT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int32, C.c_void_p, C.c_int32)
WIRE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int32, C.c_void_p, C.c_size_t)
WIRE_HEADER_BYTES, WIRE_DET_BYTES = 14, 40

EXPORTS = [
    "zl_config_default", "zl_engine_create", "zl_engine_destroy", "zl_engine_load_weights",
    "zl_engine_load_weights_mem", "zl_engine_prepare_weights", "zl_engine_commit_weights", "zl_engine_discard_weights", "zl_engine_warmup", "zl_engine_set_callback", "zl_engine_submit",
    "zl_engine_set_wire_callback", "zl_infer_batch_wire", "zl_engine_queue_size", "zl_engine_drain", "zl_engine_get_stats", "zl_infer_batch", "zl_preprocess",
    "zl_forward_raw", "zl_decode_nms", "zl_engine_num_anchors", "zl_engine_upload_resident",
    "zl_engine_run_resident", "zl_engine_profile", "zl_engine_profile_stalls", "zl_bench_e2e", "zl_bench_h2d", "zl_bench_latency", "zl_bench_preprocess", "zl_bench_decode_nms",
    "zl_model_probe", "zl_host_alloc", "zl_host_free", "zl_last_error", "zl_version", "zl_device_count",
]
TEST_EXPORTS = ["zl_test_conv", "zl_test_sppf_pool", "zl_probe_umma", "zl_probe_tma"]     # libzl_b200_test.so only


class ZlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"zl_b200 error {code}: {msg}")
        self.code = code
        self.message = msg


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZlError(SYSTEM_ERROR, f"{LIB_PATH} not built — run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        vp, i32, u32, u64, f32, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_float, C.c_size_t
        sig = {
            "zl_config_default": (None, [C.POINTER(Config)]),
            "zl_engine_create": (i32, [C.POINTER(Config), C.POINTER(vp)]),
            "zl_engine_destroy": (i32, [vp]),
            "zl_engine_load_weights": (i32, [vp, C.c_char_p]),
            "zl_engine_load_weights_mem": (i32, [vp, vp, sz]),
            "zl_engine_prepare_weights": (i32, [vp, C.c_char_p]),
            "zl_engine_commit_weights": (i32, [vp]),
            "zl_engine_discard_weights": (i32, [vp]),
            "zl_engine_warmup": (i32, [vp, i32]),
            "zl_engine_set_callback": (i32, [vp, RESULT_FN, vp]),
            "zl_engine_submit": (i32, [vp, u32, u32, u64, i32, i32, vp, sz, i32]),
            "zl_engine_set_wire_callback": (i32, [vp, WIRE_FN, vp]),
            "zl_infer_batch_wire": (i32, [vp, vp, vp, vp, i32, vp, vp, u64, vp, sz, vp]),
            "zl_engine_queue_size": (sz, [vp]),
            "zl_engine_drain": (i32, [vp]),
            "zl_engine_get_stats": (i32, [vp, C.POINTER(Stats)]),
            "zl_infer_batch": (i32, [vp, vp, vp, vp, i32, vp, i32, vp, vp]),
            "zl_preprocess": (i32, [vp, vp, i32, i32, sz, vp]),
            "zl_forward_raw": (i32, [vp, vp, vp, vp, i32, vp]),
            "zl_decode_nms": (i32, [vp, vp, i32, i32, i32, vp, vp, f32, f32, vp, i32, vp, vp]),
            "zl_engine_num_anchors": (i32, [vp]),
            "zl_engine_upload_resident": (i32, [vp, i32, vp, vp, vp, i32]),
            "zl_engine_run_resident": (i32, [vp, i32, i32, C.POINTER(f32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
            "zl_engine_profile": (i32, [vp, i32, i32, C.POINTER(OpProfile), i32, C.POINTER(i32)]),
            "zl_engine_profile_stalls": (i32, [vp, i32, vp, i32, C.POINTER(i32)]),
            "zl_bench_e2e": (i32, [vp, vp, i32, i32, i32, i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
            "zl_bench_h2d": (i32, [vp, sz, i32, C.POINTER(C.c_double)]),
            "zl_bench_latency": (i32, [vp, vp, i32, i32, i32, i32, vp]),
            "zl_bench_preprocess": (i32, [vp, i32, i32, i32, i32, C.POINTER(f32), C.POINTER(C.c_double)]),
            "zl_bench_decode_nms": (i32, [vp, vp, i32, i32, i32, f32, f32, i32, C.POINTER(f32), C.POINTER(f32), C.POINTER(C.c_int64)]),
            "zl_model_probe": (i32, [vp, sz, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(u64)]),
            "zl_host_alloc": (vp, [sz]),
            "zl_host_free": (None, [vp]),
            "zl_last_error": (C.c_char_p, []),
            "zl_version": (C.c_char_p, []),
            "zl_device_count": (i32, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


_testlib = None


def testlib():
    """The test library: the same engine objects plus the unit-test hooks the product library does not export."""
    global _testlib
    if _testlib is None:
        if not os.path.exists(TEST_LIB_PATH):
            raise ZlError(SYSTEM_ERROR, f"{TEST_LIB_PATH} not built — run __graft_entry__.build()")
        L = C.CDLL(TEST_LIB_PATH)
        vp, i32, u32 = C.c_void_p, C.c_int32, C.c_uint32
        L.zl_test_conv.restype, L.zl_test_conv.argtypes = i32, [i32, i32, vp, i32, i32, i32, i32, vp, vp, i32, i32, i32, i32, vp, vp]
        L.zl_test_sppf_pool.restype, L.zl_test_sppf_pool.argtypes = i32, [i32, i32, vp, i32, i32, i32, i32, vp]
        L.zl_probe_umma.restype, L.zl_probe_umma.argtypes = i32, [i32, i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.zl_probe_tma.restype, L.zl_probe_tma.argtypes = i32, [i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, u32, vp, u32, C.POINTER(i32)]
        L.zl_last_error.restype, L.zl_last_error.argtypes = C.c_char_p, []
        _testlib = L
    return _testlib


def _check_t(rc):
    if rc != OK:
        raise ZlError(rc, (testlib().zl_last_error() or b"").decode(errors="replace"))


def _check(rc):
    if rc != OK:
        raise ZlError(rc, (lib().zl_last_error() or b"").decode(errors="replace"))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pinned_array(shape, dtype=np.uint8):
    """numpy array over pinned host memory from zl_host_alloc (never freed: test/bench lifetime)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib().zl_host_alloc(n)
    if not p:
        raise ZlError(INSUFFICIENT_RESOURCES, "zl_host_alloc failed")
    buf = (C.c_uint8 * n).from_address(p)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Engine:
    def __init__(self, model_w=416, model_h=416, nc=4, scale="n", precision=BF16, conf=0.5, iou=0.45,
                 max_batch=1, device=0, max_frame=(0, 0), queue_depth=8, num_lanes=1, use_graph=1,
                 batch_window_us=0, class_weights=None, emit_wire=False, letterbox=False):
        L = lib()
        cfg = Config()
        L.zl_config_default(C.byref(cfg))
        cfg.device, cfg.model_w, cfg.model_h, cfg.num_classes = device, model_w, model_h, nc
        cfg.scale, cfg.precision = SCALE[scale], precision
        cfg.conf_threshold, cfg.iou_threshold = conf, iou
        cfg.max_batch, cfg.max_frame_w, cfg.max_frame_h = max_batch, max_frame[0], max_frame[1]
        cfg.queue_depth, cfg.num_lanes, cfg.use_graph, cfg.batch_window_us = queue_depth, num_lanes, use_graph, batch_window_us
        cfg.emit_wire = 1 if emit_wire else 0
        cfg.preprocess_mode = 1 if letterbox else 0          # ZL_PRE_LETTERBOX is NOT a parity mode (the reference stretches)
        self._cw = None
        if class_weights is not None:
            self._cw = np.ascontiguousarray(class_weights, np.float32)
            cfg.class_weights = self._cw.ctypes.data_as(C.POINTER(C.c_float))
        self.h = C.c_void_p()
        _check(L.zl_engine_create(C.byref(cfg), C.byref(self.h)))
        self.nc, self.model_w, self.model_h, self.max_batch = nc, model_w, model_h, max_batch
        self.A = L.zl_engine_num_anchors(self.h)
        self._cb = None

    def close(self):
        if self.h:
            lib().zl_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_weights_blob(self, blob: bytes):
        buf = np.frombuffer(blob, np.uint8)
        _check(lib().zl_engine_load_weights_mem(self.h, _ptr(buf), buf.size))

    def load_weights(self, path: str):
        _check(lib().zl_engine_load_weights(self.h, path.encode()))

    def prepare_weights(self, path: str):
        _check(lib().zl_engine_prepare_weights(self.h, path.encode()))

    def commit_weights(self):
        _check(lib().zl_engine_commit_weights(self.h))

    def discard_weights(self):
        _check(lib().zl_engine_discard_weights(self.h))

    def warmup(self, iters=3):
        _check(lib().zl_engine_warmup(self.h, iters))

    @staticmethod
    def _frame_args(frames):
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
        n = len(frames)
        ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
        ws = np.array([f.shape[1] for f in frames], np.int32)
        hs = np.array([f.shape[0] for f in frames], np.int32)
        return frames, ptrs, ws, hs, n

    def infer(self, frames, capacity=None):
        """frames: list of [h,w,3] uint8 BGR.  Returns list of structured det arrays (one per frame)."""
        frames, ptrs, ws, hs, n = self._frame_args(frames)
        cap = capacity if capacity is not None else n * self.A
        dets = np.empty(max(cap, 1), DET_DTYPE)          # the library fills [0, sum(counts)); nothing else is read
        counts = np.zeros(n, np.int32)
        offs = np.zeros(n, np.int32)
        _check(lib().zl_infer_batch(self.h, ptrs, _ptr(ws), _ptr(hs), n, _ptr(dets), cap, _ptr(counts), _ptr(offs)))
        return [dets[offs[i]:offs[i] + counts[i]].copy() for i in range(n)]

    def infer_wire(self, frames, frame_ids, timestamps, det_timestamp_ms):
        """Results in the reference's wire layout (DetectionResultPacket body per frame).  Returns a list of bytes objects."""
        frames, ptrs, ws, hs, n = self._frame_args(frames)
        ids = np.ascontiguousarray(frame_ids, np.uint32)
        tss = np.ascontiguousarray(timestamps, np.uint64)
        cap = n * (WIRE_HEADER_BYTES + WIRE_DET_BYTES * self.A)
        out = np.empty(cap, np.uint8)
        offs = np.zeros(n + 1, np.uint32)
        _check(lib().zl_infer_batch_wire(self.h, ptrs, _ptr(ws), _ptr(hs), n, _ptr(ids), _ptr(tss), int(det_timestamp_ms), _ptr(out), cap, _ptr(offs)))
        return [out[offs[i]:offs[i + 1]].tobytes() for i in range(n)]

    def set_wire_callback(self, fn):
        """fn(client_id, frame_id, timestamp, status, body bytes)"""
        def tramp(user, cid, fid, ts, status, bptr, nbytes):
            body = bytes((C.c_uint8 * nbytes).from_address(bptr)) if nbytes and bptr else b""
            fn(cid, fid, ts, status, body)
        self._wcb = WIRE_FN(tramp)
        _check(lib().zl_engine_set_wire_callback(self.h, self._wcb, None))

    def forward_raw(self, frames):
        frames, ptrs, ws, hs, n = self._frame_args(frames)
        raw = np.zeros((n, 4 + self.nc, self.A), np.float32)
        _check(lib().zl_forward_raw(self.h, ptrs, _ptr(ws), _ptr(hs), n, _ptr(raw)))
        return raw

    def preprocess(self, frame_bytes, width, height):
        buf = np.ascontiguousarray(frame_bytes, np.uint8).reshape(-1)
        out = np.zeros((3, self.model_h, self.model_w), np.float32)
        _check(lib().zl_preprocess(self.h, _ptr(buf), width, height, buf.size, _ptr(out)))
        return out

    def decode_nms(self, raw, img_w, img_h, conf, iou):
        raw = np.ascontiguousarray(raw, np.float32)
        n, nc, A = raw.shape[0], raw.shape[1] - 4, raw.shape[2]
        iw = np.ascontiguousarray(np.broadcast_to(np.asarray(img_w, np.int32), (n,)))
        ih = np.ascontiguousarray(np.broadcast_to(np.asarray(img_h, np.int32), (n,)))
        dets = np.zeros(n * A, DET_DTYPE)
        counts = np.zeros(n, np.int32)
        offs = np.zeros(n, np.int32)
        _check(lib().zl_decode_nms(self.h, _ptr(raw), n, nc, A, _ptr(iw), _ptr(ih), conf, iou, _ptr(dets), n * A, _ptr(counts), _ptr(offs)))
        return [dets[offs[i]:offs[i] + counts[i]].copy() for i in range(n)]

    # ---- async path
    def set_callback(self, fn):
        """fn(client_id, frame_id, timestamp, status, dets ndarray)"""
        def tramp(user, cid, fid, ts, status, dptr, n):
            d = np.zeros(0, DET_DTYPE)
            if n > 0 and dptr:
                d = np.frombuffer((C.c_uint8 * (n * DET_DTYPE.itemsize)).from_address(dptr), DET_DTYPE).copy()
            fn(cid, fid, ts, status, d)
        self._cb = RESULT_FN(tramp)
        _check(lib().zl_engine_set_callback(self.h, self._cb, None))

    def submit(self, client_id, frame_id, timestamp, frame, width=None, height=None, nbytes=None):
        f = np.ascontiguousarray(frame, np.uint8)
        h = height if height is not None else f.shape[0]
        w = width if width is not None else f.shape[1]
        return lib().zl_engine_submit(self.h, client_id, frame_id, timestamp, w, h, _ptr(f), nbytes if nbytes is not None else f.size, 0)

    def drain(self):
        _check(lib().zl_engine_drain(self.h))

    def queue_size(self):
        return lib().zl_engine_queue_size(self.h)

    def stats(self):
        s = Stats()
        _check(lib().zl_engine_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    # ---- measurement
    def upload_resident(self, set_idx, frames):
        frames, ptrs, ws, hs, n = self._frame_args(frames)
        _check(lib().zl_engine_upload_resident(self.h, set_idx, ptrs, _ptr(ws), _ptr(hs), n))

    def run_resident(self, n_sets, steps):
        ms, launches, dets = C.c_float(), C.c_int64(), C.c_int64()
        _check(lib().zl_engine_run_resident(self.h, n_sets, steps, C.byref(ms), C.byref(launches), C.byref(dets)))
        return ms.value, launches.value, dets.value

    def profile(self, set_idx=0, iters=3):
        arr = (OpProfile * 256)()
        n = C.c_int32()
        _check(lib().zl_engine_profile(self.h, set_idx, iters, arr, 256, C.byref(n)))
        return [dict(name=arr[i].name.decode(), kind=arr[i].kind, ms=arr[i].ms, flops=arr[i].flops, bytes=arr[i].bytes)
                for i in range(n.value)]

    def profile_stalls(self, set_idx=0):
        """[n_ops, 16] uint64 cycle counters of the instrumented persistent conv kernel (see zl_b200.h)."""
        out = np.zeros((256, 16), np.uint64)
        n = C.c_int32()
        _check(lib().zl_engine_profile_stalls(self.h, set_idx, _ptr(out), 256, C.byref(n)))
        return out[:n.value]

    def bench_e2e(self, batches, steps_total):
        """batches: list (one per host thread / lane) of pinned uint8 arrays [n,h,w,3].  Returns (seconds, dets of the last step)."""
        n, h, w = batches[0].shape[:3]
        ptrs = (C.c_void_p * len(batches))(*[b.ctypes.data for b in batches])
        sec, nd = C.c_double(), C.c_int64()
        _check(lib().zl_bench_e2e(self.h, ptrs, len(batches), n, w, h, steps_total, C.byref(sec), C.byref(nd)))
        return sec.value, nd.value

    def bench_h2d(self, nbytes=256 << 20, iters=8):
        g = C.c_double()
        _check(lib().zl_bench_h2d(self.h, nbytes, iters, C.byref(g)))
        return g.value

    def bench_latency(self, frame, warmup=50, iters=500):
        f = frame if isinstance(frame, np.ndarray) else np.ascontiguousarray(frame, np.uint8)
        out = np.zeros(iters, np.float32)
        _check(lib().zl_bench_latency(self.h, _ptr(f), f.shape[1], f.shape[0], warmup, iters, _ptr(out)))
        return out

    def bench_preprocess(self, w, h, n, iters=20):
        ms, b = C.c_float(), C.c_double()
        _check(lib().zl_bench_preprocess(self.h, w, h, n, iters, C.byref(ms), C.byref(b)))
        return ms.value, b.value

    def bench_decode_nms(self, raw, conf, iou, iters=5):
        raw = np.ascontiguousarray(raw, np.float32)
        n, nc, A = raw.shape[0], raw.shape[1] - 4, raw.shape[2]
        mf, mn, kept = C.c_float(), C.c_float(), C.c_int64()
        _check(lib().zl_bench_decode_nms(self.h, _ptr(raw), n, nc, A, conf, iou, iters, C.byref(mf), C.byref(mn), C.byref(kept)))
        return mf.value, mn.value, kept.value


def test_conv(x_nhwc, w_ohwi, bias, stride=1, act=True, res=None, impl=0, device=0, out_f32=False, ntile_hint=0, fp16=False):
    """One conv through the engine's kernels. x: [n,h,w,cin] fp32, w: [cout,k,k,cin] fp32."""
    x = np.ascontiguousarray(x_nhwc, np.float32)
    w = np.ascontiguousarray(w_ohwi, np.float32)
    b = np.ascontiguousarray(bias, np.float32)
    n, h, wd, cin = x.shape
    cout, k = w.shape[0], w.shape[1]
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // stride + 1, (wd + 2 * pad - k) // stride + 1
    y = np.zeros((n, ho, wo, cout), np.float32)
    r = np.ascontiguousarray(res, np.float32) if res is not None else None
    flags = (1 if act else 0) | (2 if out_f32 else 0) | (4 if fp16 else 0) | (ntile_hint << 8)
    _check_t(testlib().zl_test_conv(device, impl, _ptr(x), n, h, wd, cin, _ptr(w), _ptr(b), cout, k, stride, flags,
                              _ptr(r) if r is not None else None, _ptr(y)))
    return y


def test_sppf_pool(x_nhwc, dtype="fp16", device=0):
    """SPPF pools of x [n,h,w,c] through the engine's kernel: returns the concat buffer [n,h,w,4c] = x | p1 | p2 | p3 (fp32)."""
    x = np.ascontiguousarray(x_nhwc, np.float32)
    n, h, w, c = x.shape
    cat = np.zeros((n, h, w, 4 * c), np.float32)
    _check_t(testlib().zl_test_sppf_pool(device, {"fp32": 0, "bf16": 1, "fp16": 2}[dtype], _ptr(x), n, h, w, c, _ptr(cat)))
    return cat


def probe_umma(N, swz=128, sbo=None, nacc=1, count=512, shift_rows=0, ksteps=4, grid=1, device=0):
    a, b = C.c_int64(), C.c_int64()
    _check_t(testlib().zl_probe_umma(device, N, swz, sbo if sbo is not None else 8 * swz, nacc, count, shift_rows, ksteps, grid, C.byref(a), C.byref(b)))
    return a.value / count, b.value / count


def probe_tma(x_f16, box, estride, swizzle, coords, expect_bytes, dump_bytes, device=0):
    """x_f16: [n,h,w,c] float16 array. Returns (completed, uint16 dump)."""
    x = np.ascontiguousarray(x_f16, np.float16)
    n, h, w, c = x.shape
    dump = np.zeros(dump_bytes, np.uint8)
    done = C.c_int32()
    _check_t(testlib().zl_probe_tma(device, _ptr(x), n, h, w, c, box[0], box[1], box[2], estride, swizzle,
                              coords[0], coords[1], coords[2], coords[3], expect_bytes, _ptr(dump), dump_bytes, C.byref(done)))
    return done.value, dump.view(np.float16)


def model_probe(blob: bytes):
    """Host-only: (scale, nc, n_tensors, checksum) of a ZLW1 or ONNX model file."""
    buf = np.frombuffer(blob, np.uint8)
    sc, nc, nt, ck = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
    _check(lib().zl_model_probe(_ptr(buf), buf.size, C.byref(sc), C.byref(nc), C.byref(nt), C.byref(ck)))
    return sc.value, nc.value, nt.value, ck.value
