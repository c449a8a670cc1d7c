"""CPU fp32 restatement of the YOLOv8 graph the reference executes through ONNX Runtime.

TEST INFRASTRUCTURE ONLY (see oracle/zl_oracle.c header for the rule).

The arithmetic of `Ort::Session::Run` (src/inference/onnx_engine.cpp:577-585)
is not in the reference tree: it is ONNX Runtime v1.8.1 (start.sh:74) running
an ultralytics YOLOv8 export (start.sh:122-125), neither vendored nor available
offline.  This module restates the published YOLOv8 definition (SURVEY.md
Appendix A) with torch CPU fp32 ops as "PyTorch-CPU stand-in for the ORT-CPU
session".  PARITY UNPINNED: no golden vector of the reference exists; the
restatement is pinned only against ultralytics' published FLOP / parameter
totals (tests/test_oracle_model.py).

I/O contract kept from the reference: input `images` [B,3,H,W] fp32 RGB in
[0,1] (onnx_engine.cpp:49,560); output `output0` [B,4+nc,A] fp32, rows 0-3 =
cx,cy,w,h in model-input pixels, rows 4.. = sigmoid class scores
(onnx_engine.cpp:50,767-784).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from oracle.zlw import SCALES, _ch, _rep, conv_specs


class _Net:
    """Functional forward over a dict of folded tensors (name -> np/torch array)."""

    def __init__(self, tensors, scale: str, nc: int, dtype=torch.float32):
        self.dtype = dtype
        self.t = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32).to(dtype) for k, v in tensors.items()}
        self.scale, self.nc = scale, nc
        depth, width, max_c = SCALES[scale]
        self.n = [_rep(3, depth), _rep(6, depth), _rep(6, depth), _rep(3, depth)]
        self.nh = _rep(3, depth)
        self.specs = {s["name"]: s for s in conv_specs(scale, nc)}
        self.trace = None

    def conv(self, name, x):
        sp = self.specs[name]
        y = F.conv2d(x, self.t[name + ".weight"], self.t[name + ".bias"], stride=sp["s"], padding=sp["k"] // 2)
        if sp["act"]:
            y = F.silu(y)
        if self.trace is not None:
            self.trace.append((name, tuple(y.shape[2:])))
        return y

    def c2f(self, idx, x, n, shortcut):
        t = self.conv(f"model.{idx}.cv1.conv", x)
        c = t.shape[1] // 2
        ys = [t[:, :c], t[:, c:]]
        for j in range(n):
            z = self.conv(f"model.{idx}.m.{j}.cv2.conv", self.conv(f"model.{idx}.m.{j}.cv1.conv", ys[-1]))
            ys.append(ys[-1] + z if shortcut else z)
        return self.conv(f"model.{idx}.cv2.conv", torch.cat(ys, 1))

    def sppf(self, x):
        a = self.conv("model.9.cv1.conv", x)
        p1 = F.max_pool2d(a, 5, 1, 2)
        p2 = F.max_pool2d(p1, 5, 1, 2)
        p3 = F.max_pool2d(p2, 5, 1, 2)
        return self.conv("model.9.cv2.conv", torch.cat([a, p1, p2, p3], 1))

    def features(self, x):
        n = self.n
        x = self.conv("model.0.conv", x)
        x = self.conv("model.1.conv", x)
        x = self.c2f(2, x, n[0], True)
        x = self.conv("model.3.conv", x)
        p3 = self.c2f(4, x, n[1], True)
        x = self.conv("model.5.conv", p3)
        p4 = self.c2f(6, x, n[2], True)
        x = self.conv("model.7.conv", p4)
        x = self.c2f(8, x, n[3], True)
        p5 = self.sppf(x)
        x = torch.cat([F.interpolate(p5, scale_factor=2, mode="nearest"), p4], 1)
        h12 = self.c2f(12, x, self.nh, False)
        x = torch.cat([F.interpolate(h12, scale_factor=2, mode="nearest"), p3], 1)
        o3 = self.c2f(15, x, self.nh, False)
        x = torch.cat([self.conv("model.16.conv", o3), h12], 1)
        o4 = self.c2f(18, x, self.nh, False)
        x = torch.cat([self.conv("model.19.conv", o4), p5], 1)
        o5 = self.c2f(21, x, self.nh, False)
        return [o3, o4, o5]

    def head_maps(self, feats):
        """Per level: (box logits [B,64,h,w], class logits [B,nc,h,w])."""
        boxes, clss = [], []
        for l, f in enumerate(feats):
            b = self.conv(f"model.22.cv2.{l}.2", self.conv(f"model.22.cv2.{l}.1.conv", self.conv(f"model.22.cv2.{l}.0.conv", f)))
            boxes.append(b)
        for l, f in enumerate(feats):
            c = self.conv(f"model.22.cv3.{l}.2", self.conv(f"model.22.cv3.{l}.1.conv", self.conv(f"model.22.cv3.{l}.0.conv", f)))
            clss.append(c)
        return boxes, clss


def dfl_decode(boxes, clss, strides=(8, 16, 32)):
    """Detect tail: DFL softmax expectation, dist2bbox(xywh), x stride, sigmoid."""
    B = boxes[0].shape[0]
    dt = boxes[0].dtype
    box = torch.cat([b.reshape(B, 64, -1) for b in boxes], 2)            # [B,64,A]
    cls = torch.cat([c.reshape(B, c.shape[1], -1) for c in clss], 2)     # [B,nc,A]
    A = box.shape[2]
    anc, strd = [], []
    for b, s in zip(boxes, strides):
        h, w = b.shape[2:]
        sy, sx = torch.meshgrid(torch.arange(h, dtype=dt) + 0.5,
                                torch.arange(w, dtype=dt) + 0.5, indexing="ij")
        anc.append(torch.stack([sx.reshape(-1), sy.reshape(-1)], 0))      # [2, h*w] (x, y)
        strd.append(torch.full((h * w,), float(s), dtype=dt))
    anc = torch.cat(anc, 1)                                               # [2,A]
    strd = torch.cat(strd)                                                # [A]
    prob = box.view(B, 4, 16, A).softmax(2)
    dist = (prob * torch.arange(16, dtype=dt).view(1, 1, 16, 1)).sum(2)   # [B,4,A] l,t,r,b
    lt, rb = dist[:, :2], dist[:, 2:]
    x1y1 = anc.unsqueeze(0) - lt
    x2y2 = anc.unsqueeze(0) + rb
    cxcy = (x1y1 + x2y2) / 2
    wh = x2y2 - x1y1
    dbox = torch.cat([cxcy, wh], 1) * strd.view(1, 1, A)
    return torch.cat([dbox, cls.sigmoid()], 1)                            # [B,4+nc,A]


@torch.inference_mode()
def forward_raw(tensors, scale, nc, images_nchw, return_maps=False, fp64=False):
    """images_nchw: [B,3,H,W] fp32 -> output0 [B,4+nc,A] fp32 (numpy).

    fp64=False: fp32 arithmetic end to end, like the reference's ORT CPU session.
    fp64=True : same fp32 weights and inputs, every op evaluated in float64 and the
    result rounded once to fp32 — the summation-order-free value both the fp32 CPU
    session and the CUDA exact mode approximate (see DESIGN.md "fp32 parity").
    """
    dt = torch.float64 if fp64 else torch.float32
    net = _Net(tensors, scale, nc, dtype=dt)
    x = torch.as_tensor(np.asarray(images_nchw), dtype=torch.float32).to(dt)
    boxes, clss = net.head_maps(net.features(x))
    out = dfl_decode(boxes, clss).to(torch.float32)
    if return_maps:
        return out.numpy(), [b.to(torch.float32).numpy() for b in boxes], [c.to(torch.float32).numpy() for c in clss]
    return out.numpy()


class Session:
    """Reusable stand-in for the reference's Ort::Session (weights converted once)."""

    def __init__(self, tensors, scale, nc, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.net = _Net(tensors, scale, nc)

    @torch.inference_mode()
    def run(self, images_nchw):
        x = torch.as_tensor(np.asarray(images_nchw), dtype=torch.float32)
        boxes, clss = self.net.head_maps(self.net.features(x))
        return dfl_decode(boxes, clss).numpy()


@torch.inference_mode()
def trace_shapes(scale, nc, h, w):
    """(ho, wo) of every conv in conv_specs() order, from a dry run on zeros."""
    specs = conv_specs(scale, nc)
    tensors = {}
    for sp in specs:
        tensors[sp["name"] + ".weight"] = np.zeros((sp["cout"], sp["cin"], sp["k"], sp["k"]), np.float32)
        tensors[sp["name"] + ".bias"] = np.zeros(sp["cout"], np.float32)
    net = _Net(tensors, scale, nc)
    net.trace = []
    net.head_maps(net.features(torch.zeros(1, 3, h, w)))
    got = dict(net.trace)
    return [got[sp["name"]] for sp in specs]


def num_anchors(h, w):
    return sum((h // s) * (w // s) for s in (8, 16, 32))


def _round_sig(v, sig=3):
    return float(f"{float(v):.{sig}g}")


@torch.inference_mode()
def calibrate(tensors, scale, nc, size=256, target_frac=0.02, cls_std=1.5, dfl_beta=0.5, dfl_mu0=4.0, dfl_mu_std=0.35):
    """Data-dependent rescale of random-init weights (LSUV style) so the synthetic
    model carries a real signal to the head instead of exploding / vanishing:
    every conv's (w, b) is multiplied by 1/std(pre-activation) measured on one
    fixed structured frame, the class head gets std `cls_std`, and each level's
    class bias is set so ~target_frac of its anchors score >= 0.5 (SURVEY.md §7
    "random-init weights give degenerate scores").  The DFL box head (cv2.*.2) is
    given the STRUCTURE a trained head has instead of 64 independent random rows:
    for every side s the 16 bin logits are  logit_k = alpha_s(x) * k - dfl_beta * k^2
    (+ a constant slope), alpha_s a linear function of the features (the seeded
    random row s*16 of the conv, rescaled), i.e. a discretised Gaussian over the bins
    with variance 1 / (2 dfl_beta) whose mean dfl_mu0 + alpha_s / (2 dfl_beta) moves
    with the input (std dfl_mu_std bins on the calibration frame).  Independent
    unit-variance logits give near-flat or multi-modal distributions: expectations of
    ~7.5 bins on every side (900-px boxes on a 416-px image) that are twenty times
    more sensitive to rounding than any trained detector's.  Scale factors are
    rounded to 3 significant digits so the result is bit-stable across CPUs.
    Returns a new tensor dict (numpy fp32).
    """
    from oracle import synth
    out = {k: np.array(v, dtype=np.float32, copy=True) for k, v in tensors.items()}
    net = _Net(out, scale, nc)
    fr = synth.frames_structured(1, size, size, seed=99)[0].astype(np.float32) / 255.0
    x = torch.as_tensor(np.ascontiguousarray(fr[..., ::-1].transpose(2, 0, 1))[None].copy())

    orig_conv = net.conv

    def conv(name, xin):
        sp = net.specs[name]
        w, b = net.t[name + ".weight"], net.t[name + ".bias"]
        y = F.conv2d(xin, w, b, stride=sp["s"], padding=sp["k"] // 2)
        if sp["act"]:
            s = _round_sig(1.0 / max(float(y.std()), 1e-12))
            w.mul_(s)
            b.mul_(s)
        elif ".cv2." in name:                   # DFL box head: unimodal bin distributions with a data-dependent mean
            w0 = w.clone()
            for side in range(4):
                row = w0[side * 16].clone()                                        # the side's seeded random direction
                alpha = F.conv2d(xin, row.view(1, -1, 1, 1), None)
                sc = _round_sig(dfl_mu_std * 2.0 * dfl_beta / max(float(alpha.std()), 1e-12))
                for kbin in range(16):
                    w[side * 16 + kbin] = row * (sc * kbin)
                    b[side * 16 + kbin] = -dfl_beta * kbin * kbin + 2.0 * dfl_beta * dfl_mu0 * kbin
        else:                                   # class logits: scale, then bias by quantile
            s = _round_sig(cls_std / max(float((y - b.view(1, -1, 1, 1)).std()), 1e-12))
            w.mul_(s)
            z = F.conv2d(xin, w, None, stride=1, padding=0)
            mx = z.amax(1).reshape(-1)
            q = float(torch.quantile(mx, 1.0 - target_frac))
            b.fill_(round(-q, 2))
        out[name + ".weight"] = w.numpy()
        out[name + ".bias"] = b.numpy()
        return orig_conv(name, xin)

    net.conv = conv
    net.head_maps(net.features(x))
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


def synthetic_model(scale="n", nc=4, seed=0):
    """The repo's standard synthetic checkpoint: seeded init + calibration."""
    from oracle import zlw
    return calibrate(zlw.make_weights(scale, nc, seed=seed, gain=1.0, cls_bias=0.0), scale, nc)
