"""Runs selected persistent-conv unit cases with the launch plan printed (ZL_PLAN_DEBUG=1)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zlb200
from test_gpu_kernels import HALO_CASES, _torch_conv, _bf
import torch
sel = [int(a) for a in sys.argv[1:]] or range(len(HALO_CASES))
for i in sel:
    n, h, w, cin, cout, use_res, fp16, out_f32, ctas = HALO_CASES[i]
    rng = np.random.default_rng(i)
    q = (lambda a: a.astype(np.float16).astype(np.float32)) if fp16 else _bf
    x = q(rng.normal(size=(n, h, w, cin)).astype(np.float32))
    wt = q((rng.normal(size=(cout, 3, 3, cin)) / np.sqrt(cin * 9)).astype(np.float32))
    b = rng.normal(size=cout).astype(np.float32)
    res = q(rng.normal(size=(n, h, w, cout)).astype(np.float32)) if use_res else None
    ref = _torch_conv(x, wt, b, 1, True, res)
    try:
        y = zlb200.test_conv(x, wt, b, stride=1, act=True, res=res, impl=3, out_f32=out_f32, ntile_hint=ctas, fp16=fp16)
        print(i, HALO_CASES[i], "max err", float(np.abs(y - ref).max()), flush=True)
    except Exception as e:
        print(i, HALO_CASES[i], "FAILED", e, flush=True)
        break
