import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200, torch
which = sys.argv[1]
H, W, C, COUT = 32, 16, 32, 64
rng = np.random.default_rng(0)
def bf(a): return torch.tensor(a).to(torch.bfloat16).to(torch.float32).numpy()
x = bf(rng.normal(size=(1, H, W, C)).astype(np.float32))
wt = bf((rng.normal(size=(COUT, 3, 3, C)) / 17).astype(np.float32))
b = np.zeros(COUT, np.float32)
if which == "rand_act":   y = zlb200.test_conv(x, wt, b, stride=2, act=True, impl=3, out_f32=True)
if which == "rand_noact": y = zlb200.test_conv(x, wt, b, stride=2, act=False, impl=3, out_f32=True)
if which == "zero_w":     y = zlb200.test_conv(x, wt * 0, b, stride=2, act=False, impl=3, out_f32=True)
if which == "zero_x":     y = zlb200.test_conv(x * 0, wt, b, stride=2, act=False, impl=3, out_f32=True)
if which == "s1":         y = zlb200.test_conv(x, wt, b, stride=1, act=False, impl=3, out_f32=True)
xt = torch.tensor(x).permute(0, 3, 1, 2); wtt = torch.tensor(wt).permute(0, 3, 1, 2)
st = 1 if which == "s1" else 2
ref = torch.nn.functional.conv2d(xt, wtt if which != "zero_w" else wtt * 0, None, stride=st, padding=1)
if which == "rand_act": ref = torch.nn.functional.silu(ref)
if which == "zero_x": ref = ref * 0
ref = ref.permute(0, 2, 3, 1).numpy()
err = np.abs(y - ref)
print(which, "ok, max err", err.max(), "at", np.unravel_index(err.argmax(), err.shape), "ref absmax", np.abs(ref).max())
if err.max() > 0.05:
    bad = (err > 0.05)
    print(" bad fraction", bad.mean(), " bad per out-row", bad.any(axis=(0, 2, 3)).astype(int), " per out-col", bad.any(axis=(0, 1, 3)).astype(int))
