import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200
H, W, C = 32, 16, int(sys.argv[1]) if len(sys.argv) > 1 else 32
COUT = int(sys.argv[2]) if len(sys.argv) > 2 else 64
x = np.zeros((1, H, W, C), np.float32)
hh, ww = np.mgrid[0:H, 0:W]
x[0, :, :, 0] = hh * 32 + ww + 1          # +1 so that zero == padding
for (r, s) in [(1, 1), (0, 0), (2, 2), (0, 1), (1, 0), (1, 2), (2, 1)]:
    wt = np.zeros((COUT, 3, 3, C), np.float32)
    wt[0, r, s, 0] = 1.0
    b = np.zeros(COUT, np.float32)
    try:
        y = zlb200.test_conv(x, wt, b, stride=2, act=False, impl=3, out_f32=True, fp16=(os.environ.get("F16","1")=="1"))
    except Exception as e:
        print("tap", r, s, "ERROR", e); break
    got = y[0, :, :, 0]
    exp = np.zeros_like(got)
    for oy in range(H // 2):
        for ox in range(W // 2):
            iy, ix = 2 * oy - 1 + r, 2 * ox - 1 + s
            if 0 <= iy < H and 0 <= ix < W:
                exp[oy, ox] = x[0, iy, ix, 0]
    ok = np.array_equal(got, exp)
    print("tap", (r, s), "match", ok)
    if not ok:
        for oy in range(0, 4):
            row = []
            for ox in range(0, 8):
                v = int(got[oy, ox]) - 1
                row.append(f"({v // 32},{v % 32})" if v >= 0 else "pad")
            print("  oy", oy, " got:", " ".join(row))
            row = []
            for ox in range(0, 8):
                v = int(exp[oy, ox]) - 1
                row.append(f"({v // 32},{v % 32})" if v >= 0 else "pad")
            print("       exp:", " ".join(row))
