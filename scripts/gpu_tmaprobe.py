import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200
H, W, C = 32, 32, 16
x = np.zeros((1, H, W, C), np.float16)
hh, ww = np.mgrid[0:H, 0:W]
for c in range(C):
    x[0, :, :, c] = hh * 32 + ww + 1
def show(tag, box, es, sw, coords, expect):
    try:
        done, d = zlb200.probe_tma(x, box, es, sw, coords, expect, 8192)
    except Exception as e:
        print(tag, "ERROR", e); return
    rows = d.reshape(-1, 16)[:, 0]            # first channel of each 32-B row (swizzle none/32 keeps ch0 in chunk 0 or 1)
    vals = []
    for v in d.reshape(-1, 16)[:40]:
        v0 = int(v[0]) if v[0] == v[0] else -9
        vals.append("pad" if v0 == 0 else ("EE" if v0 < 0 or v0 > 2000 else f"{(v0-1)//32},{(v0-1)%32}"))
    print(tag, "completed", done, "box", box, "es", es, "coords", coords, "expect", expect)
    print("   rows:", " | ".join(vals))
# no stride baseline
show("base ", (16, 4, 3), 1, 0, (0, 2, 5, 0), 16 * 4 * 3 * 2)
# stride 2, box = traversed extent, expect ceil(box/2) elements
show("s2 a ", (16, 8, 6), 2, 0, (0, 2, 5, 0), 16 * 4 * 3 * 2)
show("s2 b ", (16, 8, 6), 2, 0, (0, 2, 5, 0), 16 * 8 * 6 * 2)
show("s2 c ", (16, 7, 5), 2, 0, (0, 3, 4, 0), 16 * 4 * 3 * 2)
show("s2 neg", (16, 8, 6), 2, 0, (0, -2, -1, 0), 16 * 4 * 3 * 2)
show("s2 odd", (16, 8, 6), 2, 0, (0, -1, -1, 0), 16 * 4 * 3 * 2)
