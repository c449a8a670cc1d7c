"""NMS timing against candidate statistics on the bench's frames (not a test)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200
from oracle import synth, yolov8_ref, zlw, oracle_c
t = yolov8_ref.synthetic_model("n", 80, 0)
e = zlb200.Engine(640, 640, 80, "n", precision=zlb200.FP16, max_batch=64)
e.load_weights_blob(zlw.dumps(t, "n", 80))
frames = list(synth.frames_structured(64, 640, 640, seed=5678))
raw = e.forward_raw(frames)
stats = []
for f in range(64):
    c, _ = oracle_c.decode_filter(raw[f], 640, 640, 0.5)
    k, _ = oracle_c.postprocess(raw[f], 640, 640, 0.5, 0.45)
    bc = np.bincount(c["class_id"], minlength=80)
    kb = np.bincount(k["class_id"], minlength=80)
    stats.append((len(c), len(k), int(bc.max()), int((bc > 32).sum()), int(kb[bc > 32].sum())))
s = np.array(stats)
print("cand   min/med/max", s[:, 0].min(), np.median(s[:, 0]), s[:, 0].max())
print("kept   min/med/max", s[:, 1].min(), np.median(s[:, 1]), s[:, 1].max())
print("maxseg min/med/max", s[:, 2].min(), np.median(s[:, 2]), s[:, 2].max())
print("large segments per frame med/max", np.median(s[:, 3]), s[:, 3].max(), " kept inside large segments med/max", np.median(s[:, 4]), s[:, 4].max())
order = np.argsort(-s[:, 0])
for f in list(order[:4]) + list(order[-2:]):
    mf, mn, kept = e.bench_decode_nms(raw[f:f + 1], 0.5, 0.45, iters=5)
    print(f"frame {f}: cand {s[f,0]} kept {s[f,1]} maxseg {s[f,2]} nlarge {s[f,3]} kept_in_large {s[f,4]}: filter {mf*1e3:.1f} us, nms {mn*1e3:.1f} us")
for n in (1, 16, 64):
    mf, mn, kept = e.bench_decode_nms(raw[:n], 0.5, 0.45, iters=5)
    print(f"n={n}: filter {mf*1e3:.1f} us, nms {mn*1e3:.1f} us, kept {kept}")
