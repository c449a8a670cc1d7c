import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200
from oracle import synth, yolov8_ref, zlw, oracle_c
t = yolov8_ref.synthetic_model("n", 80, 0)
e = zlb200.Engine(640, 640, 80, "n", precision=zlb200.FP16, max_batch=16)
e.load_weights_blob(zlw.dumps(t, "n", 80))
frames = list(synth.frames_structured(16, 640, 640))
raw = e.forward_raw(frames)
for f in range(4):
    c, _ = oracle_c.decode_filter(raw[f], 640, 640, 0.5)
    k, _ = oracle_c.postprocess(raw[f], 640, 640, 0.5, 0.45)
    bc = np.bincount(c["class_id"], minlength=80)
    print("frame", f, "cand", len(c), "kept", len(k), "max class seg", bc.max(), "classes present", (bc > 0).sum(), "kept in biggest", (k["class_id"] == bc.argmax()).sum())
for n in (1, 4, 16):
    mf, mn, kept = e.bench_decode_nms(raw[:n], 0.5, 0.45, iters=5)
    print(f"n={n}: filter {mf*1e3:.1f} us, nms {mn*1e3:.1f} us, kept {kept}")
st = synth.stress_head(16, 80, 8400, seed=42)
mf, mn, kept = e.bench_decode_nms(st, 0.01, 0.45, iters=3)
print(f"stress n=16: filter {mf*1e3:.1f} us, nms {mn*1e3:.1f} us, kept {kept}")
