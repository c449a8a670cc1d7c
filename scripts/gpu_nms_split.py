"""NMS time vs batch size on the bench frames (not a test); run once per ZL_NMS_SPLIT setting (read once per process)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200
from oracle import synth, yolov8_ref, zlw
t = yolov8_ref.synthetic_model("n", 80, 0)
e = zlb200.Engine(640, 640, 80, "n", precision=zlb200.FP16, max_batch=64)
e.load_weights_blob(zlw.dumps(t, "n", 80))
frames = list(synth.frames_structured(64, 640, 640, seed=5678))
raw = e.forward_raw(frames)
print("ZL_NMS_SPLIT =", os.environ.get("ZL_NMS_SPLIT", "auto"))
for lo, n in ((61, 1), (0, 1), (0, 8), (0, 16), (0, 32), (0, 64)):
    mf, mn, kept = e.bench_decode_nms(raw[lo:lo + n], 0.5, 0.45, iters=20)
    print(f"frames {lo}..{lo + n - 1}: filter {mf * 1e3:.1f} us, nms {mn * 1e3:.1f} us, kept {kept}", flush=True)
cfg5 = synth.stress_head(128, 80, 8400, seed=42)
for n in (16, 32, 64, 128):
    mf, mn, kept = e.bench_decode_nms(cfg5[:n], 0.01, 0.45, iters=5)
    print(f"cfg5 n={n}: filter {mf * 1e3:.1f} us, nms {mn * 1e3:.1f} us, kept {kept}", flush=True)
