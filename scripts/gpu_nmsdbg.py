import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/zero-latency-yolo_b200/python")
import zlb200
from oracle import synth, yolov8_ref, zlw
t = yolov8_ref.synthetic_model("n", 80, 0)
e = zlb200.Engine(640, 640, 80, "n", precision=zlb200.FP16, max_batch=64)
e.load_weights_blob(zlw.dumps(t, "n", 80))
frames = list(synth.frames_structured(64, 640, 640, seed=5678))
raw = e.forward_raw(frames)
for f in (61, 0):
    print("frame", f, e.bench_decode_nms(raw[f:f + 1], 0.5, 0.45, iters=3), flush=True)
