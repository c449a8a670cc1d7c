"""NMS phase stamps of one frame inside the bench batch (not a test).  ZL_NMS_DEBUG=<frame> ZL_NMS_SPLIT=<1|2|4|8> python scripts/gpu_nmsdbg.py [batch]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200
from oracle import synth, yolov8_ref, zlw
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t = yolov8_ref.synthetic_model("n", 80, 0)
e = zlb200.Engine(640, 640, 80, "n", precision=zlb200.FP16, max_batch=64)
e.load_weights_blob(zlw.dumps(t, "n", 80))
frames = list(synth.frames_structured(64, 640, 640, seed=5678))
raw = e.forward_raw(frames)
lo = int(os.environ.get("ZL_NMS_DEBUG", "0")) if nb < 64 else 0
print("batch", nb, e.bench_decode_nms(raw[lo:lo + nb], 0.5, 0.45, iters=3), flush=True)
