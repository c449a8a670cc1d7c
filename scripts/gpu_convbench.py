"""Micro-benchmark of single conv layers through the engine's profile path (used under ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200  # noqa: E402
from oracle import synth, yolov8_ref, zlw  # noqa: E402

if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    t = yolov8_ref.synthetic_model("n", 80, 0)
    e = zlb200.Engine(640, 640, 80, "n", precision=zlb200.FP16, max_batch=64)
    e.load_weights_blob(zlw.dumps(t, "n", 80))
    e.upload_resident(0, list(synth.frames_structured(64, 640, 640)))
    prof = e.profile(0, iters)
    print(sum(p["ms"] for p in prof))
    e.close()
