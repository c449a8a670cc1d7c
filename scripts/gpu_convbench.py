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
    scale = sys.argv[2] if len(sys.argv) > 2 else "n"
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    t = yolov8_ref.synthetic_model(scale, 80, 0)
    e = zlb200.Engine(640, 640, 80, scale, precision=zlb200.FP16, max_batch=batch)
    e.load_weights_blob(zlw.dumps(t, scale, 80))
    frames = list(synth.frames_structured(min(batch, 64), 640, 640))
    e.upload_resident(0, [frames[i % len(frames)] for i in range(batch)])
    prof = e.profile(0, iters)
    print(sum(p["ms"] for p in prof))
    e.close()
