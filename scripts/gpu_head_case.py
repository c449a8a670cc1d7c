"""Two resident steps of the b=64 YOLOv8n pipeline (the fused head kernel is the profiling target: ncu -k regex:head_decode)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200  # noqa: E402
from oracle import synth, yolov8_ref, zlw  # noqa: E402

if __name__ == "__main__":
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    scale = sys.argv[2] if len(sys.argv) > 2 else "n"
    t = yolov8_ref.synthetic_model(scale, 80, 0)
    e = zlb200.Engine(640, 640, 80, scale, precision=zlb200.FP16, max_batch=batch, use_graph=0)
    e.load_weights_blob(zlw.dumps(t, scale, 80))
    frames = list(synth.frames_structured(batch, 640, 640))
    e.upload_resident(0, frames)
    ms, launches, dets = e.run_resident(1, 2)
    print("ms/step", ms / 2, "launches", launches, "dets", dets)
    e.close()
