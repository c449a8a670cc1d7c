"""NMS phase stamps (ZL_NMS_DEBUG=0) of the b=1 416x416 nc=4 latency configuration, un-captured."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200  # noqa: E402
from oracle import synth, yolov8_ref, zlw  # noqa: E402

if __name__ == "__main__":
    t = yolov8_ref.synthetic_model("n", 4, 0)
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=1, use_graph=0)
    e.load_weights_blob(zlw.dumps(t, "n", 4))
    frames = [synth.frames_structured(1, 416, 416)[0]] * 4       # the frame bench.py / gpu_latency.py time
    for f in frames:
        d = e.infer([f])
        print("dets", len(d[0]), flush=True)
    e.close()
