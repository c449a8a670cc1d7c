"""Serialised per-step kernel time (zl_engine_profile) for A/B runs against another build: ZL_B200_LIB=<path>."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200  # noqa: E402
from oracle import synth, yolov8_ref, zlw  # noqa: E402

if __name__ == "__main__":
    for scale, batch in (("n", 64), ("s", 256), ("m", 256)):
        t = yolov8_ref.synthetic_model(scale, 80, 0)
        e = zlb200.Engine(640, 640, 80, scale, precision=zlb200.FP16, max_batch=batch, num_lanes=2)
        e.load_weights_blob(zlw.dumps(t, scale, 80))
        frames = list(synth.frames_structured(64, 640, 640))
        for s in range(2):
            e.upload_resident(s, [frames[i % 64] for i in range(batch)])
        prof = e.profile(0, 3)
        conv = sum(p["ms"] for p in prof if p["kind"] in (1, 9))
        e.run_resident(2, 4)
        ms, _, _ = e.run_resident(2, 12)
        print(f"{os.environ.get('ZL_B200_LIB', 'HEAD')[-24:]:24s} {scale} b={batch}: conv kernels {conv:8.3f} ms, all {sum(p['ms'] for p in prof):8.3f} ms, resident {batch * 12 / ms * 1e3:9.0f} frames/s", flush=True)
        e.close()
