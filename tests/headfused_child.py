"""Child process of tests/test_gpu_headfused.py: runs the cases with whatever ZL_FUSE_HEAD the parent set and writes the
detections (raw bytes per frame) plus the hot path's op names to an .npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))

CASES = [
    # name, scale, nc, size, frames, precision, conf, weighted
    ("n80_fp16", "n", 80, 640, 5, "fp16", 0.25, False),
    ("n80_bf16", "n", 80, 640, 3, "bf16", 0.25, False),
    ("n80_fp16_lowconf", "n", 80, 640, 2, "fp16", 0.02, False),
    ("n80_fp16_weighted", "n", 80, 640, 2, "fp16", 0.2, True),
    ("n4_416_b1", "n", 4, 416, 1, "fp16", 0.25, False),
    ("n4_416_b7", "n", 4, 416, 7, "bf16", 0.25, False),
    ("s80_fp16", "s", 80, 640, 2, "fp16", 0.25, False),
    ("n3_odd", "n", 3, 416, 3, "fp16", 0.25, False),
    # class counts that take the rolled scan loops (neither 16 nor 80 padded classes): 2, 7 and 12 chunks of 16
    ("n20_loop", "n", 20, 320, 2, "fp16", 0.25, False),
    ("n100_loop", "n", 100, 320, 2, "bf16", 0.25, False),
    ("n192_max", "n", 192, 320, 2, "fp16", 0.2, False),
]


def run_case(case):
    import zlb200
    from oracle import synth, yolov8_ref, zlw
    name, scale, nc, size, n, prec, conf, weighted = case
    t = yolov8_ref.synthetic_model(scale, nc, 0)
    blob = zlw.dumps(t, scale, nc)
    frames = list(synth.frames_structured(n, size, size, seed=900 + n))
    cw = None
    if weighted:
        cw = np.random.default_rng(5).uniform(0.3, 1.4, nc).astype(np.float32)
    e = zlb200.Engine(size, size, nc, scale, precision=zlb200.FP16 if prec == "fp16" else zlb200.BF16, conf=conf, iou=0.45,
                      max_batch=n, class_weights=cw)
    e.load_weights_blob(blob)
    e.warmup(1)
    dets = e.infer(frames)
    e.upload_resident(0, frames)
    ops = [o["name"] for o in e.profile(0, 1)]
    e.close()
    return dets, ops


def main(out_path):
    out = {}
    for case in CASES:
        dets, ops = run_case(case)
        for i, d in enumerate(dets):
            out[f"{case[0]}/{i}"] = np.frombuffer(d.tobytes(), np.uint8)
        out[f"{case[0]}/ops"] = np.array(ops)
    np.savez(out_path, **out)


if __name__ == "__main__":
    main(sys.argv[1])
