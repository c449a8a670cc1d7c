"""Ad-hoc sweep: unusual model sizes / class counts / batch sizes through the 16-bit engine against the oracle contract."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zlb200  # noqa: E402
from oracle import synth  # noqa: E402
from conftest import synthetic_model  # noqa: E402
import test_gpu_engine as T  # noqa: E402

CASES = [("n", 5, 352, 608, 3), ("n", 1, 32, 32, 1), ("n", 2, 64, 96, 2), ("n", 80, 1280, 1280, 2), ("s", 7, 224, 416, 5),
         ("n", 33, 96, 96, 9), ("m", 4, 160, 160, 1), ("n", 80, 640, 384, 17)]
for scale, nc, w, h, n in CASES:
    try:
        tensors, blob = synthetic_model(scale, nc)
        frames = list(synth.frames_structured(n, h, w, seed=7))
        for prec in (zlb200.FP16, zlb200.BF16):
            e = zlb200.Engine(w, h, nc, scale, precision=prec, max_batch=n)
            e.load_weights_blob(blob)
            e.warmup(1)
            T._check_16bit(e, tensors, scale, nc, frames, w, h, "fp16" if prec == zlb200.FP16 else "bf16", strict=False)
            e.close()
        print("OK  ", scale, nc, w, h, n, flush=True)
    except Exception as ex:
        print("FAIL", scale, nc, w, h, n, repr(ex)[:300], flush=True)
        traceback.print_exc()
