import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
import zlb200
from oracle import synth, yolov8_ref, zlw
t4 = yolov8_ref.synthetic_model("n", 4, seed=0)
e1 = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=1)
e1.load_weights_blob(zlw.dumps(t4, "n", 4)); e1.warmup(3)
pf = zlb200.pinned_array((416, 416, 3)); pf[:] = synth.frames_structured(1, 416, 416)[0]
lat = np.sort(e1.bench_latency(pf, warmup=200, iters=2000))
print("p50 %.4f ms p99 %.4f ms device %.4f ms" % (lat[len(lat)//2], lat[int(len(lat)*0.99)], e1.stats()["avg_device_time_ms"]))
e1.upload_resident(0, [np.asarray(pf)])
prof = e1.profile(0, 20)
tot = sum(p["ms"] for p in prof)
kinds = {}
for p in prof: kinds[p["kind"]] = kinds.get(p["kind"], 0) + p["ms"]
print("profile total %.3f ms; by kind:" % tot, {k: round(v, 3) for k, v in kinds.items()}, "launches", len(prof))
for p in sorted(prof, key=lambda p: -p["ms"])[:8]: print("  ", p["name"], p["kind"], round(p["ms"]*1e3, 1), "us")
