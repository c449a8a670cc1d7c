"""GPU diagnostics: bf16-vs-oracle statistics and the per-kernel profile table (not a test)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import zlb200  # noqa: E402
from oracle import oracle_c, synth, yolov8_ref, zlw  # noqa: E402
from test_gpu_engine import box_iou, oracle_pipeline  # noqa: E402


def bf16_stats(scale, nc, hw, nframes, precision=None):
    t = yolov8_ref.synthetic_model(scale, nc, 0)
    frames = list(synth.frames_structured(nframes, hw, hw, seed=5678))
    e = zlb200.Engine(hw, hw, nc, scale, precision=zlb200.BF16 if precision is None else precision, max_batch=nframes)
    e.load_weights_blob(zlw.dumps(t, scale, nc))
    e.warmup(1)
    raw_ref, det_ref = oracle_pipeline(t, scale, nc, frames, hw, hw)
    raw = e.forward_raw(frames)
    dets = e.infer(frames)
    ds = np.abs(raw[:, 4:] - raw_ref[:, 4:])
    db = np.abs(raw[:, :4] - raw_ref[:, :4])
    out = {"cfg": f"{scale} nc{nc} {hw} prec{precision}", "score_max": float(ds.max()), "score_med": float(np.median(ds)), "score_p999": float(np.quantile(ds, 0.999)),
           "box_max_px": float(db.max()), "box_med_px": float(np.median(db)), "kept_ref": [len(d) for d in det_ref], "kept": [len(d) for d in dets]}
    ious = []
    for d, r in zip(dets, det_ref):
        for i in range(len(r)):
            same = d[d["class_id"] == r["class_id"][i]]
            ious.append(float(box_iou(same, r[i]).max()) if len(same) else 0.0)
    ious = np.array(ious)
    out["n_oracle_dets"] = int(sum(len(r) for r in det_ref))
    out["n_below_0.99"] = int(out["n_oracle_dets"] - (ious >= 0.99).sum())      # north_star: every matched detection IoU >= 0.99
    out["kept_equal"] = bool(all(len(d) == len(r) for d, r in zip(dets, det_ref)))
    out["iou_hist_lt0.9_0.9to0.99_ge0.99"] = [int((ious < 0.9).sum()), int(((ious >= 0.9) & (ious < 0.99)).sum()), int((ious >= 0.99).sum())]
    out["matched_ge_0.5"] = int((ious >= 0.5).sum())
    m = ious[ious >= 0.5]
    out["iou_min"] = float(m.min()); out["iou_p01"] = float(np.quantile(m, 0.01)); out["iou_med"] = float(np.median(m))
    out["frac_iou_ge_0.99"] = float((m >= 0.99).mean())
    e.close()
    return out


def profile_table():
    t = yolov8_ref.synthetic_model("n", 80, 0)
    e = zlb200.Engine(640, 640, 80, "n", precision=zlb200.BF16, max_batch=64)
    e.load_weights_blob(zlw.dumps(t, "n", 80))
    e.warmup(1)
    e.upload_resident(0, list(synth.frames_structured(64, 640, 640)))
    prof = e.profile(0, 5)
    e.close()
    return prof


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    res = {"bf16": [bf16_stats("n", 4, 416, 4), bf16_stats("n", 80, 640, 2),
                    bf16_stats("n", 4, 416, 4, zlb200.FP16), bf16_stats("n", 80, 640, 2, zlb200.FP16)],
           "profile": profile_table() if "--profile" in sys.argv else []}
    json.dump(res, open("gpurun_out/diag.json", "w"), indent=1)
    for b in res["bf16"]:
        print(b)
    tot = sum(p["ms"] for p in res["profile"])
    print(f"profile total {tot:.3f} ms")
    for p in res["profile"]:
        tf = p["flops"] / (p["ms"] * 1e-3) / 1e12 if p["ms"] > 0 else 0
        gb = p["bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else 0
        print(f"{p['name']:28s} k{p['kind']} {p['ms']*1e3:8.1f} us  {tf:7.1f} TF/s  {gb:7.1f} GB/s  {100*p['ms']/tot:5.1f}%")
