#!/usr/bin/env python3
"""Mints tests/golden/*.npz from oracle/_ref — the reference's OWN preProcess / postProcess / applyNMS / calculateIoU
text compiled with strict IEEE flags (oracle/ref/shim.cpp).  Run in the container that has /root/reference:

    python oracle/mint_golden.py

Inputs are rebuilt from oracle/synth.py's version-independent generators (golden_bytes / golden_head), so a fixture
stores only the case parameters and the reference's OUTPUT (in full for the small cases, as count + SHA-256 of the
canonical bytes for the full-size ones)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_c, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (src_w, src_h, model_w, model_h, seed, store_full)
PRE_CASES = [
    (16, 12, 32, 32, 1, True), (37, 53, 64, 64, 2, True), (1, 1, 32, 32, 3, True), (33, 31, 32, 32, 4, True),
    (416, 416, 416, 416, 5, False), (800, 600, 416, 416, 6, False), (1920, 1080, 640, 640, 7, False),
    (37, 53, 416, 416, 8, False), (415, 417, 416, 416, 9, False), (640, 360, 640, 640, 10, False), (640, 640, 640, 640, 11, False),
]
# (name, nc, A, seed, img_w, img_h, conf, iou, ties, clusters, store_full)
POST_CASES = [
    ("cfg1_defaults", 4, 3549, 21, 416, 416, 0.5, 0.45, False, 0, True),
    ("cfg1_low_conf", 4, 3549, 22, 800, 600, 0.05, 0.45, False, 0, True),
    ("nc80_small", 80, 336, 23, 640, 640, 0.01, 0.45, False, 0, True),
    ("clusters_no_ties", 4, 2100, 24, 640, 640, 0.001, 0.45, False, 24, True),
    ("single_class", 1, 1000, 25, 320, 320, 0.01, 0.3, False, 8, True),
    ("nothing_passes", 4, 500, 26, 416, 416, 1.5, 0.45, False, 0, True),
    ("cfg5_full", 80, 8400, 27, 640, 640, 0.01, 0.45, False, 0, False),
    ("cfg5_clusters", 80, 8400, 28, 640, 640, 0.01, 0.45, False, 64, False),
    ("cfg5_iou_07", 80, 8400, 29, 1920, 1080, 0.01, 0.7, False, 64, False),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    assert ref_c.lib().zlr_sizeof_detection() == 40
    manifest = {"minted_from": "oracle/_ref/libzl_ref.so (reference text src/inference/onnx_engine.cpp:649-700,758-834,837-878,881-909; strict IEEE build)",
                "preprocess": [], "postprocess": [], "iou": []}
    arrays = {}
    for (w, h, mw, mh, seed, full) in PRE_CASES:
        img = synth.golden_bytes((h, w, 3), seed)
        code, out = ref_c.preprocess(img, w, h, mw, mh)
        assert code == 0
        rec = dict(w=w, h=h, mw=mw, mh=mh, seed=seed, sha256=sha(out), full=full)
        if full:
            arrays[f"pre_{seed}"] = out
        manifest["preprocess"].append(rec)
    code, _ = ref_c.preprocess(np.zeros(10, np.uint8), 4, 4, 8, 8)
    manifest["preprocess_wrong_length_code"] = int(code)
    for (name, nc, A, seed, iw, ih, conf, iou, ties, clusters, full) in POST_CASES:
        raw = synth.golden_head(nc, A, seed, img=max(iw, ih), ties=ties, clusters=clusters)
        det = ref_c.postprocess(raw, iw, ih, conf, iou)
        can = ref_c.canonical(det)
        rec = dict(name=name, nc=nc, A=A, seed=seed, img_w=iw, img_h=ih, conf=conf, iou=iou, ties=ties, clusters=clusters,
                   count=int(len(det)), sha256_in_order=sha(det), sha256_canonical=sha(can), full=full, raw_sha256=sha(raw))
        if full:
            arrays[f"post_{name}"] = det
        manifest["postprocess"].append(rec)
        print(f"{name:18s} kept {len(det)}")
    # calculateIoU on a fixed grid of box pairs (centre format)
    b = synth.golden_unit((512, 8), 31)
    b[:, 2:4] = b[:, 2:4] * np.float32(0.5)
    b[:, 6:8] = b[:, 6:8] * np.float32(0.5)
    b[:8, 2:4] = 0
    b[:4, 6:8] = 0                                      # union == 0 guard and degenerate boxes
    b[8:16, 4:8] = b[8:16, 0:4]                         # identical boxes
    ious = np.array([ref_c.iou(r[:4], r[4:]) for r in b], np.float32)
    arrays["iou_boxes"], arrays["iou_values"] = b, ious
    np.savez_compressed(os.path.join(OUT, "ref_vectors.npz"), **arrays)
    json.dump(manifest, open(os.path.join(OUT, "ref_manifest.json"), "w"), indent=1)
    print("wrote", os.path.join(OUT, "ref_vectors.npz"), os.path.getsize(os.path.join(OUT, "ref_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
