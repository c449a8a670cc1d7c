"""The CUDA path (through the C-ABI) against the golden fixtures minted from the reference's own compiled text
(tests/golden/, oracle/mint_golden.py): zl_preprocess for P1, zl_decode_nms for F1 + N1.  Bit-exact; inside an exact
(class, confidence) tie group the reference's order is unspecified (std::sort), so lists are compared canonically."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAN = json.load(open(os.path.join(GOLD, "ref_manifest.json")))
VEC = np.load(os.path.join(GOLD, "ref_vectors.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def canonical(d):
    order = np.lexsort((d["h"], d["w"], d["y"], d["x"], -d["confidence"].astype(np.float64), d["class_id"]))
    return d[order]


@pytest.mark.parametrize("rec", MAN["preprocess"], ids=lambda r: f"{r['w']}x{r['h']}to{r['mw']}x{r['mh']}")
def test_cuda_preprocess_matches_reference_golden(built_lib, rec):
    import zlb200
    e = zlb200.Engine(rec["mw"], rec["mh"], 4, "n", precision=zlb200.FP32, max_batch=1, max_frame=(max(rec["w"], rec["mw"]), max(rec["h"], rec["mh"])))
    img = synth.golden_bytes((rec["h"], rec["w"], 3), rec["seed"])
    out = e.preprocess(img, rec["w"], rec["h"])
    assert sha(out) == rec["sha256"]
    if rec["full"]:
        assert np.array_equal(out.view(np.uint32), VEC[f"pre_{rec['seed']}"].view(np.uint32))
    e.close()


@pytest.mark.parametrize("rec", MAN["postprocess"], ids=lambda r: r["name"])
def test_cuda_decode_nms_matches_reference_golden(built_lib, rec):
    import zlb200
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP32, max_batch=1)
    raw = synth.golden_head(rec["nc"], rec["A"], rec["seed"], img=max(rec["img_w"], rec["img_h"]), ties=rec["ties"], clusters=rec["clusters"])
    assert sha(raw) == rec["raw_sha256"], "input generator drifted"
    det = e.decode_nms(raw[None], rec["img_w"], rec["img_h"], rec["conf"], rec["iou"])[0]
    assert len(det) == rec["count"]
    assert sha(canonical(det)) == rec["sha256_canonical"]
    key = det["class_id"].astype(np.float64) * 4 - det["confidence"].astype(np.float64)
    assert np.all(np.diff(key) >= 0)
    if rec["full"]:
        assert np.array_equal(canonical(det).view(np.uint8), canonical(VEC[f"post_{rec['name']}"]).view(np.uint8))
    e.close()


def test_cuda_decode_nms_golden_cases_as_one_batch(built_lib):
    """The three full-size cfg5 golden cases in ONE launch with different request-frame sizes per frame."""
    import zlb200
    recs = [r for r in MAN["postprocess"] if r["A"] == 8400 and r["nc"] == 80 and r["iou"] == 0.45]
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP32, max_batch=1)
    raws = np.stack([synth.golden_head(80, 8400, r["seed"], img=max(r["img_w"], r["img_h"]), clusters=r["clusters"]) for r in recs])
    dets = e.decode_nms(raws, [r["img_w"] for r in recs], [r["img_h"] for r in recs], 0.01, 0.45)
    for d, r in zip(dets, recs):
        assert len(d) == r["count"] and sha(canonical(d)) == r["sha256_canonical"]
    e.close()
