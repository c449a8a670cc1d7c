"""The CPU oracle against the golden fixtures minted from the reference's own compiled text (tests/golden/, made by
oracle/mint_golden.py from oracle/_ref).  Needs neither /root/reference nor oracle/_ref at run time."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle_c, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAN = json.load(open(os.path.join(GOLD, "ref_manifest.json")))
VEC = np.load(os.path.join(GOLD, "ref_vectors.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def canonical(d):
    order = np.lexsort((d["h"], d["w"], d["y"], d["x"], -d["confidence"].astype(np.float64), d["class_id"]))
    return d[order]


@pytest.mark.parametrize("rec", MAN["preprocess"], ids=lambda r: f"{r['w']}x{r['h']}to{r['mw']}x{r['mh']}")
def test_oracle_preprocess_matches_golden(rec):
    img = synth.golden_bytes((rec["h"], rec["w"], 3), rec["seed"])
    code, out = oracle_c.preprocess(img, rec["w"], rec["h"], rec["mw"], rec["mh"])
    assert code == 0 and sha(out) == rec["sha256"]
    if rec["full"]:
        assert np.array_equal(out.view(np.uint32), VEC[f"pre_{rec['seed']}"].view(np.uint32))


def test_oracle_wrong_length_code_matches_golden():
    assert oracle_c.preprocess(np.zeros(10, np.uint8), 4, 4, 8, 8)[0] == MAN["preprocess_wrong_length_code"] == 203


@pytest.mark.parametrize("rec", MAN["postprocess"], ids=lambda r: r["name"])
def test_oracle_postprocess_matches_golden(rec):
    raw = synth.golden_head(rec["nc"], rec["A"], rec["seed"], img=max(rec["img_w"], rec["img_h"]), ties=rec["ties"], clusters=rec["clusters"])
    assert sha(raw) == rec["raw_sha256"], "input generator drifted"
    det, _ = oracle_c.postprocess(raw, rec["img_w"], rec["img_h"], rec["conf"], rec["iou"])
    assert len(det) == rec["count"]
    # order inside an exact (class, confidence) tie group is unspecified in the reference (std::sort): compare canonically
    assert sha(canonical(det)) == rec["sha256_canonical"]
    key = det["class_id"].astype(np.float64) * 4 - det["confidence"].astype(np.float64)
    assert np.all(np.diff(key) >= 0), "output must be sorted by (class asc, confidence desc)"
    if rec["full"]:
        assert np.array_equal(canonical(det).view(np.uint8), canonical(VEC[f"post_{rec['name']}"]).view(np.uint8))


def test_oracle_iou_matches_golden():
    b, want = VEC["iou_boxes"], VEC["iou_values"]
    got = np.array([oracle_c.iou(tuple(r[:4]), tuple(r[4:])) for r in b], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
