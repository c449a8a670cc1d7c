"""Known-answer tests for the CPU oracle, derived from the reference SOURCE TEXT
(src/inference/onnx_engine.cpp:649-909).  The reference ships no tests or golden
vectors, so these pin the oracle to the behaviours the source spells out."""
import numpy as np
import pytest

from oracle import oracle_c
from oracle.oracle_c import DET_DTYPE


def _py_stretch(src, dst):
    scale = np.float32(src) / np.float32(dst)
    return np.array([min(int(np.float32(i) * scale), src - 1) for i in range(dst)], np.int32)


@pytest.mark.parametrize("src,dst", [(416, 416), (800, 416), (600, 416), (1920, 640), (1080, 640), (37, 416), (53, 416), (1, 32)])
def test_stretch_index_table(src, dst):
    # onnx_engine.cpp:673-682: scale = float(src)/dst ; idx = min(int(i*scale), src-1)
    got = oracle_c.stretch_index(src, dst)
    assert np.array_equal(got, _py_stretch(src, dst))
    assert got[0] == 0 and got.max() <= src - 1
    if src == dst:
        assert np.array_equal(got, np.arange(dst))


def test_preprocess_bgr_to_rgb_and_normalise():
    # one 2x2 BGR image, identity size: channel c of the output reads byte 2-c (onnx_engine.cpp:685), /255.0f (:693)
    img = np.array([[[10, 20, 30], [40, 50, 60]], [[70, 80, 90], [255, 0, 128]]], np.uint8)
    code, out = oracle_c.preprocess(img, 2, 2, 2, 2)
    assert code == 0 and out.shape == (3, 2, 2)
    assert np.array_equal(out[0], (img[..., 2].astype(np.float32) / np.float32(255.0)))   # R plane
    assert np.array_equal(out[1], (img[..., 1].astype(np.float32) / np.float32(255.0)))   # G
    assert np.array_equal(out[2], (img[..., 0].astype(np.float32) / np.float32(255.0)))   # B
    assert out[0, 1, 1] == np.float32(128) / np.float32(255)


def test_preprocess_is_stretch_not_letterbox():
    # 800x600 -> 416x416: both axes stretched independently, no padding
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (600, 800, 3), dtype=np.uint8)
    code, out = oracle_c.preprocess(img, 800, 600, 416, 416)
    assert code == 0
    iy, ix = _py_stretch(600, 416), _py_stretch(800, 416)
    ref = img[iy][:, ix][..., ::-1].transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    assert np.array_equal(out, ref)


def test_preprocess_rejects_wrong_length():
    # onnx_engine.cpp:659-665 -> ErrorCode::INVALID_INPUT (203)
    code, _ = oracle_c.preprocess(np.zeros(10, np.uint8), 4, 4, 8, 8)
    assert code == 203
    code, _ = oracle_c.preprocess(np.zeros(4 * 4 * 3 + 1, np.uint8), 4, 4, 8, 8)
    assert code == 203


def test_iou_hand_cases():
    # calculateIoU, onnx_engine.cpp:881-909 (centre format)
    assert oracle_c.iou((0.5, 0.5, 0.2, 0.2), (0.5, 0.5, 0.2, 0.2)) == pytest.approx(1.0, abs=1e-6)
    assert oracle_c.iou((0.2, 0.2, 0.1, 0.1), (0.8, 0.8, 0.1, 0.1)) == 0.0          # disjoint
    assert oracle_c.iou((0.5, 0.5, 0.0, 0.0), (0.5, 0.5, 0.0, 0.0)) == 0.0          # union == 0 guard (:904-908)
    # half-overlap along x: boxes 2x2 centred at (1,1) and (2,1): inter 2, union 6
    assert oracle_c.iou((1, 1, 2, 2), (2, 1, 2, 2)) == pytest.approx(1.0 / 3.0, rel=1e-6)
    # touching edges -> overlap 0
    assert oracle_c.iou((1, 1, 2, 2), (3, 1, 2, 2)) == 0.0


def _raw(nc, A):
    return np.zeros((4 + nc, A), np.float32)


def test_decode_threshold_and_argmax_rules():
    raw = _raw(3, 6)
    raw[0] = [10, 20, 30, 40, 50, 60]; raw[1] = 5; raw[2] = 8; raw[3] = 4
    raw[4:, 0] = [0.0, 0.0, 0.0]        # all scores <= 0 -> dropped: max_conf starts at 0.0f, strict '>' (:787-796)
    raw[4:, 1] = [0.5, 0.2, 0.1]        # == threshold -> kept ('>=' at :799)
    raw[4:, 2] = [0.7, 0.7, 0.3]        # tie -> lowest class index (strict '>' at :792)
    raw[4:, 3] = [0.49999, 0.1, 0.2]    # just below -> dropped
    raw[4:, 4] = [-1.0, -0.5, -0.1]     # negative -> dropped (id stays -1)
    raw[4:, 5] = [0.1, 0.2, 0.9]
    dets, anchors = oracle_c.decode_filter(raw, 100, 50, 0.5)
    assert list(anchors) == [1, 2, 5]                      # anchor order
    assert list(dets["class_id"]) == [0, 0, 2]
    assert list(dets["confidence"]) == [np.float32(0.5), np.float32(0.7), np.float32(0.9)]
    # boxes divided by REQUEST frame dims (:802-805), centre format kept
    assert dets["x"][0] == np.float32(20) / np.float32(100) and dets["y"][0] == np.float32(5) / np.float32(50)
    assert dets["w"][0] == np.float32(8) / np.float32(100) and dets["h"][0] == np.float32(4) / np.float32(50)


def _dets(rows):
    d = np.zeros(len(rows), DET_DTYPE)
    for i, r in enumerate(rows):
        d[i] = r
    return d


def test_nms_sort_order_and_cross_class():
    # sorted (class asc, conf desc) (:846-851); suppression only within a class (:866)
    d = _dets([(0.5, 0.5, 0.2, 0.2, 0.6, 1), (0.5, 0.5, 0.2, 0.2, 0.9, 0), (0.5, 0.5, 0.2, 0.2, 0.8, 1),
               (0.51, 0.5, 0.2, 0.2, 0.7, 0), (0.1, 0.1, 0.05, 0.05, 0.55, 0)])
    kept, anchors = oracle_c.nms(d, None, 0.45)
    assert list(kept["class_id"]) == [0, 0, 1]
    assert list(kept["confidence"]) == [np.float32(0.9), np.float32(0.55), np.float32(0.8)]
    assert list(anchors) == [1, 4, 2]


def test_nms_strict_threshold():
    # IoU exactly == threshold is KEPT ('>' at :871).  Boxes 2x2 at x=1 and x=2 -> IoU = 1/3 exactly in fp32? use computed value
    a, b = (1, 1, 2, 2), (2, 1, 2, 2)
    thr = oracle_c.iou(a, b)
    d = _dets([(*a, 0.9, 0), (*b, 0.8, 0)])
    kept, _ = oracle_c.nms(d, None, thr)
    assert len(kept) == 2
    kept, _ = oracle_c.nms(d, None, np.nextafter(np.float32(thr), np.float32(0)))
    assert len(kept) == 1


def test_nms_tie_break_is_anchor_order():
    # equal (class, conf): the reference's std::sort leaves the order unspecified; the oracle fixes anchor asc
    d = _dets([(0.5, 0.5, 0.2, 0.2, 0.7, 0), (0.52, 0.5, 0.2, 0.2, 0.7, 0), (0.9, 0.9, 0.1, 0.1, 0.7, 0)])
    kept, anchors = oracle_c.nms(d, np.array([7, 3, 5], np.int32), 0.45)
    assert list(anchors) == [3, 5]     # anchor 3 wins the tie, suppresses 7; 5 is disjoint


def test_nms_single_and_empty_passthrough():
    # <=1 candidate: returned as is (:841-843)
    d = _dets([(0.5, 0.5, 0.2, 0.2, 0.7, 3)])
    kept, _ = oracle_c.nms(d, None, 0.45)
    assert len(kept) == 1 and kept[0] == d[0]
    kept, _ = oracle_c.nms(_dets([]), None, 0.45)
    assert len(kept) == 0


def test_postprocess_no_cap_on_detections():
    # no max-detections cap, no top-k: 500 disjoint boxes all survive
    A = 500
    raw = _raw(2, A)
    raw[0] = np.arange(A) * 10 + 5; raw[1] = 5; raw[2] = 4; raw[3] = 4
    raw[4] = 0.9
    kept, _ = oracle_c.postprocess(raw, 5000, 10, 0.5, 0.45)
    assert len(kept) == A


def test_postprocess_matches_pure_python():
    rng = np.random.default_rng(7)
    nc, A = 5, 300
    raw = _raw(nc, A)
    raw[0:2] = rng.uniform(0, 100, (2, A)); raw[2:4] = rng.uniform(5, 40, (2, A))
    raw[4:] = rng.uniform(0, 1, (nc, A)) ** 3
    kept, anchors = oracle_c.postprocess(raw, 100, 100, 0.3, 0.45)
    # independent restatement in numpy/python
    cls = raw[4:].argmax(0); conf = raw[4:].max(0)
    idx = [i for i in range(A) if conf[i] >= np.float32(0.3) and conf[i] > 0]
    idx.sort(key=lambda i: (cls[i], -conf[i], i))
    box = (raw[:4] / np.float32(100)).T
    removed, out = set(), []
    for a, i in enumerate(idx):
        if i in removed:
            continue
        out.append(i)
        for j in idx[a + 1:]:
            if j in removed or cls[j] != cls[i]:
                continue
            if oracle_c.iou(box[i], box[j]) > np.float32(0.45):
                removed.add(j)
    assert list(anchors) == out
    assert np.array_equal(kept["class_id"], cls[out])


def test_reciprocal_normalisation_equals_division_after_16bit_rounding():
    """The tensor-core stem evaluates x/255 as x*(1/255).  In fp32 the two differ for 126 byte values, but after rounding
    to fp16 or bf16 (what the 16-bit modes store) they agree for all 256, so the stem feeds layer 0 exactly the
    reference's preprocessed values rounded once."""
    import torch
    u = np.arange(256, dtype=np.float32)
    a, b = u / np.float32(255.0), u * np.float32(1.0 / 255.0)
    assert not np.array_equal(a, b)
    assert np.array_equal(a.astype(np.float16), b.astype(np.float16))
    assert bool((torch.tensor(a).to(torch.bfloat16) == torch.tensor(b).to(torch.bfloat16)).all())
