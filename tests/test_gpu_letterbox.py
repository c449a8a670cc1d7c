"""ZL_PRE_LETTERBOX (north_star (1)): aspect-preserving resize, centred, 114-grey border, boxes mapped back to the request
frame.  NOT a parity mode — the reference stretches (onnx_engine.cpp:673-693) — so it is checked against a numpy
restatement of the same mapping and against the parity path run on a pre-letterboxed image."""
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


def letterbox_numpy(img, mw, mh):
    h, w = img.shape[:2]
    gain = min(np.float32(mw) / np.float32(w), np.float32(mh) / np.float32(h))
    nw = min(int(np.float32(w) * gain + np.float32(0.5)), mw)
    nh = min(int(np.float32(h) * gain + np.float32(0.5)), mh)
    px, py = (mw - nw) // 2, (mh - nh) // 2
    sw, sh = np.float32(w) / np.float32(nw), np.float32(h) / np.float32(nh)
    ix = np.minimum((np.arange(nw, dtype=np.float32) * sw).astype(np.int32), w - 1)
    iy = np.minimum((np.arange(nh, dtype=np.float32) * sh).astype(np.int32), h - 1)
    out = np.full((mh, mw, 3), 114, np.uint8)
    out[py:py + nh, px:px + nw] = img[iy][:, ix]
    return out, float(gain), px, py


@pytest.mark.parametrize("w,h", [(800, 600), (600, 800), (416, 416), (1920, 1080), (37, 53)])
def test_letterbox_preprocess_matches_numpy(built_lib, w, h):
    import zlb200
    img = synth.golden_bytes((h, w, 3), w + h)
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP32, max_batch=1, max_frame=(max(w, 416), max(h, 416)), letterbox=True)
    got = e.preprocess(img, w, h)
    lb, _, _, _ = letterbox_numpy(img, 416, 416)
    want = lb[..., ::-1].transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    e.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_letterbox_detections_equal_parity_path_on_prepadded_image(built_lib, model_n4, precision):
    import zlb200
    tensors, blob = model_n4
    prec = zlb200.FP32 if precision == "fp32" else zlb200.FP16
    frames = [synth.frames_structured(1, 600, 800, seed=5)[0], synth.frames_structured(1, 800, 600, seed=6)[0], synth.frames_structured(1, 416, 416, seed=7)[0]]
    lbx = zlb200.Engine(416, 416, 4, "n", precision=prec, max_batch=4, max_frame=(800, 800), letterbox=True)
    lbx.load_weights_blob(blob)
    par = zlb200.Engine(416, 416, 4, "n", precision=prec, max_batch=4)
    par.load_weights_blob(blob)
    got = lbx.infer(frames)
    n = 0
    for f, g in zip(frames, got):
        img, gain, px, py = letterbox_numpy(f, 416, 416)
        ref = par.infer([img])[0]                          # parity path, identity-size sampling: boxes normalised by 416
        assert len(g) == len(ref) and np.array_equal(g["class_id"], ref["class_id"]) and np.array_equal(g["confidence"], ref["confidence"])
        h, w = f.shape[:2]
        assert np.allclose(g["x"], (ref["x"] * 416 - px) / gain / w, atol=2e-6)
        assert np.allclose(g["y"], (ref["y"] * 416 - py) / gain / h, atol=2e-6)
        assert np.allclose(g["w"], ref["w"] * 416 / gain / w, atol=2e-6)
        assert np.allclose(g["h"], ref["h"] * 416 / gain / h, atol=2e-6)
        n += len(g)
    assert n > 10
    lbx.close(); par.close()
