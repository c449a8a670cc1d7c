"""ONNX initialiser import.  The engine must read the Conv initialisers of an ultralytics YOLOv8 export
(the file the reference loads: start.sh:122-125, onnx_engine.cpp:977-1040).  onnx / ultralytics are not available
offline, so the test writes the ModelProto itself (tests/onnx_writer.py) from the same tensors as the ZLW1 container."""
import numpy as np
import pytest

import onnx_writer


@pytest.mark.parametrize("mode", ["raw", "float_data", "mixed"])
def test_probe_checksum_equals_zlw(built_lib, model_n4, mode):
    import zlb200
    tensors, blob = model_n4
    want = zlb200.model_probe(blob)
    got = zlb200.model_probe(onnx_writer.model(tensors, mode))
    assert got == want                      # (scale, nc, n_tensors, checksum over names, shapes and values)
    assert want[0] == 0 and want[1] == 4 and want[2] == 2 * 63


def test_probe_fp16_export_and_scales(built_lib):
    import zlb200
    from conftest import synthetic_model
    for scale, sid in (("s", 1), ("m", 2)):
        tensors, blob = synthetic_model(scale, 80)
        sc, nc, nt, _ = zlb200.model_probe(onnx_writer.model(tensors, "raw"))
        assert (sc, nc) == (sid, 80) and nt == 2 * (63 if scale == "s" else 83)
    tensors, blob = synthetic_model("n", 4)
    half = {k: np.asarray(v).astype(np.float16).astype(np.float32) for k, v in tensors.items()}
    from oracle import zlw
    assert zlb200.model_probe(onnx_writer.model(tensors, "fp16")) == zlb200.model_probe(zlw.dumps(half, "n", 4))


@pytest.mark.parametrize("damage", ["truncated", "no_graph", "huge_dims", "zero_dim", "size_mismatch"])
def test_corrupt_onnx_is_rejected_cleanly(built_lib, model_n4, damage):
    import zlb200
    tensors, _ = model_n4
    good = onnx_writer.model(tensors)
    if damage == "truncated":
        bad = good[: len(good) // 2]
    elif damage == "no_graph":
        bad = onnx_writer._key(1, 0) + onnx_writer._varint(8) + onnx_writer._ld(2, b"x" * 64)
    else:
        t = dict(tensors)
        name = "model.0.conv.weight"
        body = onnx_writer.tensor(name, np.asarray(t.pop(name)))
        if damage == "huge_dims":      # dims whose product wraps 64 bits must not pass the size checks
            body = b"".join(onnx_writer._key(1, 0) + onnx_writer._varint(d) for d in (1 << 31, 1 << 31, 4, 1)) + body[body.index(onnx_writer._key(2, 0)):]
        elif damage == "zero_dim":
            body = b"".join(onnx_writer._key(1, 0) + onnx_writer._varint(d) for d in (16, 0, 3, 3)) + body[body.index(onnx_writer._key(2, 0)):]
        else:                          # raw_data shorter than the dims say
            body = body[:-8]
            body = body[: body.rindex(onnx_writer._key(9, 2))] + onnx_writer._ld(9, b"\0" * 100)
        g = onnx_writer._ld(5, body)
        for k, v in sorted(t.items()):
            g += onnx_writer._ld(5, onnx_writer.tensor(k, v))
        bad = onnx_writer._ld(7, g)
    with pytest.raises(zlb200.ZlError) as ei:
        zlb200.model_probe(bad)
    assert ei.value.code == zlb200.MODEL_LOAD_FAILED


@pytest.mark.gpu
def test_onnx_weights_give_bit_identical_detections(built_lib, model_n4, tmp_path):
    import zlb200
    from oracle import synth
    tensors, blob = model_n4
    frames = list(synth.frames_structured(3, 416, 416, seed=5678))
    a = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=4)
    a.load_weights_blob(blob)
    b = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=4)
    path = tmp_path / "yolov8n.onnx"
    path.write_bytes(onnx_writer.model(tensors, "mixed"))
    b.load_weights(str(path))                                        # the model_path route the adapter uses
    da, db = a.infer(frames), b.infer(frames)
    assert sum(len(d) for d in da) > 10
    for x, y in zip(da, db):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    # wrong class count / scale for this engine: refused, the running model stays
    from conftest import synthetic_model
    t80, _ = synthetic_model("n", 80)
    with pytest.raises(zlb200.ZlError) as ei:
        b.load_weights_blob(onnx_writer.model(t80))
    assert ei.value.code == zlb200.MODEL_LOAD_FAILED
    assert np.array_equal(b.infer(frames)[0].view(np.uint8), da[0].view(np.uint8))
    a.close(); b.close()
