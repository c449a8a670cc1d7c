"""the result wire layout written by the device.  The expected bytes are the reference's
DetectionResultPacket::serializeBody (src/common/protocol.h:541-567) restated here with struct.pack:
    uint32 frame_id | uint64 timestamp | uint16 count | count x Detection(40 B, src/common/types.h:20-26)."""
import struct
import threading

import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu


def serialize_body(frame_id, timestamp, dets, det_ts):
    """The reference's serialisation of one GameState, field by field (little endian, Detection padded to 40 bytes)."""
    out = struct.pack("<IQH", frame_id, timestamp, len(dets) & 0xFFFF)
    for d in dets:
        out += struct.pack("<fffffiI4xQ", d["x"], d["y"], d["w"], d["h"], d["confidence"], int(d["class_id"]), 0, det_ts)
    return out


def test_wire_blocks_equal_reference_serialisation(built_lib, model_n4):
    import zlb200
    tensors, blob = model_n4
    frames = list(synth.frames_structured(5, 416, 416, seed=5678)) + [synth.frames_const(1, 416, 416)[0]]
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=8, emit_wire=True)
    e.load_weights_blob(blob)
    e.warmup(1)
    ids = [7, 8, 0xFFFFFFFF, 10, 11, 12]
    tss = [1000 + i for i in range(5)] + [0xFFFFFFFFFFFFFFF0]
    det_ts = 1_729_250_000_123
    dets = e.infer(frames)
    wire = e.infer_wire(frames, ids, tss, det_ts)
    assert sum(len(d) for d in dets) > 20
    for i in range(len(frames)):
        assert len(wire[i]) == zlb200.WIRE_HEADER_BYTES + zlb200.WIRE_DET_BYTES * len(dets[i])
        assert wire[i] == serialize_body(ids[i], tss[i], dets[i], det_ts)
    # a batch that is not a power of two is padded internally: the padding frames' blocks never reach the caller
    w3 = e.infer_wire(frames[:3], ids[:3], tss[:3], det_ts)
    assert w3 == wire[:3]
    # more detections than the inline copy window (conf 0.01 keeps thousands)
    e2 = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=2, conf=0.01, emit_wire=True)
    e2.load_weights_blob(blob)
    d2 = e2.infer(frames[:2])
    w2 = e2.infer_wire(frames[:2], [1, 2], [3, 4], 5)
    assert sum(len(d) for d in d2) > 2 * 64
    for i in range(2):
        assert w2[i] == serialize_body([1, 2][i], [3, 4][i], d2[i], 5)
    e.close(); e2.close()


def test_wire_callback_async_path(built_lib, model_n4):
    import zlb200
    tensors, blob = model_n4
    frames = list(synth.frames_structured(6, 416, 416, seed=77))
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=4, queue_depth=16, emit_wire=True)
    e.load_weights_blob(blob)
    e.warmup(1)
    ref = e.infer(frames)
    got, lock = [], threading.Lock()

    def cb(cid, fid, ts, status, body):
        with lock:
            got.append((cid, fid, ts, status, body))

    e.set_wire_callback(cb)
    for i, f in enumerate(frames):
        assert e.submit(3, 100 + i, 5000 + i, f) == 0
    e.drain()
    assert [g[1] for g in got] == [100 + i for i in range(6)]
    for i, (cid, fid, ts, status, body) in enumerate(got):
        assert cid == 3 and status == 0
        fid_w, ts_w, cnt = struct.unpack_from("<IQH", body, 0)
        assert (fid_w, ts_w, cnt) == (100 + i, 5000 + i, len(ref[i]))
        assert len(body) == 14 + 40 * cnt
        rec = np.frombuffer(body[14:], np.dtype([("x", "<f4"), ("y", "<f4"), ("w", "<f4"), ("h", "<f4"), ("confidence", "<f4"), ("class_id", "<i4"),
                                                 ("track_id", "<u4"), ("pad", "<u4"), ("timestamp", "<u8")]))
        for k in ("x", "y", "w", "h", "confidence", "class_id"):
            assert np.array_equal(rec[k], ref[i][k])
        assert np.all(rec["track_id"] == 0) and np.all(rec["pad"] == 0) and (cnt == 0 or rec["timestamp"].min() > 1_600_000_000_000)
    e.close()


def test_wire_needs_emit_wire(built_lib, model_n4):
    import zlb200
    tensors, blob = model_n4
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=1)
    e.load_weights_blob(blob)
    with pytest.raises(zlb200.ZlError) as ei:
        e.infer_wire([synth.frames_const(1, 416, 416)[0]], [1], [2], 3)
    assert ei.value.code == zlb200.INVALID_ARGUMENT
    e.close()
