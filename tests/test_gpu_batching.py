"""SURVEY 8f N1 (dynamic batching behind submitInference) and the class-weighted confidence filter (north_star (3))."""
import threading
import time

import numpy as np
import pytest

from oracle import oracle_c, synth

pytestmark = pytest.mark.gpu


def test_batch_window_coalesces_frames_and_keeps_order(built_lib, model_n4):
    """batch_window_us > 0: frames submitted inside the window run as ONE batched launch (batches < frames), every frame
    still gets exactly one callback, in submission order, with the detections it gets when run alone."""
    import zlb200
    tensors, blob = model_n4
    frames = list(synth.frames_structured(12, 416, 416, seed=300))
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=8, queue_depth=32, num_lanes=1, batch_window_us=200000)
    e.load_weights_blob(blob)
    e.warmup(1)
    alone = e.infer(frames)
    got, lock = [], threading.Lock()

    def cb(cid, fid, ts, status, dets):
        with lock:
            got.append((cid, fid, status, dets))

    e.set_callback(cb)
    b0, c0 = e.stats()["batches"], e.stats()["inference_count"]
    for i, f in enumerate(frames):
        assert e.submit(9, i, 1000 + i, f) == 0
    e.drain()
    st = e.stats()
    assert [g[1] for g in got] == list(range(12)) and all(g[2] == 0 for g in got)
    formed = st["batches"] - b0
    assert st["inference_count"] - c0 == 12
    assert formed <= 3, f"12 frames inside a 200 ms window with max_batch 8 must coalesce, got {formed} launches"   # 8 + 4 (the first may run alone)
    for i, g in enumerate(got):
        assert np.array_equal(g[3].view(np.uint8), alone[i].view(np.uint8))      # batching never changes a frame's result
    # window 0: take what is queued, never wait
    e0 = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=8, queue_depth=32, num_lanes=1, batch_window_us=0)
    e0.load_weights_blob(blob)
    e0.warmup(1)
    done = []
    e0.set_callback(lambda cid, fid, ts, status, dets: done.append(fid))
    t0 = time.perf_counter()
    assert e0.submit(1, 0, 0, frames[0]) == 0
    e0.drain()
    assert done == [0] and time.perf_counter() - t0 < 0.15
    e.close(); e0.close()


def test_full_queue_drops_with_inference_error(built_lib, model_n4):
    import zlb200
    tensors, blob = model_n4
    frames = list(synth.frames_structured(2, 416, 416, seed=301))
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=1, queue_depth=2, num_lanes=1, batch_window_us=0)
    e.load_weights_blob(blob)
    e.warmup(1)
    seen = []
    gate = threading.Event()
    e.set_callback(lambda cid, fid, ts, status, dets: (seen.append(fid), gate.wait(5)))   # the callback thread is held: the queue backs up
    codes = [e.submit(1, i, 0, frames[i % 2]) for i in range(8)]
    gate.set()
    e.drain()
    assert codes.count(zlb200.INFERENCE_ERROR) >= 1 and codes.count(0) >= 2          # network_server.cpp:213-215 treats 200 as "dropped"
    assert len(seen) == codes.count(0) and e.stats()["dropped_frames"] == codes.count(zlb200.INFERENCE_ERROR)
    e.close()


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_class_weights_scale_scores_before_the_argmax(built_lib, model_n80, precision):
    """class_weights (default 1.0 = the reference's behaviour): score_j *= w_j BEFORE the strict-> argmax and the >= threshold
    test.  Checked bit for bit against the oracle run on the engine's own raw head tensor with the multiply applied."""
    import zlb200
    tensors, blob = model_n80
    frames = list(synth.frames_structured(2, 640, 640, seed=55))
    rng = np.random.default_rng(3)
    w = rng.uniform(0.25, 1.5, 80).astype(np.float32)
    w[::7] = 0.0                                                   # a class switched off entirely
    prec = zlb200.FP32 if precision == "fp32" else zlb200.FP16
    plain = zlb200.Engine(640, 640, 80, "n", precision=prec, max_batch=2, conf=0.3)
    plain.load_weights_blob(blob)
    weighted = zlb200.Engine(640, 640, 80, "n", precision=prec, max_batch=2, conf=0.3, class_weights=w)
    weighted.load_weights_blob(blob)
    raw = plain.forward_raw(frames)                                # [n, 84, A]: the reference's output0
    got = weighted.infer(frames)
    base = plain.infer(frames)
    assert sum(len(d) for d in base) > 50
    changed = False
    for i in range(2):
        rw = raw[i].copy()
        rw[4:] = rw[4:] * w[:, None]                               # fp32 multiply, round to nearest: what the kernel does
        want, _ = oracle_c.postprocess(rw, 640, 640, 0.3, 0.45)
        assert np.array_equal(got[i].view(np.uint8), want.view(np.uint8))
        assert not np.any(np.isin(got[i]["class_id"], np.arange(0, 80, 7)))
        changed |= len(got[i]) != len(base[i]) or not np.array_equal(got[i].view(np.uint8), base[i].view(np.uint8))
    assert changed
    # through the two-step entry point as well
    plain.close(); weighted.close()
