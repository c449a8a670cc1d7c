"""GPU parity tests of the individual kernels, through the C-ABI, against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle_c, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(built_lib, model_n4):
    import zlb200
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP32, max_batch=2, max_frame=(1920, 1080))
    yield e
    e.close()


# ---------------------------------------------------------------- P1: bit-exact
@pytest.mark.parametrize("w,h", [(416, 416), (800, 600), (1920, 1080), (37, 53), (415, 417), (1, 1), (640, 360)])
def test_preprocess_bit_exact(eng, w, h):
    img = synth.frames_noise(1, h, w, seed=w * 7 + h)[0]
    code, ref = oracle_c.preprocess(img, w, h, 416, 416)
    assert code == 0
    out = eng.preprocess(img, w, h)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_preprocess_wrong_length_is_invalid_input(eng):
    import zlb200
    with pytest.raises(zlb200.ZlError) as ei:
        eng.preprocess(np.zeros(100, np.uint8), 10, 10)
    assert ei.value.code == zlb200.INVALID_INPUT
    assert "expected 300, got 100" in ei.value.message


def test_preprocess_other_model_size(built_lib):
    import zlb200
    e = zlb200.Engine(640, 384, 4, "n", precision=zlb200.BF16, max_batch=1, max_frame=(800, 600))
    img = synth.frames_structured(1, 600, 800)[0]
    _, ref = oracle_c.preprocess(img, 800, 600, 640, 384)
    assert np.array_equal(e.preprocess(img, 800, 600), ref)
    e.close()


# ---------------------------------------------------------------- convs
def _torch_conv(x, w, b, stride, act, res):
    xt = torch.tensor(x).permute(0, 3, 1, 2)
    wt = torch.tensor(w).permute(0, 3, 1, 2)
    y = torch.nn.functional.conv2d(xt, wt, torch.tensor(b), stride=stride, padding=w.shape[1] // 2)
    if act:
        y = torch.nn.functional.silu(y)
    y = y.permute(0, 2, 3, 1).numpy()
    if res is not None:
        y = y + res
    return y


def _bf(a):
    return torch.tensor(a).to(torch.bfloat16).to(torch.float32).numpy()


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, act, res
    (1, 13, 13, 64, 64, 3, 1, True, False),
    (2, 26, 26, 32, 32, 3, 1, True, True),
    (1, 20, 20, 16, 16, 3, 1, True, True),
    (1, 40, 40, 16, 32, 3, 2, True, False),
    (2, 27, 31, 64, 128, 3, 2, True, False),
    (1, 13, 13, 128, 256, 1, 1, True, False),
    (1, 52, 52, 48, 32, 1, 1, True, False),
    (2, 13, 13, 384, 256, 1, 1, True, False),
    (1, 26, 26, 64, 4, 1, 1, False, False),
    (1, 26, 26, 64, 80, 1, 1, False, False),
    (1, 10, 10, 256, 64, 3, 1, True, False),
    (1, 9, 9, 192, 576, 1, 1, True, False),
    (3, 16, 16, 96, 48, 3, 1, True, True),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fp32_simt(built_lib, case):
    import zlb200
    n, h, w, cin, cout, k, s, act, use_res = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = rng.normal(size=(n, h, w, cin)).astype(np.float32)
    wt = (rng.normal(size=(cout, k, k, cin)) / np.sqrt(cin * k * k)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    res = rng.normal(size=(n, ho, wo, cout)).astype(np.float32) if use_res else None
    y = zlb200.test_conv(x, wt, b, stride=s, act=act, res=res, impl=0)
    ref = _torch_conv(x, wt, b, s, act, res)
    assert np.abs(y - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tcgen05(built_lib, case, impl):
    import zlb200
    n, h, w, cin, cout, k, s, act, use_res = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = _bf(rng.normal(size=(n, h, w, cin)).astype(np.float32))
    wt = _bf((rng.normal(size=(cout, k, k, cin)) / np.sqrt(cin * k * k)).astype(np.float32))
    b = rng.normal(size=cout).astype(np.float32)
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    res = _bf(rng.normal(size=(n, ho, wo, cout)).astype(np.float32)) if use_res else None
    ref = _torch_conv(x, wt, b, s, act, res)
    # fp32 output: only accumulation order differs (fp32 accumulate in TMEM)
    y32 = zlb200.test_conv(x, wt, b, stride=s, act=act, res=res, impl=impl, out_f32=True)
    assert np.abs(y32 - ref).max() < 2e-3 * max(1.0, np.abs(ref).max()), "fp32-out mismatch"
    # bf16 output: one rounding on top
    y16 = zlb200.test_conv(x, wt, b, stride=s, act=act, res=res, impl=impl, out_f32=False)
    assert np.abs(y16 - ref).max() < 1.2e-2 * max(1.0, np.abs(ref).max()), "bf16-out mismatch"


HALO_CASES = [
    # n, h, w, cin, cout, res, fp16, out_f32, ctas (0 = one per SM)
    (1, 16, 8, 64, 64, False, False, True, 0),      # exactly one tile
    (2, 80, 80, 64, 64, True, False, False, 0),     # P3 head shape, SW128
    (1, 40, 40, 32, 32, True, False, True, 0),      # SW64
    (1, 24, 24, 16, 16, True, False, True, 0),      # SW32
    (2, 37, 29, 64, 80, False, False, True, 0),     # ragged tiles, Cout 80
    (1, 52, 52, 80, 80, False, False, True, 0),     # Cin 80 -> five 16-channel chunks
    (3, 33, 47, 64, 64, True, True, False, 3),      # fp16, 3 persistent CTAs -> many tiles per CTA, both TMEM buffers
    (2, 64, 64, 32, 48, False, True, True, 5),
    (1, 160, 40, 16, 16, True, True, False, 0),     # four stacked sub-tiles per tile
    (8, 20, 20, 128, 128, True, True, False, 0),    # weights do not fit: Cout split over CTAs (N-split)
    (4, 20, 20, 256, 64, False, False, True, 0),    # N-split, four 64-channel chunks
    (4, 40, 40, 128, 80, False, True, False, 0),    # uneven N-split (48 + 32)
    (2, 20, 20, 256, 80, False, True, True, 12),    # N-split with few CTAs
    (2, 70, 30, 32, 32, True, False, True, 2),      # two sub-tiles, ragged in both directions
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv_halo(built_lib, case):
    import zlb200
    n, h, w, cin, cout, use_res, fp16, out_f32, ctas = case
    cvt = _h if fp16 else _bf
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = cvt(rng.normal(size=(n, h, w, cin)).astype(np.float32))
    wt = cvt((rng.normal(size=(cout, 3, 3, cin)) / np.sqrt(cin * 9)).astype(np.float32))
    b = rng.normal(size=cout).astype(np.float32)
    res = cvt(rng.normal(size=(n, h, w, cout)).astype(np.float32)) if use_res else None
    ref = _torch_conv(x, wt, b, 1, True, res)
    y = zlb200.test_conv(x, wt, b, stride=1, act=True, res=res, impl=3, out_f32=out_f32, fp16=fp16, ntile_hint=ctas)
    tol = 2e-3 if (out_f32 or fp16) else 1.2e-2
    err = np.abs(y - ref)
    assert err.max() < tol * max(1.0, np.abs(ref).max()), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


PERSIST_S2_CASES = [
    # n, h, w, cin, cout, fp16, out_f32, ctas   (3x3 stride 2, pad 1)
    (1, 32, 16, 32, 64, False, True, 0),       # exactly one output tile
    (2, 160, 160, 32, 64, False, False, 0),    # layer 3 shape
    (1, 320, 320, 16, 32, True, False, 0),     # layer 1 shape: sub = 2, 32-B swizzle
    (2, 80, 80, 64, 128, True, False, 0),      # layer 5: two 32-channel chunks
    (4, 40, 40, 128, 256, False, False, 0),    # layer 7: N-split
    (3, 38, 54, 64, 64, True, True, 5),        # ragged output tiles (19 x 27), few CTAs
    (2, 80, 80, 64, 64, False, False, 0),      # layer 16
]


@pytest.mark.parametrize("case", PERSIST_S2_CASES)
def test_conv_persistent_stride2(built_lib, case):
    import zlb200
    n, h, w, cin, cout, fp16, out_f32, ctas = case
    cvt = _h if fp16 else _bf
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = cvt(rng.normal(size=(n, h, w, cin)).astype(np.float32))
    wt = cvt((rng.normal(size=(cout, 3, 3, cin)) / np.sqrt(cin * 9)).astype(np.float32))
    b = rng.normal(size=cout).astype(np.float32)
    ref = _torch_conv(x, wt, b, 2, True, None)
    y = zlb200.test_conv(x, wt, b, stride=2, act=True, impl=3, out_f32=out_f32, fp16=fp16, ntile_hint=ctas)
    tol = 2e-3 if (out_f32 or fp16) else 1.2e-2
    err = np.abs(y - ref)
    assert err.max() < tol * max(1.0, np.abs(ref).max()), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


PERSIST_1X1_CASES = [
    # n, h, w, cin, cout, fp16, out_f32, ctas
    (1, 16, 8, 64, 64, False, True, 0),       # one tile
    (2, 80, 80, 32, 32, False, False, 0),     # sub = 2
    (1, 160, 160, 48, 32, True, False, 0),    # Cin 48 -> three 16-channel chunks (layer 2 cv2)
    (2, 40, 40, 384, 128, True, False, 0),    # six 64-channel chunks (layer 12 cv1)
    (1, 20, 20, 256, 256, False, False, 4),   # N = 256: both accumulators fill TMEM
    (4, 20, 20, 128, 64, True, True, 3),      # fp32 output (direct store path)
    (1, 52, 52, 64, 80, False, True, 0),      # Cout 80
    (4, 26, 26, 192, 128, True, False, 7),
    (1, 24, 24, 16, 16, True, False, 0),      # sub = 4 (36 rows of 8)
    (8, 20, 20, 384, 256, True, False, 0),    # weights 196 KB: N-split
    (1, 13, 13, 256, 128, True, False, 0),    # 169 pixels (not a multiple of 8): native 13x13 view
    (3, 13, 13, 128, 64, False, True, 0),
    (2, 20, 20, 512, 256, False, False, 0),   # SPPF cv2 shape
]


@pytest.mark.parametrize("case", PERSIST_1X1_CASES)
def test_conv_persistent_1x1(built_lib, case):
    import zlb200
    n, h, w, cin, cout, fp16, out_f32, ctas = case
    cvt = _h if fp16 else _bf
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = cvt(rng.normal(size=(n, h, w, cin)).astype(np.float32))
    wt = cvt((rng.normal(size=(cout, 1, 1, cin)) / np.sqrt(cin)).astype(np.float32))
    b = rng.normal(size=cout).astype(np.float32)
    ref = _torch_conv(x, wt, b, 1, True, None)
    y = zlb200.test_conv(x, wt, b, stride=1, act=True, impl=3, out_f32=out_f32, fp16=fp16, ntile_hint=ctas)
    tol = 2e-3 if (out_f32 or fp16) else 1.2e-2
    err = np.abs(y - ref)
    assert err.max() < tol * max(1.0, np.abs(ref).max()), f"max err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"


def _h(a):
    return torch.tensor(a).to(torch.float16).to(torch.float32).numpy()


@pytest.mark.parametrize("case", CONV_CASES[:6])
def test_conv_tcgen05_fp16(built_lib, case):
    import zlb200
    n, h, w, cin, cout, k, s, act, use_res = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = _h(rng.normal(size=(n, h, w, cin)).astype(np.float32))
    wt = _h((rng.normal(size=(cout, k, k, cin)) / np.sqrt(cin * k * k)).astype(np.float32))
    b = rng.normal(size=cout).astype(np.float32)
    pad = k // 2
    ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    res = _h(rng.normal(size=(n, ho, wo, cout)).astype(np.float32)) if use_res else None
    ref = _torch_conv(x, wt, b, s, act, res)
    y32 = zlb200.test_conv(x, wt, b, stride=s, act=act, res=res, impl=2, out_f32=True, fp16=True)
    assert np.abs(y32 - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())
    y16 = zlb200.test_conv(x, wt, b, stride=s, act=act, res=res, impl=2, out_f32=False, fp16=True)
    assert np.abs(y16 - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())      # half has 3 more mantissa bits than bf16


@pytest.mark.parametrize("hint", [16, 32, 64])
def test_conv_tcgen05_n_split(built_lib, hint):
    import zlb200
    rng = np.random.default_rng(hint)
    x = _bf(rng.normal(size=(1, 13, 13, 128)).astype(np.float32))
    wt = _bf((rng.normal(size=(128, 3, 3, 128)) / 34).astype(np.float32))
    b = rng.normal(size=128).astype(np.float32)
    res = _bf(rng.normal(size=(1, 13, 13, 128)).astype(np.float32))
    ref = _torch_conv(x, wt, b, 1, True, res)
    y = zlb200.test_conv(x, wt, b, stride=1, act=True, res=res, impl=1, out_f32=True, ntile_hint=hint)
    assert np.abs(y - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())


# ---------------------------------------------------------------- F1 + N1: bit-exact
def _check_post(eng, raw, img_w, img_h, conf, iou):
    got = eng.decode_nms(raw, img_w, img_h, conf, iou)
    for f in range(raw.shape[0]):
        ref, _ = oracle_c.postprocess(raw[f], img_w, img_h, conf, iou)
        assert len(got[f]) == len(ref), f"frame {f}: kept {len(got[f])} vs oracle {len(ref)}"
        assert np.array_equal(got[f].view(np.uint8), ref.view(np.uint8)), f"frame {f}: detections differ bitwise"

@pytest.mark.parametrize("dtype", ["fp16", "bf16", "fp32"])
@pytest.mark.parametrize("shape", [(3, 20, 20, 32), (2, 13, 13, 16), (1, 7, 5, 48), (2, 40, 40, 128)])
def test_sppf_pools_are_exact(built_lib, dtype, shape):
    """SPPF = three chained 5x5/s1/p2 max-pools == windows 5 / 9 / 13 of the input with -inf padding.
    max() is exact in every format, so the kernel (packed 16-bit or fp32) must equal numpy on the rounded input bit for bit."""
    import torch
    import zlb200
    rng = np.random.default_rng(hash((dtype,) + shape) % (1 << 31))
    x = rng.standard_normal(shape).astype(np.float32) * 3
    x[rng.random(shape) < 0.05] = 0.0
    cat = zlb200.test_sppf_pool(x, dtype)
    c = shape[3]
    xq = cat[..., :c]                                             # the input as the kernel saw it (rounded to the format)
    t = torch.from_numpy(xq).permute(0, 3, 1, 2)
    if dtype != "fp32":
        assert np.array_equal(xq, (torch.from_numpy(x).to(torch.float16 if dtype == "fp16" else torch.bfloat16).float().numpy()))
    for i in range(3):
        t = torch.nn.functional.max_pool2d(t, 5, 1, 2)
        want = t.permute(0, 2, 3, 1).numpy()
        got = cat[..., (i + 1) * c:(i + 2) * c]
        assert np.array_equal(got, want), f"pool {i + 1} differs: {np.abs(got - want).max()}"



def test_post_small_random(eng):
    _check_post(eng, synth.stress_head(4, 4, 3549, seed=1, img=416), 416, 416, 0.25, 0.45)


def test_post_reference_defaults_nc80(eng):
    _check_post(eng, synth.stress_head(3, 80, 8400, seed=2), 640, 640, 0.5, 0.45)


def test_post_stress_low_threshold(eng):
    # cfg5 shape at a batch the oracle finishes quickly: nearly every anchor is a candidate
    _check_post(eng, synth.stress_head(6, 80, 8400, seed=42), 640, 640, 0.01, 0.45)


def test_post_adversarial_ties_and_clusters(eng):
    _check_post(eng, synth.stress_head_adversarial(3, 80, 8400, seed=43), 640, 640, 0.01, 0.45)
    _check_post(eng, synth.stress_head_adversarial(2, 4, 3549, seed=44, img=416), 800, 600, 0.3, 0.5)


def test_post_empty_single_and_odd_sizes(eng):
    raw = np.zeros((3, 6, 77), np.float32)                      # nothing passes
    raw[1, 0:4, 5] = [10, 10, 4, 4]; raw[1, 4 + 1, 5] = 0.9     # exactly one candidate in frame 1
    raw[2, 0:4, :] = np.array([[20.0], [20.0], [8.0], [8.0]]); raw[2, 4, :] = 0.8   # 77 identical boxes -> one survivor
    got = eng.decode_nms(raw, 100, 100, 0.5, 0.45)
    assert [len(g) for g in got] == [0, 1, 1]
    _check_post(eng, raw, 100, 100, 0.5, 0.45)


def test_post_threshold_edges(eng):
    raw = np.zeros((1, 7, 64), np.float32)
    raw[0, 0] = np.arange(64) * 3 + 10; raw[0, 1] = 50; raw[0, 2] = 12; raw[0, 3] = 12
    raw[0, 4] = 0.5                       # == threshold -> kept
    raw[0, 5, ::2] = 0.5                  # tie with class 0 -> class 0 wins
    raw[0, 6, 1::4] = np.nextafter(np.float32(0.5), np.float32(1))
    for iou in (0.1, 0.45, 0.6):
        _check_post(eng, raw, 640, 480, 0.5, iou)


@pytest.mark.parametrize("frames,nc,conf", [(20, 80, 0.05), (40, 80, 0.05), (24, 1, 0.3), (36, 3, 0.2), (9, 2, 0.01)])
def test_post_frames_split_over_cluster(eng, frames, nc, conf):
    """N1 with a frame's classes dealt to the 2 / 4 / 8 CTAs of a thread-block cluster (launch_nms picks the cluster size
    from the batch: 8 up to 16 frames, 4 up to 32, 2 up to 72).  Few classes leave ranks without any candidate; one class
    puts the whole frame on one rank.  Bit-exact against the oracle like every other N1 case."""
    raw = synth.stress_head(frames, nc, 8400 if nc > 3 else 2100, seed=900 + frames)
    got = eng.decode_nms(raw, 640, 640, conf, 0.45)
    ref, counts = oracle_c.postprocess_batch(raw, 640, 640, conf, 0.45)
    assert [len(g) for g in got] == list(counts)
    assert np.array_equal(np.concatenate(got).view(np.uint8), ref.view(np.uint8))


def test_post_split_with_dominant_class(eng):
    """One class holds most candidates (the shape of the bench frames: an 800-candidate class next to many small ones)."""
    raw = synth.stress_head(5, 80, 8400, seed=77)
    raw[:, 4 + 17, ::3] = np.maximum(raw[:, 4 + 17, ::3], 0.6)         # class 17 wins a third of the anchors
    _check_post(eng, raw, 640, 640, 0.3, 0.45)
    raw = synth.stress_head_adversarial(20, 80, 8400, seed=78)
    _check_post(eng, raw, 640, 640, 0.3, 0.5)


@pytest.mark.parametrize("conf", [0.3, 0.01])
def test_post_more_anchors_than_the_shared_memory_sort_holds(eng, conf):
    """A > 16384 anchors per frame (a 1280x1280 input has 33600): the keys are sorted in place in global memory, and at the
    low threshold (every anchor a candidate) the sorted boxes live in the global scratch as well — the paths no 640x640
    case reaches."""
    raw = synth.stress_head(2, 8, 20000, seed=505)
    _check_post(eng, raw, 640, 640, conf, 0.45)


def test_post_full_cfg5_batch128(eng):
    # BASELINE.json config 5 at full size: A=8400, nc=80, conf 0.01, batch 128
    raw = synth.stress_head(128, 80, 8400, seed=42)
    got = eng.decode_nms(raw, 640, 640, 0.01, 0.45)
    ref, counts = oracle_c.postprocess_batch(raw, 640, 640, 0.01, 0.45)
    assert [len(g) for g in got] == list(counts)
    assert np.array_equal(np.concatenate(got).view(np.uint8), ref.view(np.uint8))
    # size-independent property: NMS is idempotent — survivors fed back (as one-hot heads) all survive
    f = 0
    k = got[f]
    raw2 = np.zeros((1, 84, len(k)), np.float32)
    raw2[0, 0] = k["x"] * 640; raw2[0, 1] = k["y"] * 640; raw2[0, 2] = k["w"] * 640; raw2[0, 3] = k["h"] * 640
    raw2[0, 4 + k["class_id"], np.arange(len(k))] = k["confidence"]
    again = eng.decode_nms(raw2, 640, 640, 0.01, 0.45)[0]
    assert len(again) == len(k)
