"""The C++ host adapter (B200InferenceEngine : IInferenceEngine) driven the way the reference server drives its engine."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "zero-latency-yolo_b200", "host")
EXE = os.path.join(ROOT, "zero-latency-yolo_b200", "lib", "host_test")


@pytest.fixture(scope="module")
def host_exe(built_lib):
    subprocess.check_call(["make", "-s", "-C", HOST])
    return EXE


def test_registry_errors_and_no_fallback(host_exe):
    out = subprocess.run([host_exe, "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_test cpu: ok" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("results", ["device", "host"])
def test_adapter_end_to_end_matches_c_abi(host_exe, model_n4, tmp_path, results):
    import zlb200
    from oracle import synth
    tensors, blob = model_n4
    wpath = tmp_path / "model.zlw"
    wpath.write_bytes(blob)
    frames = synth.frames_structured(6, 600, 800, seed=17)           # the reference client's default 800x600 frames
    fpath = tmp_path / "frames.bin"
    fpath.write_bytes(frames.tobytes())
    opath = tmp_path / "dets.bin"
    # results = "device": Detection records written by the GPU in the wire layout; "host": 24-byte records
    # widened by the adapter.  Both must give the same 40-byte Detections; the run also checks the four EventBus events.
    out = subprocess.run([host_exe, "gpu", str(wpath), "4", "fp16", str(opath), str(fpath), "6", "800", "600", results],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    e = zlb200.Engine(416, 416, 4, "n", precision=zlb200.FP16, max_batch=4, max_frame=(800, 600))
    e.load_weights_blob(blob)
    ref = e.infer(list(frames))
    e.close()
    raw = opath.read_bytes()
    off = 0
    for i in range(6):
        c = int(np.frombuffer(raw, "<u4", 1, off)[0]); off += 4
        recs = np.frombuffer(raw, np.uint8, c * 40, off).reshape(c, 40); off += c * 40
        assert c == len(ref[i])
        assert np.array_equal(recs[:, :24].reshape(-1), ref[i].view(np.uint8).reshape(-1))   # Detection's first 24 bytes == zl_det
        assert np.all(recs[:, 24:28] == 0)                                                   # track_id = 0 (onnx_engine.cpp:812)
        if c:
            ts = np.frombuffer(recs[:, 32:40].tobytes(), "<u8")
            assert ts.min() > 1_600_000_000_000 and len(np.unique(ts)) == 1                  # one wall-clock ms stamp per batch
    assert sum(len(r) for r in ref) > 5


@pytest.mark.gpu
def test_model_hot_reload(host_exe, tmp_path):
    """SURVEY.md §8f N4 / onnx_engine.cpp:473-515: the model file is replaced while frames flow; the monitor thread
    swaps the weights, detections change, a corrupt file leaves the running model in place."""
    from oracle import synth, yolov8_ref, zlw
    a = yolov8_ref.synthetic_model("n", 4, seed=0)
    b = yolov8_ref.synthetic_model("n", 4, seed=7)
    pa, pb = tmp_path / "a.zlw", tmp_path / "b.zlw"
    pa.write_bytes(zlw.dumps(a, "n", 4)); pb.write_bytes(zlw.dumps(b, "n", 4))
    frame = synth.frames_structured(1, 416, 416, seed=5678)[0]
    fpath = tmp_path / "frame.bin"
    fpath.write_bytes(frame.tobytes())
    out = subprocess.run([host_exe, "reload", str(pa), str(pb), "4", str(fpath), "416", "416"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_test reload: ok" in out.stdout
