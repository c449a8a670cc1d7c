import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zero-latency-yolo_b200", "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree shared library; built on demand (nvcc cross-compiles without a GPU)."""
    import zlb200
    if not os.path.exists(zlb200.LIB_PATH) or not os.path.exists(zlb200.TEST_LIB_PATH):
        import subprocess
        subprocess.check_call(["bash", os.path.join(ROOT, "zero-latency-yolo_b200", "csrc", "build.sh")])
    return zlb200.lib()


_MODEL_CACHE = {}


def synthetic_model(scale, nc, seed=0):
    from oracle import yolov8_ref, zlw
    key = (scale, nc, seed)
    if key not in _MODEL_CACHE:
        t = yolov8_ref.synthetic_model(scale, nc, seed)
        _MODEL_CACHE[key] = (t, zlw.dumps(t, scale, nc))
    return _MODEL_CACHE[key]


@pytest.fixture(scope="session")
def model_n4():
    return synthetic_model("n", 4)


@pytest.fixture(scope="session")
def model_n80():
    return synthetic_model("n", 80)
