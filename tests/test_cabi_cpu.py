"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/zl_b200.h declares, and fails loudly (no CPU fallback) when no CUDA device exists."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zl_b200.h")).read()
    return sorted(set(re.findall(r"ZL_API\s+[\w\s\*]+?\b(zl_\w+)\s*\(", hdr)))


def test_header_declares_expected_surface():
    syms = _declared_symbols()
    for must in ("zl_engine_create", "zl_engine_submit", "zl_engine_set_callback", "zl_infer_batch",
                 "zl_preprocess", "zl_forward_raw", "zl_decode_nms", "zl_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    import zlb200
    syms = _declared_symbols()
    assert sorted(zlb200.EXPORTS) == syms, "python binding and header disagree"
    for s in syms:
        assert hasattr(built_lib, s), f"{s} missing from libzl_b200.so"


def test_test_hooks_are_not_in_the_product_library(built_lib):
    """zl_test_conv / zl_probe_* live in libzl_b200_test.so (include/zl_b200_test.h), never in the drop-in library."""
    import zlb200
    hdr = open(os.path.join(ROOT, "include", "zl_b200_test.h")).read()
    decl = sorted(set(re.findall(r"ZL_API\s+[\w\s\*]+?\b(zl_\w+)\s*\(", hdr)))
    assert decl == sorted(zlb200.TEST_EXPORTS)
    t = zlb200.testlib()
    for s in decl:
        assert hasattr(t, s), f"{s} missing from libzl_b200_test.so"
        assert not hasattr(built_lib, s), f"{s} must not be exported by libzl_b200.so"


def test_error_codes_match_reference_values():
    # src/common/result.h:14-48
    hdr = open(os.path.join(ROOT, "include", "zl_b200.h")).read()
    want = {"ZL_OK": 0, "ZL_INVALID_ARGUMENT": 2, "ZL_NOT_INITIALIZED": 3, "ZL_INFERENCE_ERROR": 200,
            "ZL_MODEL_NOT_FOUND": 201, "ZL_MODEL_LOAD_FAILED": 202, "ZL_INVALID_INPUT": 203,
            "ZL_SYSTEM_ERROR": 300, "ZL_INSUFFICIENT_RESOURCES": 303}
    for k, v in want.items():
        assert re.search(rf"\b{k}\s*=\s*{v}\b", hdr), k


def test_det_layout_is_prefix_of_reference_detection():
    import zlb200
    # Detection = {BoundingBox(4 x f32), confidence f32, class_id i32, track_id u32, pad, timestamp u64} = 40 B
    # (src/common/types.h:16-26); the device record is its first 24 bytes.
    assert zlb200.DET_DTYPE.itemsize == 24
    assert [zlb200.DET_DTYPE.fields[n][1] for n in ("x", "y", "w", "h", "confidence", "class_id")] == [0, 4, 8, 12, 16, 20]


def test_no_cpu_fallback_without_device(built_lib):
    import zlb200
    if built_lib.zl_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(zlb200.ZlError) as ei:
        zlb200.Engine(416, 416, 4, "n")
    assert ei.value.code == zlb200.INSUFFICIENT_RESOURCES
    assert "no CPU fallback" in ei.value.message
    with pytest.raises(zlb200.ZlError):
        zlb200.test_conv([[[[0.0] * 16]]], [[[[0.0] * 16]]] * 16, [0.0] * 16)


def test_null_handle_is_an_error_not_a_crash(built_lib):
    assert built_lib.zl_engine_warmup(None, 1) == 2
    assert built_lib.zl_engine_submit(None, 0, 0, 0, 1, 1, None, 3, 0) == 2
    assert built_lib.zl_engine_queue_size(None) == 0
    assert built_lib.zl_engine_destroy(None) == 0
    assert b"zl_b200" in built_lib.zl_version()


def test_no_load_is_hoisted_above_the_programmatic_launch_wait(built_lib):
    """Kernels launched as programmatic dependents (PDL) start while their predecessor still runs and must not read its
    output before `griddepcontrol.wait` (SASS: ACQBULK).  nvcc treats loads through `const __restrict__` pointers / `__ldg`
    as invariant (LDG.E.CONSTANT) and moves them across asm statements whatever their memory clobber: round 2 lost
    detections to exactly that in nms_kernel (the candidate count was loaded above the wait).  This test reads the SASS of
    the product library: in every kernel that waits, no invariant global load may precede the first ACQBULK."""
    import shutil
    import subprocess
    import zlb200
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", zlb200.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels, name, before_wait, waited, bad = 0, None, 0, False, {}
    for line in sass.splitlines():
        if "Function :" in line:
            name, before_wait, waited = line.split("Function :")[1].strip(), 0, False
            continue
        if name is None or waited:
            continue
        if "ACQBULK" in line:
            waited = True
            kernels += 1
            if before_wait:
                bad[name] = before_wait
        elif "LDG" in line and ".CONSTANT" in line:
            before_wait += 1
    assert kernels >= 5, f"expected the tcgen05 conv, head and NMS kernels to contain the wait, found {kernels}"
    assert not bad, f"invariant loads above griddepcontrol.wait: {bad}"
