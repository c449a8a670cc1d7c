"""The fused head kernel (head_fused.cu: model.22.cv2.l.2 + cv3.l.2 + Detect tail + decode/threshold in one tcgen05 kernel)
must give the SAME detections, byte for byte, as the unfused chain (conv kernels -> fp32 logits -> decode_filter_kernel),
which is the chain the raw-head parity tests pin to the oracle.  The switch (ZL_FUSE_HEAD) is read once per process, so each
arm runs in its own child process."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(tmp_path, fuse, branches=8):
    out = str(tmp_path / f"dets_fuse{fuse}_br{branches}.npz")
    env = dict(os.environ, ZL_FUSE_HEAD=str(fuse), ZL_HEAD_BRANCHES=str(branches))
    subprocess.check_call([sys.executable, os.path.join(HERE, "headfused_child.py"), out], env=env, timeout=900)
    return np.load(out)


def test_fused_head_gives_the_unfused_chain_s_detections_bit_for_bit(built_lib, tmp_path):
    sys.path.insert(0, HERE)
    import headfused_child
    fused, plain = _run(tmp_path, 1), _run(tmp_path, 0)
    total = 0
    for case in headfused_child.CASES:
        name, n = case[0], case[4]
        ops_f, ops_p = list(fused[f"{name}/ops"]), list(plain[f"{name}/ops"])
        assert "head.2+decode+filter" in ops_f and "decode+filter" not in ops_f, f"{name}: the fused kernel did not run ({ops_f[-4:]})"
        assert "decode+filter" in ops_p and "head.2+decode+filter" not in ops_p
        assert len(ops_f) == len(ops_p) - 6, "six 1x1 convs and the decode kernel are replaced by one launch"
        for i in range(n):
            a, b = fused[f"{name}/{i}"], plain[f"{name}/{i}"]
            assert a.shape == b.shape and np.array_equal(a, b), f"{name} frame {i}: fused and unfused detections differ"
            total += a.size // 40
    assert total > 50, f"the cases must produce detections to compare (got {total})"


def test_head_branches_on_side_streams_do_not_change_results(built_lib, tmp_path):
    """Batches of up to 8 frames run each level's Detect head on side streams next to the rest of the neck (a CUDA graph with
    forks when captured).  Same kernels, same inputs: the detections must be byte-identical to the single-stream order."""
    sys.path.insert(0, HERE)
    import headfused_child
    forked, serial = _run(tmp_path, 1, 8), _run(tmp_path, 1, 0)
    for case in headfused_child.CASES:
        name, n = case[0], case[4]
        for i in range(n):
            assert np.array_equal(forked[f"{name}/{i}"], serial[f"{name}/{i}"]), f"{name} frame {i}: side-stream head differs"
