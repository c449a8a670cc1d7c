"""N>1 host logic on CPU with gloo, world_size 2: the frame-sharding rule and the max-over-ranks timing
reduction used by bench.py.  The data path itself has no collective (frames never share state)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    from oracle import oracle_c, synth
    lo, hi = bench.shard_frames(n_total, rank, world)
    # each rank post-processes ITS frames only (oracle stands in for the engine on CPU) ...
    raw = synth.stress_head(n_total, 4, 300, seed=5)
    counts = [len(oracle_c.postprocess(raw[i], 416, 416, 0.3, 0.45)[0]) for i in range(lo, hi)]
    np.save(os.path.join(out_dir, f"counts_{rank}.npy"), np.array([lo, hi] + counts))
    # ... and the only cross-rank step is the max of the per-rank time
    fake_ms = 10.0 + 5.0 * rank
    mx = bench.reduce_max(fake_ms, dist)
    assert mx == 10.0 + 5.0 * (world - 1)
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_and_timing_reduction_world2(tmp_path):
    world, n_total = 2, 9
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle_c, synth
    raw = synth.stress_head(n_total, 4, 300, seed=5)
    want = [len(oracle_c.postprocess(raw[i], 416, 416, 0.3, 0.45)[0]) for i in range(n_total)]
    got, covered = [], []
    for r in range(world):
        a = np.load(tmp_path / f"counts_{r}.npy")
        covered.append((int(a[0]), int(a[1])))
        got += list(a[2:])
    assert covered == [(0, 5), (5, 9)]              # disjoint, contiguous, complete
    assert got == want                              # sharded result == unsharded result


def test_shard_rule_properties():
    import bench
    for n in (1, 7, 64, 256, 257):
        for w in (1, 2, 4, 8):
            spans = [bench.shard_frames(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
