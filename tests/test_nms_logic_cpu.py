"""CPU restatements of the index arithmetic inside nms_kernel (csrc/postprocess.cu), checked exhaustively, and the
class-range split of the cluster launch checked against the oracle: the GPU tests prove the kernel, these pin the rules it
relies on so that a change to one of them fails here first (no GPU needed)."""
import numpy as np
import pytest

from oracle import oracle_c, synth


def _bitonic_network(P):
    """The kernel's schedule: k = 2..32 in registers; for k >= 64 two strides per barrier on quads while both are >= 32,
    one pair step at stride 32, then strides 16..1 in registers.  Returns the list of (i, j, ascending) comparators."""
    comps = []

    def reg_steps(k, jfirst):
        j = jfirst
        while j >= 1:
            for i in range(P):
                if i & j == 0:
                    comps.append((i, i | j, (i & k) == 0))
            j >>= 1

    k = 2
    while k <= min(32, P):
        reg_steps(k, k >> 1)
        k <<= 1
    k = 64
    while k <= P:
        j = k >> 1
        while j >= 64:
            h, lowm = j >> 1, (j >> 1) - 1
            seen = set()
            for q in range(P >> 2):
                i = (q & lowm) | ((q & ~lowm) << 2)
                quad = (i, i | h, i | j, i | j | h)
                assert not (seen & set(quad)), "quads overlap"
                seen |= set(quad)
                up = (i & k) == 0
                comps += [(quad[0], quad[2], up), (quad[1], quad[3], up), (quad[0], quad[1], up), (quad[2], quad[3], up)]
            assert len(seen) == P, "quads must cover every key exactly once"
            j >>= 2
        if j == 32:
            seen = set()
            for q in range(P >> 1):
                i = (q & 31) | ((q & ~31) << 1)
                assert i & 32 == 0
                seen |= {i, i | 32}
                comps.append((i, i | 32, (i & k) == 0))
            assert len(seen) == P
        reg_steps(k, 16)
        k <<= 1
    return comps


@pytest.mark.parametrize("P", [32, 64, 128, 512, 2048])
def test_sort_schedule_is_a_sorting_network(P):
    """Pair / quad index maps of the shared-memory steps cover every key once per step, and the whole schedule sorts
    (checked on random 64-bit keys with duplicates: a sorting network that sorts these sorts everything it will meet)."""
    comps = _bitonic_network(P)
    rng = np.random.default_rng(P)
    for trial in range(3):
        a = rng.integers(0, 1 << 62 if trial else 8, size=P, dtype=np.uint64)
        x = a.copy()
        for i, j, up in comps:
            if (x[i] > x[j]) == up:
                x[i], x[j] = x[j], x[i]
        assert np.array_equal(x, np.sort(a))


def _resolve_rows(und, rows):
    """resolve_rows of the kernel: settle, per round, every candidate whose earlier suppressors are all settled."""
    kept, rounds = 0, 0
    while und:
        kb = rb = 0
        for lane in range(32):
            if not (und >> lane) & 1:
                continue
            if rows[lane] & kept:
                rb |= 1 << lane
            elif not (rows[lane] & und):
                kb |= 1 << lane
        assert kb | rb, "the first undecided candidate always settles"
        kept |= kb
        und &= ~(kb | rb)
        rounds += 1
    return kept, rounds


def test_round_based_chain_equals_the_serial_greedy_chain():
    """applyNMS's chain over <= 32 candidates (keep the first alive, remove what it suppresses, repeat) and the round-based
    form used for the small class segments give the same kept set for any suppression matrix and any alive mask."""
    rng = np.random.default_rng(7)
    for trial in range(2000):
        density = rng.choice([0.02, 0.1, 0.3, 0.7])
        m = rng.random((32, 32)) < density
        m = np.triu(m | m.T, 1)                                  # m[t, i]: t < i and IoU(t, i) > thr (symmetric relation)
        rows = [int(sum(1 << t for t in range(i) if m[t, i])) for i in range(32)]
        alive = int(rng.integers(0, 1 << 32)) if trial % 3 else (1 << 32) - 1
        kept_serial, cur = 0, alive
        while cur:
            t = (cur & -cur).bit_length() - 1
            kept_serial |= 1 << t
            cur &= ~(1 << t)
            cur &= ~sum(1 << i for i in range(t + 1, 32) if m[t, i])
        kept_rounds, rounds = _resolve_rows(alive, rows)
        assert kept_rounds == kept_serial
        assert rounds <= max(1, bin(alive).count("1"))          # every round settles at least the first undecided candidate


def _class_owner(hist_excl, n_all, S):
    return np.minimum(S - 1, (hist_excl.astype(np.int64) * S) // n_all)


@pytest.mark.parametrize("S", [2, 4, 8])
def test_cluster_split_partitions_classes_and_preserves_the_result(S):
    """The split rule of the cluster launch: class c goes to rank min(S-1, excl_prefix(c) * S / n).  Ranks own contiguous,
    non-decreasing class ranges, and because classes never interact in applyNMS the concatenation of the ranks' results in
    rank order IS the frame's result — checked here with the oracle doing each rank's share."""
    raw = synth.stress_head(1, 12, 2100, seed=31 + S)[0]
    conf, iou = 0.15, 0.45
    whole, _ = oracle_c.postprocess(raw, 640, 640, conf, iou)
    cand, _ = oracle_c.decode_filter(raw, 640, 640, conf)
    n_all = len(cand)
    assert n_all > 256
    counts = np.bincount(cand["class_id"], minlength=12)
    excl = np.concatenate([[0], np.cumsum(counts)[:-1]])
    owner = _class_owner(excl, n_all, S)
    assert np.all(np.diff(owner) >= 0) and owner.min() >= 0 and owner.max() <= S - 1
    # every rank's share = whole classes; a class alone reproduces exactly its slice of the frame's result, so the ranks'
    # outputs concatenated in rank (= class) order are the frame's output
    for c in range(12):
        w = whole[whole["class_id"] == c]
        anchors_of_c = np.flatnonzero((np.argmax(raw[4:], axis=0) == c))
        sub = np.zeros_like(raw)
        sub[:4] = raw[:4]
        sub[4 + c, anchors_of_c] = raw[4 + c, anchors_of_c]
        alone, _ = oracle_c.postprocess(sub, 640, 640, conf, iou)
        assert np.array_equal(alone.view(np.uint8), w.view(np.uint8)), f"class {c} does not stand alone"
    kept_by_rank = [int(np.isin(whole["class_id"], np.flatnonzero(owner == r)).sum()) for r in range(S)]
    assert sum(kept_by_rank) == len(whole)
    order = np.concatenate([whole[np.isin(whole["class_id"], np.flatnonzero(owner == r))] for r in range(S)])
    assert np.array_equal(order.view(np.uint8), whole.view(np.uint8)), "rank order must be the sorted (class asc) order"
