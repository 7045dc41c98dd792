"""The batched list merge of the search kernel (search_layer_fast: rank among old entries + rank among new keys ->
final positions; per-row occupancy masks; rows rebuilt top-down as a gather), modelled lane by lane on the CPU
(tests/helpers/merge_model.py) and compared with a plain sort.  The list must stay sorted by (distance bits, id) — the
tuple order of (NotNan<f32>, usize), src/index/mod.rs:1001-1002 — keep exactly the `cap` smallest entries, keep the
'expanded' flags attached to their entries, and report how many flagged entries fell off the end."""
import numpy as np
import pytest

from helpers.merge_model import DMASK, FLAG, SENTINEL, merge


def pow2_above(cap):
    p = 32
    while p <= cap:
        p <<= 1
    return p


@pytest.mark.parametrize("R", [1, 3, 7, 15])
@pytest.mark.parametrize("seed", range(25))
def test_merge_equals_sorted_insert(R, seed):
    rng = np.random.default_rng(1000 * R + seed)
    cap, P = 32 * R, pow2_above(32 * R)
    n = int(rng.integers(1, cap + 1))
    few = bool(rng.integers(0, 2))                     # few distinct distances -> many exact ties
    dvals = rng.integers(1, 12 if few else 1 << 30, size=n + 32)
    ids = rng.permutation(100000)[:n + 32]
    old = sorted(zip(dvals[:n].tolist(), ids[:n].tolist()))
    flags = rng.integers(0, 2, size=n).astype(bool).tolist()
    Ld = [(d | (FLAG if f else 0)) for (d, _), f in zip(old, flags)] + [SENTINEL] * (P - n)
    Li = [i for _, i in old] + [0] * (P - n)
    m = int(rng.integers(1, 33))
    keys = list(zip(dvals[n:n + m].tolist(), ids[n:n + m].tolist()))
    if n == cap:  # the kernel's pre-filter: keys strictly farther than the tail of a full list never get here
        tail = old[-1][0]
        keys = [k for k in keys if not k[0] > tail] or [(tail, int(ids[-1]))]
        m = len(keys)
    ef = max(1, cap - 16)
    want = sorted([(d, i, f) for (d, i), f in zip(old, flags)] + [(d, i, False) for d, i in keys])
    dropped = want[cap:]
    want = want[:cap]
    n2, min_pos, drop_flagged, overflow = merge(Ld, Li, n, R, keys, ef)
    assert n2 == min(n + m, cap)
    got = [(Ld[j] & DMASK, Li[j], bool(Ld[j] >> 31)) for j in range(n2)]
    assert got == want
    assert all(Ld[j] == SENTINEL for j in range(n2, P))            # padding intact: the pop scan relies on it
    assert min_pos == min(want.index((d, i, False)) if (d, i, False) in want else cap for d, i in keys) or min_pos >= cap
    assert drop_flagged == sum(1 for d, i, f in dropped if f)      # only OLD entries carry flags
    if n + m > cap:
        assert overflow == (not (want[ef - 1][0] < want[cap - 1][0]))


def test_all_new_keys_beyond_a_full_list():
    """ties with the tail and larger ids: every new key lands at or beyond cap, nothing moves"""
    R, cap, P = 1, 32, 64
    Ld = [10] * 32 + [SENTINEL] * 32
    Li = list(range(32)) + [0] * 32
    keys = [(10, 100), (10, 50)]
    n2, min_pos, drop_flagged, overflow = merge(Ld, Li, 32, R, keys, 16)
    assert n2 == 32 and [x & DMASK for x in Ld[:32]] == [10] * 32 and Li[:32] == list(range(32))
    assert min_pos >= cap and drop_flagged == 0 and overflow  # a plateau: the exact answer needs the next pass
