"""The kernel's visited set, modelled lane by lane on the CPU (tests/helpers/visited_model.py), behaves like the
reference's FxHashSet (src/index/mod.rs:1009-1026): across many warp-wide insert calls — with duplicate ids inside a
call, ids that collide in a bucket, full buckets that chain, and the speculative 'preloaded' snapshot used when no
insertion happened since it was taken — `insert` returns true exactly once per id and never loses an id."""
import numpy as np
import pytest

from helpers.visited_model import EMPTY, home_bucket, insert_warp


@pytest.mark.parametrize("nbuckets,universe", [(256, 400), (64, 200), (1500, 5000), (8, 24)])
@pytest.mark.parametrize("seed", range(6))
def test_bucketed_set_equals_a_python_set(nbuckets, universe, seed):
    rng = np.random.default_rng(100 * nbuckets + seed)
    table = [EMPTY] * (4 * nbuckets)
    seen = set()
    limit = (4 * nbuckets * 3) // 4          # the kernel flags the query before the load factor passes 3/4
    for call in range(400):
        k = int(rng.integers(1, 31))
        ids = [int(x) for x in rng.integers(0, universe, size=k)]
        if rng.random() < 0.3 and k > 2:      # duplicates inside one neighbour list (MultiSetVector allows them)
            ids[1] = ids[0]
            ids[-1] = ids[0]
        ids += [EMPTY] * (32 - len(ids))
        fresh = {i for i in ids if i != EMPTY} - seen
        if len(seen) + len(fresh) > limit:
            break
        snapshot = list(table) if rng.random() < 0.5 else None   # taken after the last insertion: still valid
        is_new, overflow = insert_warp(table, nbuckets, ids, snapshot)
        assert not overflow
        got = {i for i, f in zip(ids, is_new) if f}
        assert got == fresh                                       # exactly the ids that were not visited yet ...
        assert sum(is_new) == len(fresh)                          # ... each reported by exactly one lane
        assert not any(f for i, f in zip(ids, is_new) if i == EMPTY)
        seen |= fresh
    stored = [x for x in table if x != EMPTY]
    assert sorted(stored) == sorted(seen)                         # nothing lost, nothing stored twice
    for bkt in range(nbuckets):                                   # slots fill from slot 0 upwards
        row = table[4 * bkt:4 * bkt + 4]
        used = sum(1 for x in row if x != EMPTY)
        assert all(x != EMPTY for x in row[:used]) and all(x == EMPTY for x in row[used:])


def test_full_buckets_chain_to_the_next_one():
    nbuckets = 16
    table = [EMPTY] * (4 * nbuckets)
    same = [i for i in range(100000) if home_bucket(i, nbuckets) == 5][:11]   # 11 ids with the same home bucket
    is_new, overflow = insert_warp(table, nbuckets, same + [EMPTY] * 21)
    assert all(is_new[:11]) and not overflow
    assert sorted(table[20:32]) == sorted(same + [EMPTY])        # buckets 5, 6 full, bucket 7 holds three
    again, overflow = insert_warp(table, nbuckets, same[::-1] + [EMPTY] * 21)
    assert not any(again) and not overflow                       # found through the chain
