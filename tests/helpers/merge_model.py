"""Lane-level model of the batched list merge of search_layer_fast (granne_b200/csrc/search_kernels.cuh, "rebuild rows
top-down"): 32 lanes, a list of capacity 32*R stored row-major, all passing keys of one expansion inserted in one pass.

Every step is written the way the kernel does it — lower-bound rank on a sentinel-padded array, rank among the new
keys, per-row occupancy mask (REDUX.OR), popc-based gather index, rows processed top-down with a barrier between the
reads and the writes of a row — so that the CPU test (tests/test_merge_model_cpu.py) checks the ALGORITHM against a
plain sort, independently of the GPU parity tests."""

SENTINEL = 0xFFFFFFFF
DMASK = 0x7FFFFFFF
FLAG = 0x80000000


def lower_bound(Ld, P, d):
    """branch-free lower bound over the padded array (positions >= n hold sentinels whose masked value is maximal)"""
    lo, step = 0, P // 2
    while step >= 1:
        if (Ld[lo + step - 1] & DMASK) < d:
            lo += step
        step //= 2
    return lo


def merge(Ld, Li, n, R, keys, ef):
    """Ld/Li: lists of length P >= 32*R (padded with sentinels beyond n); keys: [(dbits, id)] of the passing lanes
    (<= 32, distinct ids).  Mutates Ld/Li exactly like the kernel; returns (new n, min_pos, drop_flagged, overflow)."""
    cap = 32 * R
    P = len(Ld)
    m = len(keys)
    assert 0 < m <= 32 and n <= cap
    # rank among the old entries (distance lower bound, refined by id on exact ties), then among the new keys
    new_pos = []
    for d, i in keys:
        lo = lower_bound(Ld, P, d)
        while lo < n and (Ld[lo] & DMASK) == d and Li[lo] < i:
            lo += 1
        rank_n = sum(1 for d2, i2 in keys if (d2, i2) < (d, i))
        new_pos.append(lo + rank_n)
    total = n + m
    min_pos = min(new_pos)
    drop_flagged = mdrop = 0
    if total > cap:
        mdrop = sum(1 for p in new_pos if p >= cap)
        odrop = (total - cap) - mdrop
        drop_flagged = sum(1 for j in range(n - odrop, n) if Ld[j] >> 31)
    kept = min(total, cap)
    kge = mdrop
    rt, rbm = (kept - 1) >> 5, min_pos >> 5
    for r in range(R - 1, -1, -1):
        if r > rt or r < rbm:
            continue
        occ = 0
        for p in new_pos:
            if (p >> 5) == r:
                occ |= 1 << (p & 31)
        kge += bin(occ).count("1")
        reads = {}
        for lane in range(32):  # all lanes read ...
            p = 32 * r + lane
            below = (m - kge) + bin(occ & ((1 << lane) - 1)).count("1")
            if not ((occ >> lane) & 1) and min_pos <= p < kept:
                reads[lane] = (Ld[p - below], Li[p - below])
        for lane, (vd, vi) in reads.items():  # ... barrier ... then write
            Ld[32 * r + lane], Li[32 * r + lane] = vd, vi
    for (d, i), p in zip(keys, new_pos):
        if p < cap:
            Ld[p], Li[p] = d, i
    overflow = False
    if total > cap:
        n = cap
        overflow = not ((Ld[ef - 1] & DMASK) < (Ld[cap - 1] & DMASK))
    else:
        n = total
    return n, min_pos, drop_flagged, overflow
