"""Lane-level model of the fast pass's visited set (granne_b200/csrc/search_kernels.cuh, vis_bucket_insert): a table of
16-byte buckets (4 u32 slots, filled from slot 0 upwards, 0xFFFFFFFF = empty) private to one warp; 32 lanes probe
and insert concurrently without atomics — duplicates inside one neighbour list are resolved with match.any on the id,
lanes that share a bucket take consecutive free slots (match.any on the bucket), a full bucket chains to the next.
Semantics to reproduce: FxHashSet::insert (src/index/mod.rs:1009-1010,1016,1026) — true exactly once per id."""

EMPTY = 0xFFFFFFFF
HASH_MUL = 0x9E3779B1


def home_bucket(node_id, nbuckets):
    return (((node_id * HASH_MUL) & 0xFFFFFFFF) * nbuckets) >> 32


def insert_warp(table, nbuckets, ids, preloaded=None):
    """table: list of nbuckets*4 u32.  ids: up to 32 ids (EMPTY = lane without a neighbour).  `preloaded`: optional
    snapshot of the table taken earlier (the speculative bucket copy) that replaces the FIRST round's loads.
    Returns (is_new per lane, overflow)."""
    lanes = len(ids)
    b = [home_bucket(i, nbuckets) for i in ids]
    pending = [i != EMPTY for i in ids]
    is_new = [False] * lanes
    for probe in range(66):
        src = preloaded if (probe == 0 and preloaded is not None) else table
        v = [src[4 * b[l]:4 * b[l] + 4] for l in range(lanes)]          # every lane loads its bucket
        for l in range(lanes):
            if ids[l] in v[l]:
                pending[l] = False
        ins = [pending[l] and v[l][3] == EMPTY for l in range(lanes)]
        if any(ins):
            for l in range(lanes):                                       # same id twice: only the first lane inserts
                if ins[l] and any(ins[m] and ids[m] == ids[l] for m in range(l)):
                    ins[l] = False
                    pending[l] = False
            for l in range(lanes):                                       # same bucket: consecutive free slots
                if ins[l]:
                    used = sum(1 for x in v[l][:3] if x != EMPTY)
                    slot = used + sum(1 for m in range(l) if ins[m] and b[m] == b[l])
                    if slot < 4:
                        table[4 * b[l] + slot] = ids[l]
                        is_new[l] = True
                        pending[l] = False
        for l in range(lanes):
            if pending[l]:
                b[l] = 0 if b[l] + 1 == nbuckets else b[l] + 1           # bucket full: continue in the next one
        if not any(pending):
            return is_new, False
        if probe >= 64:
            return is_new, True
    return is_new, True
