"""Scalar model of the device kernel's merged candidate list (granne_b200/csrc/search_kernels.cuh, search_layer).

Host-logic test aid: proves on CPU that the single sorted list with 'expanded' flags, the ef-th-expanded threshold and
the strict-drop rule reproduce the reference's two-heap search_for_neighbors (src/index/mod.rs:999-1037) exactly,
and that plateau overflows are detected rather than silently mis-answered.
"""
import numpy as np


class Overflow(Exception):
    pass


def search_layer_model(get_neighbors, dist, entrypoint, ef, cap):
    """Returns (results [(id, d)], n_dist, n_expand).  Raises Overflow when an exact answer needs the slow path."""
    assert cap > ef
    L = []  # entries [dbits(int), id, expanded(bool)] sorted by (dbits, id)
    visited = {entrypoint}
    d0 = dist(entrypoint)
    n_dist, n_expand = 1, 0
    L.append([d0, entrypoint, False])
    while True:
        px = next((i for i, e in enumerate(L) if not e[2]), None)
        if px is None:
            break
        n_exp = sum(1 for e in L if e[2])
        thr = None
        if n_exp >= ef:
            thr = [e for e in L if e[2]][ef - 1][0]
        xd, xid, _ = L[px]
        if thr is not None and xd > thr:
            break
        L[px][2] = True
        n_exp += 1
        thr = [e for e in L if e[2]][ef - 1][0] if n_exp >= ef else None
        n_expand += 1
        for nb in get_neighbors(xid):
            if nb in visited:
                continue
            visited.add(nb)
            dn = dist(nb)
            n_dist += 1
            n_exp = sum(1 for e in L if e[2])
            if n_exp >= ef:
                thr = [e for e in L if e[2]][ef - 1][0]
                if not dn < thr:
                    continue
            key = (dn, nb)
            if len(L) == cap:
                last = L[-1]
                guard = L[ef - 1][0]
                if key > (last[0], last[1]):
                    if not guard < dn:
                        raise Overflow()
                    continue
                if not guard < last[0]:
                    raise Overflow()
                L.pop()
            pos = 0
            while pos < len(L) and (L[pos][0], L[pos][1]) < key:
                pos += 1
            L.insert(pos, [dn, nb, False])
    res = [(e[1], e[0]) for e in L if e[2]][:ef]
    return res, n_dist, n_expand


# ----------------------------------------------------------------------------------------------------------------------
# v2: mirrors the kernel's bookkeeping step by step (batched rank-based insertion, incremental threshold position,
# cursor), so that the logic can be validated on the CPU before it is transcribed to CUDA.
# ----------------------------------------------------------------------------------------------------------------------
FLAG = 1 << 63
MASK = FLAG - 1


def _key(dbits, idx):
    return (int(dbits) << 32) | int(idx)


def search_layer_model_v2(get_neighbors, dist_bits, entrypoint, ef, cap, batch=32, trace=None):
    """dist_bits(id) -> u32 bit pattern of the (non-negative) f32 distance.  Returns (res, n_dist, n_expand)."""
    assert cap > ef
    L = [0] * cap
    visited = {entrypoint}
    n_dist, n_expand = 1, 0
    L[0] = _key(dist_bits(entrypoint), entrypoint)
    n, n_exp, cursor = 1, 0, 0
    pos_thr, thr = -1, 0  # valid iff n_exp >= ef

    def ev(name):
        if trace is not None:
            trace[name] = trace.get(name, 0) + 1

    def first_unexpanded(start):
        for j in range(start, n):
            if not (L[j] & FLAG):
                return j
        return -1

    while True:
        px = first_unexpanded(cursor)
        if px < 0:
            break
        x = L[px]
        xd = (x >> 32) & 0x7FFFFFFF
        if n_exp >= ef and xd > thr:
            break
        L[px] = x | FLAG
        cursor = px + 1
        n_exp += 1
        if n_exp == ef:
            ev("became_full")
            # res just became full: its max is the last expanded entry of L
            pos_thr = max(j for j in range(n) if L[j] & FLAG)
            thr = (L[pos_thr] >> 32) & 0x7FFFFFFF
        elif n_exp > ef:
            ev("expand_before_thr" if px < pos_thr else "expand_tie_behind_thr")
            if px < pos_thr:
                # MaxSizeHeap::push replaced the max: new max = previous expanded entry before the old one
                j = pos_thr - 1
                while not (L[j] & FLAG):
                    j -= 1
                pos_thr = j
                thr = (L[pos_thr] >> 32) & 0x7FFFFFFF
                n_exp -= 0  # entries beyond pos_thr stay flagged but are dead; n_exp counts flagged entries <= pos_thr
                n_exp = ef
            else:
                # equal-distance tie beyond the max: expanded, not kept in res (dead entry)
                n_exp = ef
        n_expand += 1
        nbrs = [nb for nb in get_neighbors(key_id(x))]
        for b0 in range(0, len(nbrs), batch):
            new = []
            for nb in nbrs[b0:b0 + batch]:
                if nb not in visited:
                    visited.add(nb)
                    new.append(nb)
            if not new:
                continue
            keys = [_key(dist_bits(nb), nb) for nb in new]
            n_dist += len(new)
            passing = [k for k in keys if n_exp < ef or ((k >> 32) & 0x7FFFFFFF) < thr]
            if not passing:
                continue
            m = len(passing)
            # ranks
            rankL = [sum(1 for j in range(n) if (L[j] & MASK) < k) for k in passing]
            rankN = [sum(1 for k2 in passing if k2 < k) for k in passing]
            newpos = [a + b for a, b in zip(rankL, rankN)]
            shift = [sum(1 for r in rankL if r <= j) for j in range(n)]
            total = n + m
            nn = min(total, cap)
            merged = {}
            dropped = []
            for j in range(n - 1, -1, -1):
                np_ = j + shift[j]
                if np_ < cap:
                    merged[np_] = L[j]
                else:
                    dropped.append(L[j])
            for k, p in zip(passing, newpos):
                if p < cap:
                    merged[p] = k
                else:
                    dropped.append(k)
            assert sorted(merged.keys()) == list(range(nn))
            for p, v in merged.items():
                L[p] = v
            if dropped:
                ev("drop_batch")
                guard = (L[ef - 1] >> 32) & 0x7FFFFFFF
                dmin = min((v >> 32) & 0x7FFFFFFF for v in dropped)
                if not guard < dmin:
                    raise Overflow()
                dropped_flagged_live = 0
                if n_exp >= ef:
                    # every inserted key is < thr entry: it moves up by m
                    if pos_thr + m >= cap:
                        ev("thr_fell_off")
                        # the res max itself fell off: res now spans evicted entries -> "not full" regime
                        n_exp = sum(1 for j in range(nn) if L[j] & FLAG)
                        assert n_exp < ef
                    else:
                        pos_thr += m
                else:
                    n_exp -= sum(1 for v in dropped if v & FLAG)
            else:
                if n_exp >= ef:
                    pos_thr += m
            n = nn
            cursor = min(cursor, min(newpos))
            if n_exp >= ef:
                assert L[pos_thr] & FLAG and ((L[pos_thr] >> 32) & 0x7FFFFFFF) == thr
    out = [(key_id(v), (v >> 32) & 0x7FFFFFFF) for v in L[:n] if v & FLAG][:ef]
    return out, n_dist, n_expand


def key_id(v):
    return v & 0xFFFFFFFF
