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
