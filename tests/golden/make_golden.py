"""Generates tests/golden/*.npz: small granne-format fixtures (index + elements file images, queries) with the search
results of the CPU oracle.  The reference itself cannot run in this image (Rust, no toolchain), so these vectors pin
the oracle against regressions and give the GPU tests a fixture that does not depend on rebuilding an index.
Run: python tests/golden/make_golden.py   (deterministic: single-threaded oracle build, fixed seeds)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers.data import build_fixture, random_vectors  # noqa: E402
from oracle import granne_oracle as go  # noqa: E402

CASES = [
    # name, kind, n, dim, M, ef_build, ef_search, k, extra
    ("angular_2000x32_m10", "angular", 2000, 32, 10, 50, 50, 10, {}),
    ("angular_int_800x100_m20", "angular_int", 800, 100, 20, 20, 40, 10, {}),
    ("embeddings_600x20_m16", "embeddings", 600, 20, 16, 30, 40, 10, {"num_embeddings": 100}),
]

for name, kind, n, dim, m, efb, efs, k, extra in CASES:
    el, g, ib, eb, mb = build_fixture(go, kind, n, dim, seed=n + dim, num_neighbors=m, max_search=efb, **extra)
    q = random_vectors(64, dim, seed=dim + 1)
    ids, dists, counts, stats = g.search_batch(q, efs, k, with_stats=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), kind=kind, index=np.frombuffer(ib, dtype=np.uint8),
                        elements=np.frombuffer(eb, dtype=np.uint8),
                        embeddings=np.frombuffer(mb, dtype=np.uint8) if mb is not None else np.zeros(0, np.uint8),
                        queries=q, max_search=efs, k=k, ids=ids, dists=dists, counts=counts, stats=stats)
    print(name, "index", len(ib), "elements", len(eb), "mean n_dist", stats[:, 0].mean())
