"""Profiling driver for ncu: builds a small index (oracle, threaded) and runs a few search batches.
Usage: python tools/prof_search.py [n] [nq] [iters]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import granne_b200  # noqa: E402
from helpers.data import clustered_vectors  # noqa: E402
from oracle import granne_oracle as go  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
el = go.Elements.angular(clustered_vectors(n, 128, seed=1))
g = go.GranneBuilder(el, num_neighbors=30, max_search=200).build(threads=os.cpu_count())
p = granne_b200.Granne.from_bytes(g.to_bytes(), "angular", el.to_bytes())
tq = torch.from_numpy(clustered_vectors(nq, 128, seed=2)).cuda()
out = None
for _ in range(iters):
    out = p.search_batch_device(tq, 200, 10, out=out)
torch.cuda.synchronize()
p.stream_status()
print("done")
