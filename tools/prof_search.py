"""Profiling driver for ncu: builds the bench index with the GPU builder and runs a few search batches.
Usage: python tools/prof_search.py [n] [nq] [iters] [dim]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import granne_b200  # noqa: E402
from bench import clustered  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dim = int(sys.argv[4]) if len(sys.argv) > 4 else 128
nc = max(8, int(4096 * (n / 1e6) ** 0.5))
eb = granne_b200.elements_from_raw("angular", clustered(n, dim, 1234, nc))
b = granne_b200.GranneBuilder("angular", eb, num_neighbors=30, max_search=200)
b.build()
p = b.get_index()
b.close()
tq = torch.from_numpy(clustered(nq, dim, 4321, nc)).cuda()
out = None
for it in range(iters):
    if it == iters - 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("prof")
    out = p.search_batch_device(tq, 200, 10, out=out)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
p.stream_status()
print("done")
