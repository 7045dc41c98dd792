"""Profiling driver for ncu: the bench workload (same generator, same per-box index cache as bench.py) and a few search
batches; the last one sits in the NVTX range "prof".
Usage: python tools/prof_search.py [config=c2] [nq=32768] [iters=3] [elements=0]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import granne_b200  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sys.argv = [sys.argv[0], "--config", cfg] + (["--elements", sys.argv[4]] if len(sys.argv) > 4 and int(sys.argv[4]) else [])
a = bench.parse_args()
a.nq = nq
granne_b200.load_library()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
tables = bench.device_tables(torch, dev, a, a.n)
cont = bench.make_container(torch, granne_b200, dev, a, a.n, bench.DATA_SEED, tables)
p, _, prov = bench.build_or_load_index(torch, granne_b200, a, dev, cont, bench.DATA_SEED)
if a.kind == "embeddings":
    tq = torch.from_numpy(cont.queries(nq, bench.QUERY_SEED)).to(dev)
else:
    tq = bench.make_queries_device(torch, dev, a, nq, bench.QUERY_SEED, tables)
cont.free()
stats = torch.zeros((nq, 4), dtype=torch.int64, device=dev)
out = None
for it in range(iters):
    if it == iters - 1:
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("prof")
    out = p.search_batch_device(tq, a.max_search, a.k, out=out, stats=stats if it == 0 else None)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
p.stream_status()
s = stats.cpu().numpy()
print("done: config %s n=%d nq=%d index %s; n_expand total %d, n_dist total %d" % (
    cfg, a.n, nq, prov.get("source"), int(s[:, 1].sum()), int(s[:, 0].sum())))
