"""Scratch perf probe (NOT the bench): builds a mid-size index with the oracle's threaded builder and times the search
kernel with CUDA events.  Usage: python tools/quick_perf.py [n] [dim] [nq] [ef]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import granne_b200  # noqa: E402
from helpers.data import clustered_vectors  # noqa: E402
from oracle import granne_oracle as go  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
ef = int(sys.argv[4]) if len(sys.argv) > 4 else 200
threads = os.cpu_count()
print("host cores", threads, flush=True)
t = time.time()
raw = clustered_vectors(n, dim, seed=1)
el = go.Elements.angular(raw)
g = go.GranneBuilder(el, num_neighbors=30, max_search=200).build(threads=threads)
print("oracle build %.1fs layers %s" % (time.time() - t, [g.layer_len(l) for l in range(g.num_layers())]), flush=True)
ib, eb = g.to_bytes(), el.to_bytes()
p = granne_b200.Granne.from_bytes(ib, "angular", eb)
q = clustered_vectors(max(nq, 16384), dim, seed=2)
# recall vs brute force (GPU matmul, off the hot path)
rows = torch.from_numpy(el.rows()).cuda()
for batch in sorted({nq, 16384}):
    qb = q[:batch]
    tq = torch.from_numpy(qb).cuda()
    stats = torch.zeros((batch, 4), dtype=torch.int64, device="cuda")
    out = p.search_batch_device(tq, ef, 10, stats=stats)
    torch.cuda.synchronize()
    p.stream_status()
    qn = torch.nn.functional.normalize(tq, dim=1)
    gt = torch.topk(qn @ rows.T, 10, dim=1).indices
    ids = out[0].long()
    rec = np.mean([len(set(gt[i].tolist()) & set(ids[i].tolist())) / 10 for i in range(min(batch, 512))])
    st = stats.cpu().numpy()
    nd, ne, nn = st[:, 0].mean(), st[:, 1].mean(), st[:, 2].mean()
    print("   slow-path queries %d  max n_expand %d" % (int((st[:, 3] & 1).sum()), int(st[:, 1].max())))
    bytes_q = nd * dim * 4 + nn * 4 + dim * 4
    times = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        p.search_batch_device(tq, ef, 10, out=out)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times[2:]))
    print("batch %d: %.3f ms  QPS %.0f  recall@10 %.3f  n_dist %.0f n_expand %.0f  bytes/q %.0f  -> %.1f GB/s"
          % (batch, ms, batch / ms * 1e3, rec, nd, ne, bytes_q, bytes_q * batch / ms / 1e6), flush=True)
    t = time.time()
    for _ in range(5):
        p.search_batch(qb, ef, 10)
    print("   e2e host API: %.3f ms/batch" % ((time.time() - t) / 5 * 1e3), flush=True)
# CPU oracle timing
gi = go.Granne.from_bytes(ib, el)
t = time.time()
gi.search_batch(q[:2048], ef, 10, threads=threads)
dt = time.time() - t
print("oracle CPU (%d threads, compressed adjacency): %.0f QPS" % (threads, 2048 / dt), flush=True)
