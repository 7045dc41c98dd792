"""Throughput of one large search launch vs the number of resident CTAs per SM (GRANNE_B200_MAX_CTAS_PER_SM is read
once per process, so each point runs in its own process).  Usage: python tools/occ_sweep.py <cap>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import granne_b200
from bench import clustered
n = 1_000_000
eb = granne_b200.elements_from_raw("angular", clustered(n, 128, 1234, 4096))
b = granne_b200.GranneBuilder("angular", eb, num_neighbors=30, max_search=200); b.build(); p = b.get_index(); b.close()
tq = torch.from_numpy(clustered(32768, 128, 4321, 4096)).cuda()
out = p.search_batch_device(tq, 200, 10)
torch.cuda.synchronize()
ts = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); p.search_batch_device(tq, 200, 10, out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print("cap %s: 32768 queries in %.2f ms -> %.2fM QPS" % (os.environ.get("GRANNE_B200_MAX_CTAS_PER_SM"), ms, 32768 / ms / 1e3))
