"""Prints traversal statistics of the bench workload (uses the GPU builder). Usage: python tools/bench_stats.py [n]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import granne_b200
from bench import clustered
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nc = max(8, int(4096 * (n / 1e6) ** 0.5))
eb = granne_b200.elements_from_raw("angular", clustered(n, 128, 1234, nc))
b = granne_b200.GranneBuilder("angular", eb, num_neighbors=30, max_search=200); b.build(); p = b.get_index(); b.close()
tq = torch.from_numpy(clustered(4096, 128, 4321, nc)).cuda()
stats = torch.zeros((4096, 4), dtype=torch.int64, device="cuda")
p.search_batch_device(tq, 200, 10, stats=stats); torch.cuda.synchronize(); p.stream_status()
st = stats.cpu().numpy()
print("n_dist %.0f n_expand %.0f (max %d) n_nbr %.0f slow %d" % (st[:,0].mean(), st[:,1].mean(), st[:,1].max(), st[:,2].mean(), (st[:,3]&1).sum()))
degs = [len(p.get_neighbors(i)) for i in range(0, n, n // 2000)]
print("bottom degree mean %.1f min %d max %d" % (np.mean(degs), min(degs), max(degs)))
