"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py
Mode 1 (replicated, query-sharded) must be bit-identical to the oracle's search of the whole batch; mode 2
(range-partitioned) must equal the oracle's per-shard searches merged by (distance, global id)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import granne_b200  # noqa: E402
from granne_b200.distributed import FusedGather, PartitionedGranne, ReplicatedGranne, merge_topk_host, shard_bounds  # noqa: E402
from helpers.data import build_fixture, random_vectors  # noqa: E402
from oracle import granne_oracle as go  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    check(rank, world, local, dev)
    if rank == 0:
        print("multi-gpu check ok: world=%d replicated + partitioned parity" % world)
    dist.destroy_process_group()


def check(rank, world, local, dev):
    """The assertions proper, on an initialised NCCL process group (bench.py runs them before every N > 1 timing)."""
    granne_b200.load_library()
    q = random_vectors(1000, 32, seed=77)
    tq = torch.from_numpy(q).to(dev)

    # mode 1: replicated index (deterministic oracle build -> identical file image on every rank)
    el, g, ib, eb, _ = build_fixture(go, "angular", 10_000, 32, seed=1234, num_neighbors=10, max_search=50)
    idx = granne_b200.Granne.from_bytes(ib, "angular", eb, device=local)
    rep = ReplicatedGranne(idx)
    ids, d = rep.search_batch(tq, 50, 10)
    torch.cuda.synchronize()
    idx.stream_status()
    ref_ids, ref_d, _ = g.search_batch(q, 50, 10)
    assert np.array_equal(ids.cpu().numpy().view(np.uint32), ref_ids), "replicated ids"
    assert np.array_equal(d.cpu().numpy().view(np.uint32), ref_d.view(np.uint32)), "replicated dists"

    # mode 1 with the fused gather: kernels store their tiles straight into every peer's buffer (no collective)
    per = min(128, q.shape[0] // world)  # every rank needs a full slice of the 1000 test queries
    fg = FusedGather(per, 10, slots=2)
    for step in range(3):
        lq = tq[rank * per:(rank + 1) * per]
        idx.search_batch_device_gather(lq, fg.spec(step % 2, step + 1), 50, 10)
    torch.cuda.synchronize()
    idx.stream_status()
    dist.barrier()
    torch.cuda.synchronize()
    gi = fg.ids(0).cpu().numpy().view(np.uint32)   # slot 0 holds step 2 (same queries every step)
    gd = fg.dists(0).cpu().numpy()
    assert np.array_equal(gi, ref_ids[:world * per]), "fused gather ids"
    assert np.array_equal(gd.view(np.uint32), ref_d[:world * per].view(np.uint32)), "fused gather dists"
    assert fg.flags(0).cpu().tolist() == [3] * world and fg.flags(1).cpu().tolist() == [2] * world, "gather flags"

    # mode 2: one index per contiguous id range
    sizes = [3000 + 500 * r for r in range(world)]
    bases = [int(sum(sizes[:r])) for r in range(world)]
    el2, g2, ib2, eb2, _ = build_fixture(go, "angular", sizes[rank], 32, seed=100 + rank, num_neighbors=10,
                                         max_search=50)
    shard = granne_b200.Granne.from_bytes(ib2, "angular", eb2, device=local)
    part = PartitionedGranne(shard, shard_base=bases[rank])
    gi, gd = part.search_batch(tq, 50, 10)
    torch.cuda.synchronize()
    assert part.bases == bases
    parts = []
    for r in range(world):  # the oracle searches every shard on every rank (small)
        _, gr, _, _, _ = build_fixture(go, "angular", sizes[r], 32, seed=100 + r, num_neighbors=10, max_search=50)
        parts.append(gr.search_batch(q, 50, 10))
    ei, ed = merge_topk_host(np.stack([x[0] for x in parts]), np.stack([x[1] for x in parts]), bases, 10)
    assert np.array_equal(gi.cpu().numpy(), ei), "partitioned ids"
    assert np.array_equal(gd.cpu().numpy().view(np.uint32), ed.view(np.uint32)), "partitioned dists"
    dist.barrier()
    shard.close()
    idx.close()


if __name__ == "__main__":
    main()
