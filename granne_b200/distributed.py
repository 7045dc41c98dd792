"""Multi-GPU search: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the plumbing.

Queries are independent (Granne::search takes &self, src/index/mod.rs:140-150), so the path shards without any
data-path collective; the only exchange is the collection of the small result tiles.

Mode 1 — ReplicatedGranne: the staged index is replicated on every GPU, each rank searches its own slice of the query
batch, the [nq_local, k] result tiles are all-gathered.  Results are bit-identical to a single-GPU search of the same
queries.

Mode 2 — PartitionedGranne: the element set is split into contiguous id ranges with one independent granne index per
range (the reference's own notion of sharding, src/elements/embeddings/parsing.rs:63-100); every rank searches ALL
queries on its shard, the per-shard tiles are all-gathered and merged per query by (distance, global id) — the tuple
order of into_sorted_vec (src/index/mod.rs:1036).  The merge runs on the GPU (granne_b200_merge_topk_device).

The collective calls work on whatever backend the process group uses.  The product path merges on the GPU; only the
host-logic tests (gloo, CPU tensors, `host_merge=True`) use merge_topk_host below, the numpy statement of the same
k-way merge.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous, balanced [begin, end) of `n` items for `rank` (the first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def merge_topk_host(part_ids, part_dists, part_base, k):
    """k-way merge of per-shard tiles on the host.  part_ids: [P, nq, k] uint32 (0xFFFFFFFF padded), part_dists:
    [P, nq, k] float32, part_base: P global id offsets.  Returns (int64 [nq, k] padded with -1, float32 padded +inf)."""
    part_ids = np.asarray(part_ids)
    part_dists = np.asarray(part_dists)
    P, nq, kk = part_ids.shape
    out_ids = np.full((nq, k), -1, dtype=np.int64)
    out_d = np.full((nq, k), np.inf, dtype=np.float32)
    for q in range(nq):
        cand = []
        for p in range(P):
            for j in range(kk):
                lid = int(part_ids[p, q, j])
                if lid == 0xFFFFFFFF:
                    break
                cand.append((part_dists[p, q, j], int(part_base[p]) + lid))
        cand.sort(key=lambda t: (t[0], t[1]))
        for r, (d, gid) in enumerate(cand[:k]):
            out_ids[q, r] = gid
            out_d[q, r] = d
    return out_ids, out_d


class ReplicatedGranne:
    """Index replicated per rank, query batch sharded, results all-gathered.

    local_search(queries, max_search, k) -> (ids int32/uint32 [nq_local, k], dists float32 [nq_local, k]) as torch
    tensors on the collective's device; by default the rank's granne_b200.Granne.search_batch_device."""

    def __init__(self, index=None, group=None, local_search=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.index = index
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local = local_search or self._search_device

    def _search_device(self, q, max_search, k):
        ids, dists, _ = self.index.search_batch_device(q, max_search, k)
        return ids, dists

    def my_slice(self, nq_global):
        return shard_bounds(nq_global, self.world, self.rank)

    def search_local_shard(self, local_queries, max_search=200, k=10, out=None):
        """Every rank passes ITS slice (equal sizes); returns the gathered [world * nq_local, k] tiles in rank order."""
        import torch

        ids, dists = self._local(local_queries, max_search, k)
        if self.world == 1:
            return ids, dists
        nq = ids.shape[0]
        if out is None:
            out = (torch.empty((self.world * nq, k), dtype=ids.dtype, device=ids.device),
                   torch.empty((self.world * nq, k), dtype=dists.dtype, device=dists.device))
        self.dist.all_gather_into_tensor(out[0], ids.contiguous(), group=self.group)
        self.dist.all_gather_into_tensor(out[1], dists.contiguous(), group=self.group)
        return out

    def search_batch(self, queries, max_search=200, k=10):
        """`queries` is the same global batch on every rank; each rank searches its slice and all ranks receive the
        full result in query order (uneven slices are padded for the collective and trimmed afterwards)."""
        import torch

        nq = queries.shape[0]
        b, e = self.my_slice(nq)
        per = -(-nq // self.world)
        local = queries[b:e]
        if e - b < per:  # pad with a repeat of the last query so every rank contributes `per` rows
            pad = queries[e - 1:e] if e > b else queries[:1]
            local = torch.cat([local] + [pad] * (per - (e - b)), dim=0)
        ids, dists = self.search_local_shard(local, max_search, k)
        if self.world == 1:
            return ids[:nq], dists[:nq]
        keep = []
        for r in range(self.world):
            rb, re = shard_bounds(nq, self.world, r)
            keep.append(torch.arange(r * per, r * per + (re - rb), device=ids.device))
        keep = torch.cat(keep)
        return ids[keep], dists[keep]


class FusedGather:
    """Peer-mapped gathered result buffers for mode 1 (replicated index): the search kernels store each rank's tile
    straight into every rank's buffer over NVLink (granne_b200_search_batch_device_gather) — no collective per step.

    `slots` independent buffers ([world * nq, k] ids + dists each) allow several steps in flight.  After a step,
    flags(slot)[r] == seq tells that rank r's tile of that step has landed in THIS rank's buffer."""

    def __init__(self, nq_local, k, slots=4, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from .api import PeerGather

        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nq, self.k, self.slots = nq_local, k, slots
        dev = torch.device("cuda", torch.cuda.current_device())
        rows = self.world * nq_local
        self.tile = rows * k                       # int32 elements per ids (or dists) buffer
        per_slot = 2 * self.tile + 32              # ids | dists | flags (padded)
        self.buf = symm_mem.empty(slots * per_slot, dtype=torch.int32, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        ptrs = list(self.hdl.buffer_ptrs)
        self.per_slot = per_slot
        self._PeerGather = PeerGather
        self.specs = []
        for sl in range(slots):
            g = PeerGather()
            g.n_peers = self.world
            g.my_rank = self.rank
            g.row_offset = self.rank * nq_local
            for p in range(self.world):
                base = ptrs[p] + sl * per_slot * 4
                g.ids[p] = base
                g.dists[p] = base + self.tile * 4
                g.flags[p] = base + 2 * self.tile * 4
            self.specs.append(g)
        torch.cuda.synchronize()
        dist.barrier(group)

    def spec(self, slot, seq):
        g = self.specs[slot]
        g.seq = seq
        return g

    def ids(self, slot):
        o = slot * self.per_slot
        return self.buf[o:o + self.tile].view(self.world * self.nq, self.k)

    def dists(self, slot):
        import torch

        o = slot * self.per_slot + self.tile
        return self.buf[o:o + self.tile].view(torch.float32).view(self.world * self.nq, self.k)

    def flags(self, slot):
        o = slot * self.per_slot + 2 * self.tile
        return self.buf[o:o + self.world]


class PartitionedGranne:
    """One independent index per contiguous id range (one per rank); all queries searched on every shard; per-shard
    tiles all-gathered and merged by (distance, global id)."""

    def __init__(self, index=None, shard_base=0, group=None, local_search=None, device_index=None, host_merge=False):
        import torch
        import torch.distributed as dist

        self.host_merge = host_merge  # tests only: merge CPU tensors (gloo) with merge_topk_host
        self.dist = dist
        self.group = group
        self.index = index
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local = local_search or self._search_device
        self.device_index = device_index
        bases = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(bases, int(shard_base), group=group)
        else:
            bases = [int(shard_base)]
        self.bases = bases
        self._torch = torch

    def _search_device(self, q, max_search, k):
        ids, dists, _ = self.index.search_batch_device(q, max_search, k)
        return ids, dists

    def search_batch(self, queries, max_search=200, k=10):
        torch = self._torch
        ids, dists = self._local(queries, max_search, k)
        nq = ids.shape[0]
        if self.world > 1:
            flat_ids = torch.empty((self.world * nq, k), dtype=ids.dtype, device=ids.device)
            flat_d = torch.empty((self.world * nq, k), dtype=dists.dtype, device=dists.device)
            self.dist.all_gather_into_tensor(flat_ids, ids.contiguous(), group=self.group)
            self.dist.all_gather_into_tensor(flat_d, dists.contiguous(), group=self.group)
            all_ids, all_d = flat_ids.view(self.world, nq, k), flat_d.view(self.world, nq, k)
        else:
            all_ids, all_d = ids[None], dists[None]
        if all_ids.is_cuda:
            from .api import merge_topk_device

            # merge on the device the tiles live on (one process per GPU: that is this rank's device, whatever
            # ordinal the caller's default device has)
            dev = all_ids.device.index if self.device_index is None else self.device_index
            return merge_topk_device(dev, all_ids, all_d, self.bases)
        if not self.host_merge:
            raise RuntimeError("PartitionedGranne merges on the GPU; CPU tensors are only accepted with "
                               "host_merge=True (host-logic tests over gloo)")
        gi, gd = merge_topk_host(all_ids.numpy().view(np.uint32), all_d.numpy(), self.bases, k)
        return torch.from_numpy(gi), torch.from_numpy(gd)
