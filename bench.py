#!/usr/bin/env python
"""bench.py — QPS of the granne search path on B200 (BASELINE.json metric), with roofline, e2e and CPU baseline.

Workload (config.workload): BASELINE.json configs[1] — 1M x 128-d angular f32, M=30, build max_search=200, search
max_search=200, k=10, batches of 1024 queries per GPU per step — the largest configuration of the metric's family that
this bench can BUILD inside its time budget (the 100M x 128 configuration of the headline fits one GPU's HBM, but no
100M-element HNSW index can be constructed in minutes).  Synthetic clustered vectors
 so that recall@10 >= 0.95 is reachable; recall is measured against an exact brute force and reported.

One step = one batch of 1024 queries per rank through Granne::search semantics (granne_b200_search_batch*).  Steps
are independent batches; they are issued round-robin on 8 CUDA streams (a serving system would do the same), then
the whole timed region is bracketed by barrier + synchronize and timed with CUDA events (max over ranks).
  value   device-resident: queries already in HBM, results left in HBM (+ NCCL all-gather of the result tiles, N > 1)
  e2e     host buffers through the public API (granne_b200.Granne.search_batch): H2D of the queries and D2H of the
          results inside the timed region, issued from a few host threads
Multi-GPU: the index is replicated (rank 0 builds, the file image is broadcast), every rank searches its own 1024
queries per step ("weak" scaling: per-GPU work fixed), results are all-gathered.

`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on the same configuration.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "QPS @ recall@10>=0.95, angular f32 HNSW search (granne Granne::search), per-box aggregate"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elements", dest="n", type=int, default=1_000_000,
                    help="number of indexed elements (not --n: torchrun would claim that prefix)")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--nq", type=int, default=1024, help="queries per step per GPU")
    ap.add_argument("--max-search", type=int, default=200)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--num-neighbors", type=int, default=30)
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--graphs", type=int, default=0, help="replay each step from a CUDA graph (0 = eager launches)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: p2p = kernels store result tiles into every peer's buffer (fused epilogue over "
                         "NVLink peer memory); nccl = one all-gather per step")
    ap.add_argument("--kind", default="angular", choices=["angular", "angular_int"],
                    help="element type (angular_int = BASELINE config 3 style i8/dp4a path)")
    ap.add_argument("--reorder", type=int, default=0,
                    help="1: run Granne::reorder (GPU compute_order + host apply) on the built index before searching")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the cpu_baseline sample")
    return ap.parse_args()


def clustered(n, dim, seed, n_centers, sub_dim=16, spread=0.3, basis_seed=7):
    """Gaussian mixture on a random low-dimensional subspace (SURVEY.md §8d measurement distribution)."""
    brng = np.random.default_rng(basis_seed)
    basis = brng.standard_normal((sub_dim, dim)).astype(np.float32)
    centers = brng.standard_normal((n_centers, sub_dim)).astype(np.float32)
    rng = np.random.default_rng(seed)
    out = np.empty((n, dim), dtype=np.float32)
    step = 1 << 18
    for s in range(0, n, step):
        m = min(step, n - s)
        which = rng.integers(0, n_centers, size=m)
        pts = centers[which] + spread * rng.standard_normal((m, sub_dim)).astype(np.float32)
        out[s:s + m] = pts @ basis
    return out


def workload_config(a, impl):
    et = "f32" if a.kind == "angular" else "i8"
    return {"workload": "%dx%d angular %s HNSW (M=%d, build max_search=200), search max_search=%d k=%d, "
                        "%d queries/step/GPU" % (a.n, a.dim, et, a.num_neighbors, a.max_search, a.k, a.nq),
            "baseline_config": "BASELINE.json configs[1]" if a.kind == "angular" else "BASELINE.json configs[2] family",
            "n": a.n, "dim": a.dim, "max_search": a.max_search,
            "k": a.k, "queries_per_step_per_gpu": a.nq, "index": "replicated, queries sharded" if a.gpus > 1
            else "single GPU", "l2": "inputs larger than L2 (%.0f MB vectors + adjacency; query batches rotate)"
            % (a.n * a.dim * (4 if a.kind == "angular" else 1) / 1e6), "streams": a.streams, "impl": impl,
            "reordered": bool(getattr(a, "reorder", 0))}


class ClockSampler:
    """nvidia-smi sampled during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(index_bytes, elements_bytes, queries, a, seconds):
    """The CPU restatement of the reference (oracle/) on this box's host cores, bounded sample."""
    from oracle import granne_oracle as go

    threads = os.cpu_count() or 1
    el = go.Elements.from_bytes(a.kind, elements_bytes)
    g = go.Granne.from_bytes(index_bytes, el)          # compressed adjacency decoded per expansion (faithful)
    gf = g.to_fixed()                                  # pre-decoded adjacency (the stronger CPU variant)
    probe = queries[:max(threads * 2, 64)]
    t = time.time()
    gf.search_batch(probe, a.max_search, a.k, threads=threads)
    rate = probe.shape[0] / max(time.time() - t, 1e-6)
    nsample = int(min(queries.shape[0], max(probe.shape[0], rate * seconds / 2)))
    sample = queries[:nsample]
    out = {}
    for name, idx in (("compressed_adjacency", g), ("decoded_adjacency", gf)):
        t = time.time()
        idx.search_batch(sample, a.max_search, a.k, threads=threads)
        out[name] = nsample / max(time.time() - t, 1e-9)
    best = max(out.values())
    return {"value": best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d queries of the bench workload, %d threads, one query per task in static chunks; best of "
                      "compressed (%.0f QPS) and pre-decoded (%.0f QPS) adjacency" %
                      (nsample, threads, out["compressed_adjacency"], out["decoded_adjacency"])}


def run_reference(a):
    """Reference arm: the oracle's CPU search (all host threads) on the same configuration; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import granne_oracle as go

    threads = os.cpu_count() or 1
    n_centers = max(8, int(4096 * (a.n / 1e6) ** 0.5))
    raw = clustered(a.n, a.dim, seed=1234, n_centers=n_centers)
    el = go.Elements.angular(raw) if a.kind == "angular" else go.Elements.angular_int(raw)
    del raw
    t0 = time.time()
    g = go.GranneBuilder(el, num_neighbors=a.num_neighbors, max_search=200).build(threads=threads).to_fixed()
    build_s = time.time() - t0
    queries = clustered(a.nq * 16, a.dim, seed=4321, n_centers=n_centers)
    probe = queries[:max(threads * 2, 64)]
    t = time.time()
    g.search_batch(probe, a.max_search, a.k, threads=threads)
    rate = probe.shape[0] / max(time.time() - t, 1e-6)
    # bounded sample per step: ~1 s, and at most ~60 s for the whole --steps/--warmup run
    per_step_s = min(1.0, 60.0 / max(1, a.steps + a.warmup))
    per_step = int(max(64, min(a.nq, rate * per_step_s)))
    for _ in range(a.warmup):
        g.search_batch(queries[:per_step], a.max_search, a.k, threads=threads)
    t0 = time.time()
    for s in range(a.steps):
        off = (s * per_step) % (queries.shape[0] - per_step + 1)
        g.search_batch(queries[off:off + per_step], a.max_search, a.k, threads=threads)
    dt = time.time() - t0
    qps = a.steps * per_step / dt
    line = {"metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if a.kind == "angular" else "i8", "data": "synthetic", "impl": "reference", "config": workload_config(a, "reference"),
            "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d queries per step (bounded), pre-decoded adjacency, %d threads; index built "
                                       "by the oracle's threaded builder in %.0f s" % (per_step, threads, build_s)},
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist

    import granne_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the granne_b200 search path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from granne_b200 import build as gb_build

    if rank == 0:
        gb_build.build()
    if world > 1:
        dist.barrier()
    granne_b200.load_library()

    # ---- synthetic data, elements, index ---------------------------------------------------------------------------
    n_centers = max(8, int(4096 * (a.n / 1e6) ** 0.5))
    t0 = time.time()
    raw = clustered(a.n, a.dim, seed=1234, n_centers=n_centers)
    elements_bytes = granne_b200.elements_from_raw(a.kind, raw, device=local)
    del raw
    t_data = time.time() - t0
    t0 = time.time()
    if rank == 0:
        builder = granne_b200.GranneBuilder(a.kind, elements_bytes, num_neighbors=a.num_neighbors, max_search=200,
                                            device=local)
        builder.build()
        # very large single-GPU runs skip the host-side file image (only needed for the CPU baseline / replication)
        big = a.n > 20_000_000 and world == 1
        index_bytes = None if big else builder.index_bytes()
        if world == 1:
            index = builder.get_index()
        builder_launches = 0
    t_build = time.time() - t0
    t_reorder = None
    if a.reorder and world == 1 and index_bytes is not None:
        t0 = time.time()
        index.reorder()
        t_reorder = time.time() - t0
        index_bytes = np.frombuffer(index.index_bytes(), dtype=np.uint8)
        elements_bytes = np.frombuffer(index.elements_bytes(), dtype=np.uint8)
    if world > 1:
        size = torch.tensor([len(index_bytes) if rank == 0 else 0], dtype=torch.int64, device=dev)
        dist.broadcast(size, 0)
        buf = torch.empty(int(size.item()), dtype=torch.uint8, device=dev)
        if rank == 0:
            buf.copy_(torch.from_numpy(np.asarray(index_bytes)))
        dist.broadcast(buf, 0)
        index_bytes = buf.cpu().numpy()
        del buf
        if rank == 0:
            builder.close()
        index = granne_b200.Granne.from_bytes(index_bytes, a.kind, elements_bytes, device=local)
    elif rank == 0:
        builder.close()

    # ---- queries: a rotating pool of distinct batches per rank --------------------------------------------------------
    pool = 16
    q_host = clustered(a.nq * pool, a.dim, seed=4321 + rank, n_centers=n_centers)
    q_dev = torch.from_numpy(q_host).to(dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, a.streams))]
    # per stream: one int32 buffer [2, nq, k] = ids | distance bits (a single all-gather collects both), + counts
    bufs = [torch.empty((2, a.nq, a.k), dtype=torch.int32, device=dev) for _ in streams]
    outs = [(b[0], b[1].view(torch.float32), torch.empty((a.nq,), dtype=torch.int32, device=dev)) for b in bufs]
    gathered, groups, fused = None, None, None
    if world > 1 and a.gather == "p2p":
        try:
            from granne_b200.distributed import FusedGather

            fused = FusedGather(a.nq, a.k, slots=len(streams))
        except Exception as e:  # symmetric memory unavailable: use the NCCL all-gather
            fused = None
            if rank == 0:
                print("fused peer gather unavailable (%r): falling back to NCCL all-gather" % (e,), file=sys.stderr)
        ok = torch.tensor([1 if fused is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            fused = None
    if world > 1 and fused is None:
        gathered = [torch.empty((world * 2 * a.nq, a.k), dtype=torch.int32, device=dev) for _ in streams]
        # one communicator per stream: collectives of different in-flight steps do not serialise behind each other
        groups = [dist.new_group(backend="nccl") for _ in streams]

    qin = [torch.empty((a.nq, a.dim), dtype=torch.float32, device=dev) for _ in streams]
    graphs = [None] * len(streams)

    seq = [0]

    def step_body(slot):
        if fused is not None:  # result tiles go straight into every peer's gathered buffer
            seq[0] += 1
            index.search_batch_device_gather(qin[slot], fused.spec(slot, seq[0]), a.max_search, a.k,
                                             stream=streams[slot].cuda_stream)
            return
        index.search_batch_device(qin[slot], a.max_search, a.k, out=outs[slot], stream=streams[slot].cuda_stream)
        if world > 1:  # collect every rank's result tile (NCCL all-gather over NVLink)
            dist.all_gather_into_tensor(gathered[slot], bufs[slot].view(2 * a.nq, a.k), group=groups[slot])

    def device_step(s):
        slot = s % len(streams)
        st = streams[slot]
        with torch.cuda.stream(st):
            qin[slot].copy_(q_dev[(s % pool) * a.nq:(s % pool + 1) * a.nq], non_blocking=True)
            if graphs[slot] is not None:
                graphs[slot].replay()
            else:
                step_body(slot)

    def capture_graphs():
        """One CUDA graph per stream slot: search kernels (+ the all-gather) replayed with a single launch."""
        for slot, st in enumerate(streams):
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    step_body(slot)
                graphs[slot] = g
            except Exception as e:  # capture unsupported (e.g. NCCL build): fall back to eager launches
                graphs[slot] = None
                if rank == 0:
                    print("cuda graph capture failed, running eagerly: %r" % (e,), file=sys.stderr)
                break

    def sync_all():
        for st in streams:
            st.synchronize()
        torch.cuda.synchronize(dev)

    # recall + algorithmic bytes per query (counters are parity-checked against the oracle in tests/)
    stats = torch.zeros((a.nq, 4), dtype=torch.int64, device=dev)
    ids0, _, _ = index.search_batch_device(q_dev[:a.nq], a.max_search, a.k, stats=stats)
    torch.cuda.synchronize(dev)
    index.stream_status()
    st_np = stats.cpu().numpy()
    n_dist, n_expand, n_nbr = st_np[:, 0].mean(), st_np[:, 1].mean(), st_np[:, 2].mean()
    esz = 4 if a.kind == "angular" else 1
    bytes_per_query = n_dist * a.dim * esz + n_nbr * 4 + a.dim * esz  # vectors + adjacency + query
    nsamp = min(256, a.nq)
    if a.kind == "angular":
        rows = torch.from_numpy(np.frombuffer(elements_bytes, dtype=np.float32, offset=8).reshape(a.n, a.dim))
        qn = torch.nn.functional.normalize(q_dev[:nsamp], dim=1)
    else:  # ground truth under the same i8 angular distance: cosine of the quantised vectors
        rows = torch.from_numpy(np.frombuffer(elements_bytes, dtype=np.int8, offset=8).reshape(a.n, a.dim))
        qq = q_dev[:nsamp]
        qq = torch.trunc(qq * 127.0 / qq.abs().amax(dim=1, keepdim=True))
        qn = torch.nn.functional.normalize(qq, dim=1)
    best = None
    for s0 in range(0, a.n, 1 << 18):  # exact brute force in slabs (off the hot path)
        blk = rows[s0:s0 + (1 << 18)].to(dev)
        if a.kind != "angular":
            blk = torch.nn.functional.normalize(blk.float(), dim=1)
        sc = qn @ blk.T
        v, i = torch.topk(sc, a.k, dim=1)
        i = i + s0
        if best is None:
            best = (v, i)
        else:
            vv = torch.cat([best[0], v], dim=1)
            ii = torch.cat([best[1], i], dim=1)
            tv, ti = torch.topk(vv, a.k, dim=1)
            best = (tv, torch.gather(ii, 1, ti))
        del blk, sc
    gt = best[1].cpu().numpy()
    got = ids0[:nsamp].cpu().numpy()
    recall = float(np.mean([len(set(gt[i].tolist()) & set(got[i].tolist())) / a.k for i in range(nsamp)]))
    del rows

    # ---- device-resident timed region ------------------------------------------------------------------------------------
    for s in range(max(a.warmup, len(streams))):
        device_step(s)
    sync_all()
    if a.graphs and fused is None:
        capture_graphs()
        sync_all()
        for s in range(len(streams)):
            device_step(s)
        sync_all()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = index.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.nvtx.range_push("timed")
    e0.record(torch.cuda.current_stream(dev))
    for st in streams:
        st.wait_stream(torch.cuda.current_stream(dev))
    t_issue = time.perf_counter()
    for s in range(a.steps):
        device_step(a.warmup + s)
    issue_ms = (time.perf_counter() - t_issue) * 1e3
    for st in streams:
        torch.cuda.current_stream(dev).wait_stream(st)
    e1.record(torch.cuda.current_stream(dev))
    sync_all()
    torch.cuda.nvtx.range_pop()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    launches = index.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    index.stream_status()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = a.steps * a.nq * world / (ms / 1e3)

    # kernel-alone duration (single stream, one launch at a time) for the per-launch roofline
    solo = []
    for s in range(6):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        qb = q_dev[(s % pool) * a.nq:(s % pool + 1) * a.nq]
        k0.record()
        index.search_batch_device(qb, a.max_search, a.k, out=outs[0])
        k1.record()
        torch.cuda.synchronize(dev)
        solo.append(k0.elapsed_time(k1))
    solo_ms = float(np.median(solo[1:]))

    # ---- end to end through the public host API ------------------------------------------------------------------------
    nthreads = max(1, min(a.streams, 8))
    h2d = a.nq * a.dim * 4
    d2h = a.nq * a.k * 8 + a.nq * 4

    def host_steps(t, count, offset):
        for s in range(count):
            b = (offset + s * nthreads + t) % pool
            index.search_batch(q_host[b * a.nq:(b + 1) * a.nq], a.max_search, a.k)

    def run_host(count_total, offset):
        per = [count_total // nthreads + (1 if t < count_total % nthreads else 0) for t in range(nthreads)]
        ths = [threading.Thread(target=host_steps, args=(t, per[t], offset)) for t in range(nthreads)]
        [t.start() for t in ths]
        [t.join() for t in ths]

    run_host(max(a.warmup, nthreads), 0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    run_host(a.steps, 3)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_qps = a.steps * a.nq * world / e2e_s

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    total_bytes = bytes_per_query * a.nq * a.steps  # this rank's launches in the timed region
    achieved = total_bytes / (ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            if (tj.get("kind"), tj.get("n"), tj.get("dim"), tj.get("queries_per_launch")) == (a.kind, a.n, a.dim, a.nq):
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    cpu = None
    if world == 1 and index_bytes is not None and a.cpu_seconds > 0:
        cpu = cpu_baseline(np.asarray(index_bytes).tobytes(), np.asarray(elements_bytes).tobytes(), q_host, a,
                           a.cpu_seconds)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if a.kind == "angular" else "i8", "data": "synthetic", "config": workload_config(a, "ours"),
        "recall_at_10": recall,
        "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_threads": nthreads, "api": "granne_b200.Granne.search_batch (granne_b200_search_batch)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": "search_kernel<%s,7> (1 launch per step)" % ("DistF32<4>" if a.kind == "angular" else "DistI8"),
                     "algorithmic_bytes_per_query": bytes_per_query,
                     "algorithmic_bytes_per_launch": bytes_per_query * a.nq,
                     "n_dist_per_query": n_dist, "n_expand_per_query": n_expand,
                     "solo_launch_ms": solo_ms,
                     "solo_launch_gbs": bytes_per_query * a.nq / (solo_ms / 1e3) / 1e9},
        "cpu_baseline": cpu,
        "setup_s": {"data+elements": t_data, "gpu_index_build": t_build, "reorder": t_reorder},
        "host_issue_ms_per_step": issue_ms / a.steps, "cuda_graphs": bool(graphs[0] is not None),
        "multi_gpu_gather": None if world == 1 else ("p2p peer stores fused into the search kernels" if fused is not None
                                                     else "nccl all_gather"),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
