"""Per-phase instruction/stall budget of the fast search loop from an .ncu-rep, bucketed by SOURCE LINE ranges that
are looked up from marker comments in search_kernels.cuh (so the tool follows the file as it changes).
Usage: python tools/ncu_phase_lines.py rep.ncu-rep n_expansions_total"""
import csv, io, subprocess, sys, re, os
rep = sys.argv[1]; nexp = float(sys.argv[2])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "granne_b200/csrc/search_kernels.cuh")).read().split("\n")
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise SystemExit("marker not found: " + pat)
f0 = find("__device__ __forceinline__ void search_layer_fast(")
marks = [("pop+spec", find("// ---- pq.pop(): first unexpanded", f0)),
         ("res.push/thr", find("// ---- res.push ----", f0)),
         ("adjacency", find("// ---- neighbours ----", f0)),
         ("visited(call)", find("const bool is_new = vis_bucket_insert", f0)),
         ("compact", find("if (is_new) c.ids[", f0)),
         ("dist(call)", find("const float d = dist.dists(ix, c, my_id, k", f0)),
         ("spec buckets (hook)", find("candidate rows are in flight the speculative adjacency row lands too", f0)),
         ("pass filter", find("// !res.is_full() || distance < res.peek().0   (:1029)", f0)),
         ("merge: park+rank", find("// park the passing keys", f0)),
         ("merge: rank_n", find("// rank among the new keys", f0)),
         ("merge: drops", find("uint32_t drop_flagged = 0, mdrop = 0;", f0)),
         ("merge: rows", find("// rebuild rows top-down", f0)),
         ("merge: bookkeeping", find("if (pass && new_pos < cap) {", f0)),
         ("end", find("// ElementContainer::get(id) written into the query slot", f0))]
vis0, vis1 = find("__device__ __forceinline__ bool vis_bucket_insert("), find("// search_for_neighbors (src/index/mod.rs:999-1037) on one layer.")
d0 = find("struct DistF32 {"); d1 = find("// ANGULAR f32, any dim (runtime chunk count")
issue = find("cp_async_wait_all();", d0); part = find("c.tile[b * kTileStride + c.lane] = partial(", d0); osum = find("uint32_t id = 0;", d0)
marks.sort(key=lambda m: m[1])
def phase(ln):
    if vis0 <= ln < vis1: return "visited"
    if d0 <= ln < d1:
        if ln >= osum or (d0 <= ln < find("// Batches of up to stg_rows", d0) and ln >= find("__device__ __forceinline__ float ordered_finish", d0)): return "dist: ordered sum+finish"
        if ln >= part or ln < find("__device__ __forceinline__ float ordered_finish", d0): return "dist: partial"
        if ln >= issue: return "dist: wait"
        return "dist: issue"
    if marks[0][1] <= ln < marks[-1][1]:
        name = marks[0][0]
        for n, l in marks:
            if ln >= l: name = n
        return name
    return "other (kernel frame, helpers, intrinsics headers)"
def page(s):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", s], capture_output=True, text=True).stdout)))
agg = {}; cur = None; hdr = None
for r in page("cuda,sass"):
    if r and r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0].strip().isdigit():
        ln = int(r[0]); p = phase(ln)
        try: a = agg.setdefault(p, [0, 0]); a[0] += int(r[iI]); a[1] += int(r[iS])
        except ValueError: pass
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total %.1f inst/expansion, %d stall samples (all instantiations: upper layers + bottom layer)" % (ti / nexp, ts))
for p, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-52s %6.1f inst/exp %5.1f%%   %5.1f%% stall samples" % (p, i / nexp, 100.0 * i / ti, 100.0 * s / ts))
