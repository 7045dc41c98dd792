"""Buckets an ncu source-page export by phase of search_layer_fast (line ranges are looked up by marker comments)."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
src_path = "granne_b200/csrc/search_kernels.cuh"
lines = open(src_path).read().split("\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; agg = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; iA, iI, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) or r[iA] != "-": continue
    try: ln, inst, samp = int(r[0]), int(r[iI]), int(r[iS])
    except ValueError: continue
    a = agg.setdefault((ln, r[1].strip()[:60]), [0, 0]); a[0] += inst; a[1] += samp
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
# classify by source text (robust to line shifts between the profiled build and the current file)
rules = [("visited", r"ldcg_u4|stcg_u|bucket|found|used =|pending|same_id|same_b|ins_mask|nbuckets|probe|0x9E3779B1|hi\.[xyzw]|lo\.[xyzw]"),
         ("ordered_sum", r"__fadd_rn\(r|const float4 v = t\[i\]|float4\* t ="),
         ("partial_fma", r"__fmaf_rn\(v\.|partial\(|reinterpret_cast<const float4\*>\(r \+"),
         ("tma_issue_wait", r"mbar_|bulk_copy|try_wait|copy_bytes|nb = \(k - j0\)|j0 \+= rb|selp|cp\.async"),
         ("bsearch", r"lo \+ step|step >>= 1|\+\+lo"),
         ("rank_shift", r"rank_n|sh\[tt\]|sh_base|rel <|t &= t - 1|__ffs\(t\)|rj|dj|ij"),
         ("move", r"dv\[t\]|iv\[t\]|Ld\[np\]|Li\[np\]|np = j|drop_flagged|new_pos"),
         ("select_spec", r"sel_mask|base_sel|spec_|have_cur|cur_nb|px"),
         ("thr", r"pos_thr|thr_bits|n_exp"),
         ("adjacency", r"__ldg\(row|valid|vm ==|n_nbr"),
         ("compact", r"c\.ids|is_new|nm"),
         ("pass_filter", r"pass|pm =|pm ==|m = __popc"),
         ("finish", r"d != d|finish_angular|0\.0f <= d|tail|__fsub_rn")]
ph = {}
for (ln, text), (inst, samp) in agg.items():
    name = "other"
    for n, pat in rules:
        if re.search(pat, text): name = n; break
    a = ph.setdefault(name, [0, 0]); a[0] += inst; a[1] += samp
print("total inst %d samples %d" % (ti, ts))
for n, (i, s) in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    print("%-16s %5.1f%% inst  %5.1f%% stall samples" % (n, 100 * i / ti, 100 * s / ts))
