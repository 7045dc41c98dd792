"""SASS-level view of an .ncu-rep in address order: executions per expansion, stall samples, source line (hot code only).
Usage: python tools/ncu_sass.py rep.ncu-rep n_expansions_total [min_per_expansion]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; nexp = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
def page(src):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", src], capture_output=True, text=True).stdout)))
a2l = {}
cur = None
for r in page("cuda,sass"):
    if len(r) < 4 or r[0] == "Line No": continue
    if r[0].strip().isdigit(): cur = (int(r[0]), r[1].strip()); continue
    if r[2].startswith("0x") and cur: a2l[r[2]] = cur
rows = page("sass")
hdr = None; tot_i = tot_s = 0; out = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
        iL = hdr.index("stall_long_sb"); iW = hdr.index("stall_wait"); iSh = hdr.index("stall_short_sb"); iB = hdr.index("stall_branch_resolving")
        continue
    if hdr is None or len(r) < len(hdr): continue
    try: inst = int(r[iI]); samp = int(r[iS])
    except ValueError: continue
    tot_i += inst; tot_s += samp
    out.append((r[0], r[1].strip(), inst, samp, r[iL], r[iW], r[iSh], r[iB]))
print("total inst %d (%.1f per expansion), samples %d" % (tot_i, tot_i / nexp, tot_s))
acc = 0.0; last = None
for a, s, inst, samp, l, w, sh, b in out:
    per = inst / nexp
    if per >= thr:
        acc += per
        ln = a2l.get(a)
        if ln and ln != last:
            print("        ---- L%d: %s" % (ln[0], ln[1][:110])); last = ln
        print("%s %6.2f %5.2f%% %s" % (a[-5:], per, 100.0 * samp / tot_s, s))
print("listed: %.1f inst per expansion" % acc)
