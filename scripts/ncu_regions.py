#!/usr/bin/env python
"""Split an `ncu --page source --csv` dump of one kernel into code regions delimited by BAR.SYNC
and print, per region, instructions executed and warp-stall samples by reason.
usage: ncu_regions.py src.csv <kernel-substring> [occurrence]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]; occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = [i for i in starts if pat in rows[i][1]][occ]
end = min([j for j in starts if j > sel] + [len(rows)])
hdr = rows[sel + 1]
H = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions = []; cur = dict(n=0, inst=0, samples=0, first=None, **{c: 0 for c in stall_cols}); cur["ops"] = {}
def num(x):
    try: return float(x)
    except: return 0.0
for r in rows[sel + 2:end]:
    if len(r) < len(hdr): continue
    sass = r[H["Source"]].strip()
    if cur["first"] is None: cur["first"] = r[H["Address"]][-5:]
    cur["n"] += 1; cur["inst"] += num(r[H["Instructions Executed"]]); cur["samples"] += num(r[H["# Samples"]])
    op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
    op = op.split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + num(r[H["Instructions Executed"]])
    for c in stall_cols: cur[c] += num(r[H[c]])
    if "BAR.SYNC" in sass or sass.startswith("EXIT"):
        regions.append(cur); cur = dict(n=0, inst=0, samples=0, first=None, **{c: 0 for c in stall_cols}); cur["ops"] = {}
regions.append(cur)
tot_i = sum(x["inst"] for x in regions); tot_s = sum(x["samples"] for x in regions)
print("total inst %.3e samples %d" % (tot_i, tot_s))
for k, x in enumerate(regions):
    if x["inst"] == 0 and x["samples"] == 0: continue
    top = sorted(((x[c], c[6:]) for c in stall_cols), reverse=True)[:6]
    ops = sorted(((v, o) for o, v in x["ops"].items()), reverse=True)[:8]
    print("region %d @%s: %4d sass, inst %5.1f%%, samples %5.1f%% | %s" % (k, x["first"], x["n"], 100 * x["inst"] / tot_i, 100 * x["samples"] / max(tot_s, 1),
          " ".join("%s=%.1f%%" % (c, 100 * v / max(tot_s, 1)) for v, c in top)))
    print("      ops: " + " ".join("%s=%.1f%%" % (o, 100 * v / tot_i) for v, o in ops))
