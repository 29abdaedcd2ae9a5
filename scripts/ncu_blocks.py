#!/usr/bin/env python
"""Basic-block view of an `ncu --page source --csv` dump: contiguous SASS with equal execution
count, with share of executed instructions, stall samples and opcode mix.
usage: ncu_blocks.py src.csv <kernel-substring> [occurrence] [min-share-%]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]; occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
minshare = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sel = [i for i in starts if pat in rows[i][1]][occ]
end = min([j for j in starts if j > sel] + [len(rows)])
hdr = rows[sel + 1]; H = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[sel + 2:end] if len(r) >= len(hdr)]
tot = sum(float(r[H["Instructions Executed"]]) for r in body)
tsm = sum(float(r[H["# Samples"]]) for r in body)
blk = []; cur = None; nbar = 0
for r in body:
    n = float(r[H["Instructions Executed"]]); sass = r[1].strip()
    if cur is None or n != cur["n"]:
        if cur: blk.append(cur)
        cur = dict(n=n, first=r[0][-5:], cnt=0, ops={}, smp=0, bar=nbar, st={c: 0 for c in stall_cols})
    cur["cnt"] += 1; cur["smp"] += float(r[H["# Samples"]])
    for c in stall_cols: cur["st"][c] += float(r[H[c]] or 0)
    op = (sass.split()[1] if sass.startswith("@") else sass.split()[0]).split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + 1
    if "BAR.SYNC" in sass: nbar += 1
blk.append(cur)
print("total inst %.4e, samples %d, sass %d" % (tot, tsm, len(body)))
for b in blk:
    share = b["n"] * b["cnt"] / tot * 100
    if share > minshare:
        ops = sorted(b["ops"].items(), key=lambda x: -x[1])[:7]
        st = sorted(((v, c[6:]) for c, v in b["st"].items()), reverse=True)[:3]
        print("bar%d @%s n=%.3e x %4d = %5.2f%% inst, %5.2f%% smp | %s | %s" % (
            b["bar"], b["first"], b["n"], b["cnt"], share, b["smp"] / tsm * 100,
            " ".join("%s:%d" % o for o in ops), " ".join("%s=%.1f%%" % (c, 100 * v / tsm) for v, c in st)))
