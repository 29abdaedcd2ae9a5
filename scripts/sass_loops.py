#!/usr/bin/env python
"""Static SASS summary of one kernel in an object file: instruction mix between BAR.SYNC markers.
usage: sass_loops.py file.o <mangled-substring>"""
import subprocess, sys, re, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
funcs = out.split("Function : ")
sel = [f for f in funcs if sys.argv[2] in f.split("\n")[0]]
f = sel[0]
print(f.split("\n")[0][:150])
region = collections.Counter(); k = 0; n = 0
def flush():
    global region, k, n
    fp64 = sum(v for o, v in region.items() if o in ("DADD", "DMUL", "DFMA", "DSETP"))
    print("region %d: %d instrs, FP64 %d | %s" % (k, n, fp64, " ".join("%s=%d" % (o, v) for o, v in region.most_common(14))))
    region = collections.Counter(); k += 1; n = 0
for line in f.split("\n"):
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
    if not m: continue
    ins = m.group(1).split()
    op = ins[1] if ins[0].startswith("@") else ins[0]
    if op == "NOP": continue
    region[op.split(".")[0]] += 1; n += 1
    if op.startswith("BAR") or op == "EXIT": flush()
flush()
