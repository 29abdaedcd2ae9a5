#!/usr/bin/env python
"""tests/golden/reference_traces_lw_implode.json: the UNMODIFIED reference (baseline/_ref/serial) on
inputs/lw_implode.in -- the deck of the reference's regression suite case lw_implode_symmetry
(tst/regression/CMakeLists.txt:73-74; Liska-Wendroff implosion, 256^2 in one meshblock, reflecting
walls, VL2 + PLM + HLLC, to t = 2.5).  Kept: Parthenon's (time, dt) printout at 17 digits for
  * a 64^2 version of the deck, every one of 300 cycles (oracle test on the CPU),
  * the deck itself, every 500th cycle and the last one (about 12 minutes of the serial build).
  python tests/golden/make_lw_implode_goldens.py"""
import json
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "baseline", "_ref", "serial", "bin", "athenaPK")
DECK = "/root/reference/inputs/lw_implode.in"
CYCLE = re.compile(r"cycle=(\d+) time=(\S+) dt=(\S+)")
COMMON = ["parthenon/output0/file_type=hst", "parthenon/output0/dt=-1"]


def run(overrides):
    with tempfile.TemporaryDirectory() as tmp:
        p = subprocess.run([BIN, "-i", DECK] + overrides + COMMON, cwd=tmp, stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True, check=True)
    rows, seen = [], set()
    for m in CYCLE.finditer(p.stdout):  # (PostExecute prints the last cycle line a second time)
        if int(m.group(1)) not in seen:
            seen.add(int(m.group(1)))
            rows.append([int(m.group(1)), m.group(2), m.group(3)])
    return rows


def main():
    small = run(["parthenon/mesh/nx1=64", "parthenon/mesh/nx2=64", "parthenon/meshblock/nx1=64",
                 "parthenon/meshblock/nx2=64", "parthenon/time/nlim=300"])
    full = run([])
    out = {"_provenance": "printed by the unmodified reference, tests/golden/make_lw_implode_goldens.py",
           "n64": {"trace": [[t, dt] for _, t, dt in small], "ncycle": small[-1][0]},
           "n256": {"ncycle": full[-1][0],
                    "samples": [r for r in full if r[0] % 500 == 0 or r[0] == full[-1][0]]}}
    json.dump(out, open(os.path.join(HERE, "reference_traces_lw_implode.json"), "w"), indent=0)
    print(out["n64"]["ncycle"], out["n256"]["ncycle"], out["n256"]["samples"][-1])


if __name__ == "__main__":
    main()
