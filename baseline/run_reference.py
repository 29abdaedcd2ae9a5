#!/usr/bin/env python
"""Run the UNMODIFIED reference binary (baseline/_ref/{serial,cuda_sm100}/bin/athenaPK, built by
baseline/build_reference.sh from /root/reference) on one of the decks in baseline/decks/ and parse
Parthenon's own per-cycle report (driver.cpp:301-306: cycle, time, dt at 17 digits and
zone-cycles/wsec_step) and the final `zone-cycles/wallsecond`.

  python baseline/run_reference.py --impl cuda --deck uniform_mhd --n 256 --mb 32 --nlim 12
  python baseline/run_reference.py --impl serial --deck orszag_tang_3d --n 64 --copies 8

Nothing here reads /root/reference: the binaries and the decks travel with the repo.  Used by
bench.py (--impl reference, cpu_baseline) and by the GPU-vs-GPU measurement in profiles/."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = {"serial": os.path.join(HERE, "_ref", "serial", "bin", "athenaPK"),
       "cuda": os.path.join(HERE, "_ref", "cuda_sm100", "bin", "athenaPK")}
CYCLE = re.compile(r"cycle=(\d+) time=(\S+) dt=(\S+) zone-cycles/wsec_step=(\S+) wsec_total=(\S+) "
                   r"wsec_step=(\S+)")


def available(impl):
    return os.path.isfile(BIN[impl]) and os.access(BIN[impl], os.X_OK)


def command(impl, deck, n, mb, nlim, extra=()):
    n = n if isinstance(n, (tuple, list)) else (n, n, n)
    mb = mb if isinstance(mb, (tuple, list)) else (mb, mb, mb)
    args = [BIN[impl], "-i", os.path.join(HERE, "decks", deck + ".in")]
    for d in range(3):
        args += ["parthenon/mesh/nx%d=%d" % (d + 1, n[d]), "parthenon/meshblock/nx%d=%d" % (d + 1, mb[d])]
    args += ["parthenon/time/nlim=%d" % nlim]
    return args + list(extra)


def parse(text):
    cyc = []
    for m in CYCLE.finditer(text):
        cyc.append({"cycle": int(m.group(1)), "time": float(m.group(2)), "dt": float(m.group(3)),
                    "time_str": m.group(2), "dt_str": m.group(3), "zc_step": float(m.group(4)),
                    "wsec_step": float(m.group(6))})
    tot = re.search(r"zone-cycles/wallsecond = (\S+)", text)
    return {"cycles": cyc, "zone_cycles_per_wallsecond": float(tot.group(1)) if tot else None}


def run(impl, deck, n, mb, nlim, extra=(), cpu=None, timeout=3600):
    """one run in a scratch directory (the binary writes .hst files into its cwd)"""
    cmd = command(impl, deck, n, mb, nlim, extra)
    if cpu is not None:
        cmd = ["taskset", "-c", str(cpu)] + cmd
    with tempfile.TemporaryDirectory() as tmp:
        t0 = time.time()
        p = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                           timeout=timeout)
        wall = time.time() - t0
    out = parse(p.stdout)
    out.update(rc=p.returncode, wall_s=wall, tail=p.stdout[-1500:] if p.returncode else "")
    return out


def steady_rate(res, skip):
    """zone-cycles/s over the cycles after the first `skip` (Parthenon's perf_cycle_offset idea,
    driver.cpp:146-149): total zones of those cycles / their wall time"""
    cs = [c for c in res["cycles"] if c["cycle"] > skip and c["wsec_step"] > 0]
    if not cs:
        return None
    zones = sum(c["zc_step"] * c["wsec_step"] for c in cs)
    return zones / sum(c["wsec_step"] for c in cs)


def run_copies(deck, n, mb, nlim, copies, skip):
    """`copies` independent serial runs pinned to cores 0..copies-1 (Kokkos Serial is single
    threaded; there is no OpenMP or MPI build here): aggregate zone-cycles/s"""
    procs = []
    tmps = []
    for c in range(copies):
        tmp = tempfile.TemporaryDirectory()
        tmps.append(tmp)
        cmd = ["taskset", "-c", str(c)] + command("serial", deck, n, mb, nlim)
        procs.append(subprocess.Popen(cmd, cwd=tmp.name, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate()[0] for p in procs]
    for t in tmps:
        t.cleanup()
    rates = [steady_rate(parse(o), skip) or 0.0 for o in outs]
    return sum(rates), rates


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="serial", choices=["serial", "cuda"])
    ap.add_argument("--deck", default="uniform_mhd")
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--mb", type=int, default=32)
    ap.add_argument("--nlim", type=int, default=12)
    ap.add_argument("--skip", type=int, default=2)
    ap.add_argument("--copies", type=int, default=1)
    ap.add_argument("extra", nargs="*")
    a = ap.parse_args()
    if not available(a.impl):
        print(json.dumps({"unavailable": BIN[a.impl] + " not built (baseline/build_reference.sh)"}))
        return 0
    if a.copies > 1:
        agg, rates = run_copies(a.deck, a.n, a.mb, a.nlim, a.copies, a.skip)
        print(json.dumps({"impl": a.impl, "deck": a.deck, "n": a.n, "mb": a.mb, "copies": a.copies,
                          "zone_cycles_per_s_aggregate": agg, "per_copy": rates}))
        return 0
    res = run(a.impl, a.deck, a.n, a.mb, a.nlim, a.extra)
    res.update(impl=a.impl, deck=a.deck, n=a.n, mb=a.mb,
               zone_cycles_per_s_steady=steady_rate(res, a.skip))
    print(json.dumps(res))
    return 0 if res["rc"] == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
