#!/usr/bin/env python
"""Generates tests/golden/reference_traces.json by RUNNING THE UNMODIFIED REFERENCE
(baseline/_ref/serial/bin/athenaPK, Kokkos Serial, built by baseline/build_reference.sh) on the
reference's own input decks (/root/reference/inputs, build container only) with the CLI overrides
recorded per case below.  What is kept: Parthenon's per-cycle `time` / `dt` printout at all 17
digits (driver.cpp:301-306) -- dt is a minimum over the whole state, i.e. a sensitive fingerprint
of every cell -- and the rows the linear-wave problem generators append to linearwave-errors.dat.

  python tests/golden/make_reference_goldens.py          (about 5 minutes on 8 cores)

The parity tests (CPU oracle here, CUDA path on the GPU box) read only the JSON."""
import json
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "baseline", "_ref", "serial", "bin", "athenaPK")
INPUTS = "/root/reference/inputs"
CYCLE = re.compile(r"cycle=(\d+) time=(\S+) dt=(\S+)")
GAMMA = "1.666666666666667"  # the decks' literal (NOT 5/3 in binary64)


def run(deck, overrides):
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [BIN, "-i", deck] + overrides
        p = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError(" ".join(cmd) + "\n" + p.stdout[-2000:])
        trace = [[m.group(2), m.group(3)] for m in CYCLE.finditer(p.stdout)]
        err = None
        f = os.path.join(tmp, "linearwave-errors.dat")
        if os.path.exists(f):
            err = open(f).read().strip().splitlines()[-1].split()
    return {"overrides": overrides, "ncycle": len(trace) - 1, "trace": trace, "lw_row": err}


def lw_case(problem, fluid, integ, recon, rs, ng, N):
    ov = ["job/problem_id=" + problem, "hydro/fluid=" + fluid,
          "parthenon/mesh/nx1=%d" % (2 * N), "parthenon/mesh/nx2=%d" % N, "parthenon/mesh/nx3=%d" % N,
          "parthenon/meshblock/nx1=%d" % (2 * N), "parthenon/meshblock/nx2=%d" % N,
          "parthenon/meshblock/nx3=%d" % N, "parthenon/mesh/nghost=%d" % ng,
          "parthenon/time/integrator=" + integ, "hydro/reconstruction=" + recon,
          "hydro/riemann=" + rs, "parthenon/output0/file_type=hst"]
    r = run(os.path.join(INPUTS, "linear_wave3d.in"), ov)
    r.update(problem=problem, fluid=fluid, integrator=integ, recon=recon, riemann=rs, ng=ng, N=N,
             gamma=GAMMA)
    return r


LW_METHODS = [("euler", "rk2", "plm", "hlle", 2), ("euler", "vl2", "plm", "hlle", 2),
              ("euler", "vl2", "plm", "hllc", 2), ("euler", "rk3", "ppm", "hlle", 3),
              ("euler", "rk3", "wenoz", "hlle", 3), ("glmmhd", "vl2", "plm", "hlle", 2),
              ("glmmhd", "vl2", "ppm", "hlld", 3), ("glmmhd", "rk3", "wenoz", "hlld", 3)]
LWM_METHODS = [("glmmhd", "vl2", "ppm", "hlld", 3), ("glmmhd", "rk3", "wenoz", "hlld", 3),
               ("glmmhd", "vl2", "plm", "hlle", 2), ("glmmhd", "rk2", "limo3", "hlld", 3)]


def main():
    jobs = {}
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for N in (16, 32):
            for m in LW_METHODS:
                jobs[("linear_wave", N) + m] = ex.submit(lw_case, "linear_wave", *m, N)
            for m in LWM_METHODS:
                jobs[("linear_wave_mhd", N) + m] = ex.submit(lw_case, "linear_wave_mhd", *m, N)
        # magnetised wave families: Alfven (1) and slow (2) at N = 16 (fast = 0 is the table above; the entropy wave has zero speed)
        for wf in (1, 2):
            jobs[("linear_wave_mhd_wf", wf)] = ex.submit(
                lambda wf=wf: dict(run(os.path.join(INPUTS, "linear_wave3d.in"), [
                    "job/problem_id=linear_wave_mhd", "hydro/fluid=glmmhd", "parthenon/mesh/nx1=32",
                    "parthenon/mesh/nx2=16", "parthenon/mesh/nx3=16", "parthenon/meshblock/nx1=32",
                    "parthenon/meshblock/nx2=16", "parthenon/meshblock/nx3=16",
                    "parthenon/mesh/nghost=3", "parthenon/time/integrator=vl2",
                    "hydro/reconstruction=ppm", "hydro/riemann=hlld",
                    "problem/linear_wave/wave_flag=%d" % wf, "parthenon/output0/file_type=hst"]),
                    wave_flag=wf, gamma=GAMMA))
        jobs["sod"] = ex.submit(run, os.path.join(INPUTS, "sod.in"), ["hydro/riemann=hllc", "parthenon/output0/file_type=hst"])
        jobs["orszag_tang_50"] = ex.submit(run, os.path.join(INPUTS, "orszag_tang.in"),
                                           ["parthenon/time/nlim=50", "parthenon/output0/file_type=hst",
                                            "parthenon/output1/file_type=hst"])
        # FirstOrderFluxCorrect TRIGGERED (hydro.cpp:1223-1342): a strong 2-D blast run at an
        # unstable CFL number with floors, so that cells fail the trial update every few cycles
        jobs["blast_fofc"] = ex.submit(run, os.path.join(INPUTS, "blast_3d_amr.in"), [
            "parthenon/mesh/refinement=none", "parthenon/mesh/nx1=64", "parthenon/mesh/nx2=64",
            "parthenon/mesh/nx3=1", "parthenon/meshblock/nx1=32", "parthenon/meshblock/nx2=32",
            "parthenon/meshblock/nx3=1", "parthenon/mesh/nghost=3", "parthenon/time/cfl=0.9",
            "parthenon/time/nlim=40", "hydro/reconstruction=ppm", "hydro/riemann=hllc",
            "hydro/first_order_flux_correct=true", "hydro/dfloor=1e-8", "hydro/pfloor=1e-8",
            "problem/blast/radius_outer=0.1", "problem/blast/radius_inner=0.1",
            "parthenon/output0/file_type=hst"])
        # BASELINE config 5 on a FROZEN refined mesh: inputs/blast_3d_amr.in with refinement = static
        # (two nested <parthenon/static_refinement> regions -> 3 levels, 56 + 56 + 64 meshblocks of
        # 8^3), GLM-MHD WENOZ + HLLD, nghost = 4 (odd nghost is rejected with mesh refinement,
        # mesh_refinement.cpp:60-70), blast centred off the origin so that it straddles the level
        # boundaries: prolongation, restriction and flux correction all act on the dt-limiting cells
        def amr(name, regions, fluid, recon, rs, ng):
            ov = ["parthenon/mesh/refinement=static", "hydro/fluid=" + fluid,
                  "hydro/reconstruction=" + recon, "hydro/riemann=" + rs,
                  "parthenon/mesh/nghost=%d" % ng, "parthenon/time/nlim=20",
                  "parthenon/output0/file_type=hst", "problem/blast/radius_outer=0.12",
                  "problem/blast/radius_inner=0.12", "problem/blast/pressure_ratio=100",
                  "problem/blast/pressure_ambient=0.1", "problem/blast/x1_0=0.22",
                  "problem/blast/x2_0=0.1", "problem/blast/x3_0=-0.05"]
            for r, (lev, lo, hi) in enumerate(regions):
                for d in range(3):
                    ov += ["parthenon/static_refinement%d/x%dmin=%r" % (r, d + 1, lo),
                           "parthenon/static_refinement%d/x%dmax=%r" % (r, d + 1, hi)]
                ov += ["parthenon/static_refinement%d/level=%d" % (r, lev)]
            jobs[name] = ex.submit(run, os.path.join(INPUTS, "blast_3d_amr.in"), ov)
        amr("amr_2level_plm_hlle", [(1, -0.2, 0.2)], "euler", "plm", "hlle", 2)
        amr("amr_3level_plm_hlle", [(1, -0.2, 0.2), (2, -0.1, 0.1)], "euler", "plm", "hlle", 2)
        amr("amr_3level_wenoz_hlld", [(1, -0.2, 0.2), (2, -0.1, 0.1)], "glmmhd", "wenoz", "hlld", 4)
        # the two same-input decks of the GPU-vs-GPU measurement (baseline/decks), small meshes
        for deck in ("uniform_mhd", "orszag_tang_3d"):
            jobs[deck] = ex.submit(run, os.path.join(ROOT, "baseline", "decks", deck + ".in"),
                                   ["parthenon/time/nlim=12"])
    out = {"_provenance": "printed by the UNMODIFIED reference (Kokkos Serial, g++ -O3, no FMA) run by "
                          "tests/golden/make_reference_goldens.py; trace = [time, dt] per cycle at 17 digits",
           "linear_wave": [], "linear_wave_mhd": [], "linear_wave_mhd_families": []}
    for k, f in jobs.items():
        r = f.result()
        if isinstance(k, tuple) and k[0] in ("linear_wave", "linear_wave_mhd"):
            out[k[0]].append(r)
        elif isinstance(k, tuple):
            out["linear_wave_mhd_families"].append(r)
        else:
            out[k] = r
        print(k, r["ncycle"], r["trace"][-1], (r.get("lw_row") or [None] * 5)[4], flush=True)
    json.dump(out, open(os.path.join(HERE, "reference_traces.json"), "w"), indent=0)


# tests/golden/reference_orszag_tang_50.hst: the history file the reference writes for
#   athenaPK -i inputs/orszag_tang.in parthenon/time/nlim=50 parthenon/output0/file_type=hst \
#            parthenon/output0/dt=-1 parthenon/output1/dt=0.005
# (copied as written: parthenon.out1.hst)


if __name__ == "__main__":
    sys.exit(main())
