#!/usr/bin/env python
"""Generates tests/golden/reference_traces_diffusion.json by RUNNING THE UNMODIFIED REFERENCE
(baseline/_ref/serial/bin/athenaPK, Kokkos Serial) with the <diffusion> block switched on
(src/hydro/diffusion/*, hydro_driver.cpp:36-344): unsplit and RKL2 super-time-stepping
integration of isotropic / anisotropic thermal conduction, isotropic viscosity and Ohmic
resistivity (fixed coefficients) on the reference's own Orszag-Tang, blast and Sod decks.
Kept: Parthenon's per-cycle `time` / `dt` printout at all 17 digits and the number of RKL2
stages the reference reports for every super step ("STS ratio: ... Taking N steps.").

  python tests/golden/make_diffusion_goldens.py          (about 1 minute)

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
STS = re.compile(r"STS ratio: \S+ Taking (\d+) steps")


def run(deck, overrides):
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [BIN, "-i", deck] + overrides
        p = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError(" ".join(cmd) + "\n" + p.stdout[-2000:])
        trace = [[m.group(2), m.group(3)] for m in CYCLE.finditer(p.stdout)]
        sts = [int(m.group(1)) for m in STS.finditer(p.stdout)]
    return {"overrides": overrides, "ncycle": len(trace) - 1, "trace": trace, "s_rkl": sts}


def diff(integrator, cond=None, visc=None, ohm=None, ratio=None, cfl=None):
    ov = ["diffusion/integrator=" + integrator]
    if cond:
        ov += ["diffusion/conduction=" + cond[0], "diffusion/conduction_coeff=fixed",
               "diffusion/thermal_diff_coeff_code=%r" % cond[1]]
    if visc:
        ov += ["diffusion/viscosity=isotropic", "diffusion/viscosity_coeff=fixed",
               "diffusion/mom_diff_coeff_code=%r" % visc]
    if ohm:
        ov += ["diffusion/resistivity=ohmic", "diffusion/resistivity_coeff=fixed",
               "diffusion/ohm_diff_coeff_code=%r" % ohm]
    if ratio:
        ov += ["diffusion/rkl2_max_dt_ratio=%r" % ratio]
    if cfl:
        ov += ["diffusion/cfl=%r" % cfl]
    return ov


def mesh(n, mb, ng):
    ov = ["parthenon/mesh/nghost=%d" % ng, "parthenon/output0/file_type=hst"]
    for d in range(3):
        ov += ["parthenon/mesh/nx%d=%d" % (d + 1, n[d]), "parthenon/meshblock/nx%d=%d" % (d + 1, mb[d])]
    return ov


OT = os.path.join(INPUTS, "orszag_tang.in")
OT3 = os.path.join(ROOT, "baseline", "decks", "orszag_tang_3d.in")
BLAST = os.path.join(INPUTS, "blast_3d_amr.in")
NOFOFC = ["hydro/first_order_flux_correct=false", "parthenon/output1/file_type=hst"]
CASES = {
    # 2-D GLM-MHD, all three processes, hyperbolic dt limits (small coefficients)
    "ot2d_unsplit": (OT, mesh((64, 64, 1), (32, 32, 1), 3) + NOFOFC + ["parthenon/time/nlim=20"] +
                     diff("unsplit", ("isotropic", 0.002), 0.002, 0.002)),
    # ... diffusive dt limits (large coefficients)
    "ot2d_unsplit_stiff": (OT, mesh((64, 64, 1), (32, 32, 1), 3) + NOFOFC + ["parthenon/time/nlim=12"] +
                           diff("unsplit", ("isotropic", 0.03), 0.05, 0.04)),
    # RKL2, several stages per super step
    "ot2d_rkl2": (OT, mesh((64, 64, 1), (32, 32, 1), 3) + NOFOFC + ["parthenon/time/nlim=12"] +
                  diff("rkl2", ("isotropic", 0.05), 0.05, 0.05)),
    # RKL2 with anisotropic (field aligned, saturated) conduction and a dt ratio cap
    "ot2d_rkl2_aniso": (OT, mesh((64, 64, 1), (32, 32, 1), 3) + NOFOFC + ["parthenon/time/nlim=12"] +
                        diff("rkl2", ("anisotropic", 0.2), 0.02, 0.03, ratio=6.0, cfl=0.25)),
    # 3-D, 2 x 2 x 2 meshblocks
    "ot3d_unsplit": (OT3, mesh((32, 32, 32), (16, 16, 16), 3) + ["parthenon/time/nlim=8"] +
                     diff("unsplit", ("isotropic", 0.004), 0.003, 0.005)),
    "ot3d_rkl2_aniso": (OT3, mesh((32, 32, 32), (16, 16, 16), 3) + ["parthenon/time/nlim=6"] +
                        diff("rkl2", ("anisotropic", 0.1), 0.04, 0.05)),
    # hydro: 2-D blast, RK3 + RKL2; 1-D Sod with outflow boundaries, unsplit
    "blast2d_rkl2_hydro": (BLAST, mesh((64, 64, 1), (32, 32, 1), 2) + [
        "parthenon/mesh/refinement=none", "parthenon/time/nlim=12", "parthenon/time/integrator=rk3",
        "hydro/reconstruction=plm", "hydro/riemann=hllc", "problem/blast/radius_outer=0.1",
        "problem/blast/radius_inner=0.1"] + diff("rkl2", ("isotropic", 0.02), 0.03)),
    "sod_unsplit": (os.path.join(INPUTS, "sod.in"), ["hydro/riemann=hllc", "parthenon/output0/file_type=hst",
                                                     "parthenon/time/nlim=40"] +
                    diff("unsplit", ("isotropic", 0.001), 0.002)),
}


def amr(regions, fluid, recon, rs, ng, extra):
    """blast_3d_amr.in on a frozen refined mesh (refinement = static), as in make_reference_goldens.py"""
    ov = ["parthenon/mesh/refinement=static", "hydro/fluid=" + fluid, "hydro/reconstruction=" + recon,
          "hydro/riemann=" + rs, "parthenon/mesh/nghost=%d" % ng, "parthenon/time/nlim=10",
          "parthenon/output0/file_type=hst", "problem/blast/radius_outer=0.12",
          "problem/blast/radius_inner=0.12", "problem/blast/pressure_ratio=100",
          "problem/blast/pressure_ambient=0.1", "problem/blast/x1_0=0.22", "problem/blast/x2_0=0.1",
          "problem/blast/x3_0=-0.05"]
    for r, (lev, lo, hi) in enumerate(regions):
        for d in range(3):
            ov += ["parthenon/static_refinement%d/x%dmin=%r" % (r, d + 1, lo),
                   "parthenon/static_refinement%d/x%dmax=%r" % (r, d + 1, hi)]
        ov += ["parthenon/static_refinement%d/level=%d" % (r, lev)]
    return (BLAST, ov + extra)


# diffusion on refined meshes: the diffusive fluxes go through the flux correction and every RKL2
# sub-step through the multilevel ghost exchange (hydro_driver.cpp:248-273, :314-338)
CASES["amr_2level_rkl2_hydro"] = amr([(1, -0.2, 0.2)], "euler", "plm", "hlle", 2,
                                     diff("rkl2", ("isotropic", 0.02), 0.03))
CASES["amr_3level_unsplit_mhd"] = amr([(1, -0.2, 0.2), (2, -0.1, 0.1)], "glmmhd", "plm", "hlld", 2,
                                      diff("unsplit", ("isotropic", 0.0002), 0.0003, 0.0002))
CASES["amr_3level_rkl2_mhd_aniso"] = amr([(1, -0.2, 0.2), (2, -0.1, 0.1)], "glmmhd", "wenoz", "hlld", 4,
                                         diff("rkl2", ("anisotropic", 0.01), 0.004, 0.005))


def main():
    out = {"_provenance": "printed by the UNMODIFIED reference (Kokkos Serial, g++ -O3, no FMA) run by "
                          "tests/golden/make_diffusion_goldens.py; trace = [time, dt] per cycle at 17 "
                          "digits, s_rkl = RKL2 stages of every super step (two per cycle)"}
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        jobs = {k: ex.submit(run, *v) for k, v in CASES.items()}
        for k, f in jobs.items():
            out[k] = f.result()
            print(k, out[k]["ncycle"], out[k]["trace"][-1], out[k]["s_rkl"][:6], flush=True)
    json.dump(out, open(os.path.join(HERE, "reference_traces_diffusion.json"), "w"), indent=0)


if __name__ == "__main__":
    sys.exit(main())
