#!/usr/bin/env python
"""Generates tests/golden/reference_traces_amr.json by RUNNING THE UNMODIFIED REFERENCE
(baseline/_ref/serial/bin/athenaPK, built by baseline/build_reference.sh) on inputs/blast_3d_amr.in
with `refinement = adaptive` (BASELINE config 5) and the CLI overrides recorded per case.  Kept per
case: Parthenon's per-cycle `time` / `dt` printout at all 17 digits (driver.cpp:301-306), the number
of meshblocks it prints every cycle (`ncycle_out_mesh = -1`, driver.cpp:311-330) with the count
per level, and the created / destroyed totals of the closing message (driver.cpp:177-181).

  python tests/golden/make_amr_goldens.py          (about 3 minutes)

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
DECK = "/root/reference/inputs/blast_3d_amr.in"
CYCLE = re.compile(r"cycle=(\d+) time=(\S+) dt=(\S+)")
MESH = re.compile(r"Total number of MeshBlocks = (\d+)\n(?:Number of .*\n)*((?:  Physical level = .*\n)+)")
LEVEL = re.compile(r"Physical level = (\d+) .*: (\d+) MeshBlocks")
TOTALS = re.compile(r"Number of MeshBlocks = (\d+); (\d+)  created, (\d+) destroyed")

BLAST = ["problem/blast/radius_outer=0.12", "problem/blast/radius_inner=0.12",
         "problem/blast/pressure_ratio=100", "problem/blast/pressure_ambient=0.1",
         "problem/blast/x1_0=0.22", "problem/blast/x2_0=0.1", "problem/blast/x3_0=-0.05"]
COMMON = ["parthenon/output0/file_type=hst", "parthenon/time/ncycle_out_mesh=-1"]

CASES = {
    # the deck as it is (Euler, PLM + HLLE, nghost = 2, 3 levels), blast off the origin
    "adaptive_3d_plm_hlle": ["parthenon/time/nlim=30"] + BLAST,
    # BASELINE config 5's method: GLM-MHD, WENOZ + HLLD, nghost = 4 (odd nghost is rejected with
    # mesh refinement, mesh_refinement.cpp:60-70)
    "adaptive_3d_wenoz_hlld": ["parthenon/time/nlim=12", "hydro/fluid=glmmhd",
                               "hydro/reconstruction=wenoz", "hydro/riemann=hlld",
                               "parthenon/mesh/nghost=4"] + BLAST,
    # 2-D, 4 levels, quick derefinement (derefine_count = 3) behind the expanding shock
    "adaptive_2d_plm_hllc": ["parthenon/time/nlim=120", "parthenon/mesh/nx1=64",
                             "parthenon/mesh/nx2=64", "parthenon/mesh/nx3=1",
                             "parthenon/meshblock/nx3=1", "parthenon/mesh/numlevel=4",
                             "parthenon/mesh/derefine_count=3", "hydro/riemann=hllc",
                             "refinement/threshold_pressure_gradient=0.2"] + BLAST,
}


# refined meshes on NON-periodic domains: physical boundary conditions on the coarse buffers before
# the prolongation and on the blocks after it (boundary_communication.cpp:437-442), the blast placed
# on the domain boundary inside the refined region
def _bc(ix1, ox1, ix2, ox2, ix3="periodic", ox3="periodic"):
    return ["parthenon/mesh/ix1_bc=" + ix1, "parthenon/mesh/ox1_bc=" + ox1,
            "parthenon/mesh/ix2_bc=" + ix2, "parthenon/mesh/ox2_bc=" + ox2,
            "parthenon/mesh/ix3_bc=" + ix3, "parthenon/mesh/ox3_bc=" + ox3]


def _blast(x, y, z):
    return BLAST[:4] + ["problem/blast/x1_0=%r" % x, "problem/blast/x2_0=%r" % y,
                        "problem/blast/x3_0=%r" % z]


_2D = ["parthenon/mesh/nx1=64", "parthenon/mesh/nx2=64", "parthenon/mesh/nx3=1",
       "parthenon/meshblock/nx3=1"]
CASES.update({
    "static_2d_outflow": ["parthenon/time/nlim=40", "parthenon/mesh/refinement=static",
                          "parthenon/static_refinement0/x1min=-0.5", "parthenon/static_refinement0/x1max=-0.2",
                          "parthenon/static_refinement0/x2min=-0.5", "parthenon/static_refinement0/x2max=-0.1",
                          "parthenon/static_refinement0/level=2"] + _2D +
                         _bc("outflow", "outflow", "outflow", "outflow") + _blast(-0.42, -0.4, 0.0),
    "adaptive_2d_reflect_outflow": ["parthenon/time/nlim=80", "parthenon/mesh/derefine_count=5",
                                    "hydro/riemann=hllc"] + _2D +
                                   _bc("reflecting", "outflow", "outflow", "reflecting") +
                                   _blast(-0.42, 0.41, 0.0),
    "static_3d_outflow_wenoz_hlld": ["parthenon/time/nlim=10", "parthenon/mesh/refinement=static",
                                     "hydro/fluid=glmmhd", "hydro/reconstruction=wenoz",
                                     "hydro/riemann=hlld", "parthenon/mesh/nghost=4",
                                     "parthenon/static_refinement0/x1min=0.1", "parthenon/static_refinement0/x1max=0.5",
                                     "parthenon/static_refinement0/x2min=-0.5", "parthenon/static_refinement0/x2max=-0.2",
                                     "parthenon/static_refinement0/x3min=-0.2", "parthenon/static_refinement0/x3max=0.2",
                                     "parthenon/static_refinement0/level=2"] +
                                    _bc("outflow", "outflow", "outflow", "outflow", "outflow", "outflow") +
                                    _blast(0.4, -0.41, 0.05),
})


# the other <refinement> criteria (src/refinement/gradient.cpp:64-95, other.cpp:18-44)
CASES.update({
    "adaptive_2d_xyvelocity": ["parthenon/time/nlim=80", "parthenon/mesh/derefine_count=5",
                               "refinement/type=xyvelocity_gradient",
                               "refinement/threshold_xyvelocity_gradient=0.05"] + _2D + BLAST,
    "adaptive_2d_maxdensity": ["parthenon/time/nlim=80", "parthenon/mesh/derefine_count=5",
                               "refinement/type=maxdensity", "refinement/maxdensity_refine_above=1.3",
                               "refinement/maxdensity_deref_below=1.1", "hydro/riemann=hllc"] + _2D + BLAST,
})


# GLM-MHD with B != 0 on an adaptive mesh: the Orszag-Tang vortex (inputs/orszag_tang.in, FOFC on)
# -- c_h = cfl * mindx / dt_hyp follows the finest level after every remesh (hydro.cpp:102-143)
CASES["adaptive_2d_orszag_tang_mhd"] = [
    "parthenon/mesh/refinement=adaptive", "parthenon/mesh/numlevel=3", "parthenon/mesh/derefine_count=5",
    "parthenon/mesh/nx1=64", "parthenon/mesh/nx2=64", "parthenon/meshblock/nx1=8",
    "parthenon/meshblock/nx2=8", "parthenon/mesh/nghost=4", "parthenon/time/nlim=60",
    "parthenon/output0/dt=-1", "parthenon/output1/dt=-1", "refinement/type=pressure_gradient",
    "refinement/threshold_pressure_gradient=0.05"]
DECK_OF = {"adaptive_2d_orszag_tang_mhd": "/root/reference/inputs/orszag_tang.in"}
# a three-stage integrator (flux correction and multilevel exchange in every stage) and RKL2 super
# time stepping (every sub-step through the exchange) on adaptive meshes
CASES.update({
    "adaptive_2d_rk3_weno3_hllc": ["parthenon/time/nlim=40", "parthenon/time/integrator=rk3",
                                   "hydro/reconstruction=weno3", "hydro/riemann=hllc",
                                   "parthenon/mesh/derefine_count=4"] + _2D + BLAST,
    "adaptive_2d_rkl2_diffusion": ["parthenon/time/nlim=30", "parthenon/mesh/derefine_count=4",
                                   "diffusion/integrator=rkl2", "diffusion/conduction=isotropic",
                                   "diffusion/conduction_coeff=fixed",
                                   "diffusion/thermal_diff_coeff_code=0.02",
                                   "diffusion/viscosity=isotropic", "diffusion/viscosity_coeff=fixed",
                                   "diffusion/mom_diff_coeff_code=0.03"] + _2D + BLAST,
})


def run(overrides, deck=DECK):
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [BIN, "-i", deck] + overrides + COMMON
        p = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError(" ".join(cmd) + "\n" + p.stdout[-2000:])
    trace = [[m.group(2), m.group(3)] for m in CYCLE.finditer(p.stdout)]
    # one mesh printout per cycle line ("Current Mesh structure" / "New Mesh structure")
    blocks = []
    for m in MESH.finditer(p.stdout):
        per = {int(a): int(b) for a, b in LEVEL.findall(m.group(2))}
        blocks.append([int(m.group(1))] + [per.get(l, 0) for l in range(max(per) + 1)])
    tot = TOTALS.search(p.stdout)  # (printed by adaptive runs only)
    return {"overrides": overrides, "ncycle": len(trace) - 1, "trace": trace, "blocks": blocks,
            "final": [int(tot.group(1)), int(tot.group(2)), int(tot.group(3))] if tot else None}


def main():
    only = sys.argv[1:]  # (a subset of the cases: the others are kept from the existing file)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        jobs = {k: ex.submit(run, ov, DECK_OF.get(k, DECK)) for k, ov in CASES.items()
                if not only or k in only}
    out = {"_provenance": "printed by the UNMODIFIED reference (Kokkos Serial, g++ -O3, no FMA) run by "
                          "tests/golden/make_amr_goldens.py; trace = [time, dt] per cycle at 17 digits, "
                          "blocks = [total, level 0, level 1, ...] per cycle, final = [blocks, created, destroyed]"}
    for k, f in jobs.items():
        out[k] = f.result()
        print(k, out[k]["ncycle"], out[k]["trace"][-1], out[k]["final"], len(out[k]["blocks"]), flush=True)
    path = os.path.join(HERE, "reference_traces_amr.json")
    if only:
        out = dict(json.load(open(path)), **out)
    json.dump(out, open(path, "w"), indent=0)


if __name__ == "__main__":
    sys.exit(main())
