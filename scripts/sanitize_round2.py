#!/usr/bin/env python
"""Tiny runs of the round-2 additions for compute-sanitizer (memcheck / racecheck): the lean PPM
recon phase with work lists (ragged tiles in x1 and x2, several x3 chunks), the row-pass A/B kernel
(APK_B200_FUSED_V=2 in the environment), the diffusion kernels (unsplit / RKL2, all processes) and a
refined (static AMR) mesh.
   compute-sanitizer --tool racecheck python scripts/sanitize_round2.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athenapk_b200 as apk  # noqa: E402

BOX = dict(xmin=(0.0, 0.0, 0.0), xmax=(1.0, 1.0, 1.0), gamma=5.0 / 3.0, cfl=0.3)
DIFF = dict(conduction="anisotropic", thermal_diff_coeff=0.05, viscosity="isotropic",
            mom_diff_coeff=0.03, resistivity="ohmic", ohm_diff_coeff=0.04)
CASES = [
    (dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(36, 10, 12),
          mb=(36, 10, 12), ng=3), "random", (3, 0.3)),
    (dict(fluid="euler", recon="ppm", riemann="hllc", integrator="rk3", nx=(64, 16, 8),
          mb=(32, 8, 8), ng=3), "random", (5, 0.3)),
    (dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(16, 16, 16),
          mb=(8, 8, 8), ng=3, diffint="rkl2", **DIFF), "ot3d", ()),
    (dict(fluid="glmmhd", recon="plm", riemann="hlle", integrator="rk2", nx=(16, 16, 1),
          mb=(8, 8, 1), ng=2, diffint="unsplit", **DIFF), "orszag_tang", ()),
    (dict(fluid="euler", recon="plm", riemann="hlle", integrator="vl2", nx=(16, 16, 16),
          mb=(4, 4, 4), ng=2, refine=[(1, (0.3,) * 3, (0.7,) * 3)]), "random", (2, 0.2)),
]
for kw, pgen, pars in CASES:
    s = apk.HydroSim(fp_mode="fast", **BOX, **kw)
    s.pgen(pgen, pars)
    s.initialize()
    s.execute(None, 1)
    print("ok", kw["recon"], kw["riemann"], s.fused_active, s.dt, s.history()[:2])
    s.close()
