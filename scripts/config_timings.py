#!/usr/bin/env python
"""Throughput of the small BASELINE configs (1: linear_wave3d 64x32x32 single block RK2+PLM+HLLE,
2: sod 256x1x1 PLM+HLLC, 3: Orszag-Tang 256^2 VL2+PPM+HLLD+FOFC) on one GPU, default FMA mode.
These are parity cases, not bench lines; the numbers show what launch-latency-bound decks cost."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athenapk_b200 as apk  # noqa: E402

OUT = ["outflow", "outflow"] + ["periodic"] * 4
CASES = {
    "config1_linear_wave3d_64x32x32_rk2_plm_hlle": (
        dict(fluid="euler", recon="plm", riemann="hlle", integrator="rk2", nx=(64, 32, 32),
             mb=(64, 32, 32), ng=2, xmin=(0.0, 0.0, 0.0), xmax=(3.0, 1.5, 1.5), gamma=5.0 / 3.0,
             cfl=0.3), "linear_wave", (0, 1e-6, 0.0), 60),
    "config2_sod_256_plm_hllc": (
        dict(fluid="euler", recon="plm", riemann="hllc", integrator="vl2", nx=(256, 1, 1),
             mb=(64, 1, 1), ng=2, bc=OUT, xmin=(0.0, -0.5, -0.5), xmax=(1.0, 0.5, 0.5), gamma=1.4,
             cfl=0.5), "sod", (), 200),
    "config3_orszag_tang_256sq_vl2_ppm_hlld_fofc": (
        dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(256, 256, 1),
             mb=(64, 256, 1), ng=3, xmin=(-0.5, -0.5, -0.5), xmax=(0.5, 0.5, 0.5), gamma=5.0 / 3.0,
             cfl=0.4, fofc=True), "orszag_tang", (), 50),
    # config 5 on a frozen 3-level mesh (176 meshblocks of 8^3, static refinement; the un-fused tasks
    # with flux correction and the six-launch multilevel exchange)
    "config5_blast_3d_static_3level_wenoz_hlld": (
        dict(fluid="glmmhd", recon="wenoz", riemann="hlld", integrator="vl2", nx=(32, 32, 32),
             mb=(8, 8, 8), ng=4, xmin=(-0.5, -0.5, -0.5), xmax=(0.5, 0.5, 0.5), gamma=5.0 / 3.0,
             cfl=0.3, refine=[(1, (-0.2,) * 3, (0.2,) * 3), (2, (-0.1,) * 3, (0.1,) * 3)]),
        "blast", (0.12, 0.1, 100.0, 1.0, 0.22, 0.1, -0.05), 50),
}
res = {}
for name, (kw, pgen, pars, ncyc) in CASES.items():
    s = apk.HydroSim(fp_mode="fast", **kw)
    s.pgen(pgen, pars)
    s.initialize()
    s.execute(None, 5)
    s.synchronize()
    t0 = time.time()
    s.execute(None, 5 + ncyc)
    s.synchronize()
    dt = time.time() - t0
    cells = s.nblocks_global * kw["mb"][0] * kw["mb"][1] * kw["mb"][2]
    res[name] = {"cycles_per_s": ncyc / dt, "cell_updates_per_s": cells * ncyc / dt,
                 "fused": s.fused_active, "us_per_cycle": 1e6 * dt / ncyc}
    s.close()
print(json.dumps(res, indent=1))
