#!/usr/bin/env python
"""Tiny un-fused / boundary-condition / FOFC runs for compute-sanitizer memcheck."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athenapk_b200 as apk  # noqa: E402

OUT = ["outflow"] * 4 + ["periodic"] * 2
REF = ["reflecting"] * 6
CASES = [
    (dict(fluid="euler", recon="plm", riemann="hllc", integrator="rk3", nx=(24, 12, 1), mb=(12, 12, 1),
          ng=2, bc=OUT, fused=False), "blast", (0.2, 0.1, 20.0, 1.0)),
    (dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(16, 16, 16),
          mb=(8, 8, 8), ng=3, fofc=True), "random", (3, 0.4)),
    (dict(fluid="glmmhd", recon="limo3", riemann="hlle", integrator="rk2", nx=(16, 8, 8),
          mb=(16, 8, 8), ng=3, bc=REF, fused=True), "random", (3, 0.1)),
    (dict(fluid="euler", recon="weno3", riemann="hlle", integrator="rk1", nx=(40, 1, 1), mb=(8, 1, 1),
          ng=2, bc=["outflow", "outflow"] + ["periodic"] * 4, fused=True), "sod", ()),
]
for kw, pgen, pars in CASES:
    s = apk.HydroSim(fp_mode="fast", xmin=(0.0, 0.0, 0.0), xmax=(1.0, 1.0, 1.0), gamma=1.4, cfl=0.3,
                     **kw)
    s.pgen(pgen, pars)
    s.initialize()
    s.execute(None, 2)
    print("ok", kw["recon"], kw["riemann"], s.fused_active, s.dt, s.history()[0])
    s.close()
