#!/usr/bin/env python
"""Tiny fused-path runs for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athenapk_b200 as apk  # noqa: E402

CASES = [
    dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(32, 16, 16),
         mb=(32, 8, 8), ng=3),
    dict(fluid="euler", recon="plm", riemann="hllc", integrator="rk3", nx=(24, 12, 1),
         mb=(12, 12, 1), ng=2),
    dict(fluid="glmmhd", recon="wenoz", riemann="hlle", integrator="rk2", nx=(20, 10, 10),
         mb=(20, 10, 10), ng=3),
]
for kw in CASES:
    s = apk.HydroSim(fp_mode="fast", xmin=(0.0, 0.0, 0.0), xmax=(1.0, 1.0, 1.0), gamma=5.0 / 3.0,
                     cfl=0.3, **kw)
    s.pgen("random", (3, 0.2))
    s.initialize()
    s.execute(None, 1)
    print("ok", kw["recon"], kw["riemann"], s.dt, s.history()[:2])
    s.close()
