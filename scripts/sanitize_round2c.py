#!/usr/bin/env python
"""Tiny runs of the additions of the second half of round 2 for compute-sanitizer (memcheck):
adaptive remeshing (tagging with all three criteria, data-move kernels, arena-allocated arrays,
rebuilt exchange plan), refined meshes with physical boundaries (BCs on the coarse buffers), the
fused stage kernel carrying passive scalars (1-D / 2-D / 3-D, ragged tiles) and the history writer.
   compute-sanitizer --tool memcheck python scripts/sanitize_round2c.py"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athenapk_b200 as apk  # noqa: E402

G = float("1.666666666666667")
BLAST = (0.12, 0.1, 100.0, 1.0, 0.22, 0.1, -0.05)
BOX = dict(xmin=(-0.5,) * 3, xmax=(0.5,) * 3, gamma=G, cfl=0.3)
AMR = [
    # (config, pgen pars, cycles)
    (dict(fluid="euler", recon="plm", riemann="hlle", integrator="vl2", nx=(32, 32, 1), mb=(8, 8, 1),
          ng=2, adaptive=dict(numlevel=3, threshold=0.1, derefine_count=2)), BLAST, 12),
    (dict(fluid="glmmhd", recon="wenoz", riemann="hlld", integrator="vl2", nx=(16, 16, 16),
          mb=(8, 8, 8), ng=4, adaptive=dict(numlevel=2, threshold=0.1)), BLAST, 3),
    (dict(fluid="euler", recon="plm", riemann="hllc", integrator="vl2", nx=(32, 32, 1), mb=(8, 8, 1),
          ng=2, bc=["reflecting", "outflow", "outflow", "reflecting", "periodic", "periodic"],
          adaptive=dict(numlevel=3, threshold=0.05, derefine_count=3, type="xyvelocity_gradient")),
     (0.12, 0.1, 100.0, 1.0, -0.42, 0.41, 0.0), 12),
    (dict(fluid="euler", recon="plm", riemann="hllc", integrator="vl2", nx=(32, 32, 1), mb=(8, 8, 1),
          ng=2, bc=["outflow"] * 4 + ["periodic"] * 2,
          adaptive=dict(numlevel=3, threshold=1.3, threshold2=1.1, derefine_count=3, type="maxdensity")),
     (0.12, 0.1, 100.0, 1.0, -0.42, 0.41, 0.0), 12),
    (dict(fluid="glmmhd", recon="plm", riemann="hlld", integrator="rk2", nx=(16, 16, 16), mb=(4, 4, 4),
          ng=2, bc=["outflow"] * 6, refine=[(1, (0.0, -0.5, -0.2), (0.5, 0.0, 0.2))]),
     (0.12, 0.1, 100.0, 1.0, 0.4, -0.41, 0.05), 2),
]
for kw, pars, ncyc in AMR:
    s = apk.HydroSim(fp_mode="fast", **BOX, **kw)
    s.pgen("blast", pars)
    s.initialize(0.1)
    nb0 = s.nblocks
    s.execute(0.1, ncyc)
    print("amr ok", kw["recon"], kw["riemann"], nb0, "->", s.nblocks, s.amr_counts()[:2], s.dt)
    s.close()

SCAL = [
    dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(36, 10, 12), mb=(36, 10, 12),
         ng=3, nscalars=1),
    dict(fluid="euler", recon="wenoz", riemann="hllc", integrator="rk3", nx=(64, 14, 8), mb=(32, 7, 8),
         ng=3, nscalars=2),
    dict(fluid="euler", recon="plm", riemann="hlle", integrator="vl2", nx=(40, 12, 1), mb=(40, 12, 1),
         ng=2, nscalars=2),
    dict(fluid="glmmhd", recon="dc", riemann="llf", integrator="rk1", nx=(24, 1, 1), mb=(12, 1, 1), ng=2,
         nscalars=1),
]
for kw in SCAL:
    s = apk.HydroSim(fp_mode="fast", xmin=(0.0,) * 3, xmax=(1.0,) * 3, gamma=5.0 / 3.0, cfl=0.3, **kw)
    s.pgen("random", (3, 0.2))
    s.initialize()
    with tempfile.TemporaryDirectory() as tmp:
        s.enable_history(os.path.join(tmp, "x.hst"), 0.0)
        s.execute(None, 2)
        s.finalize_outputs()
        rows = len(open(os.path.join(tmp, "x.hst")).read().splitlines())
    assert s.fused_active
    print("scalars ok", kw["recon"], kw["riemann"], kw["nx"], s.dt, "hst rows", rows)
    s.close()
