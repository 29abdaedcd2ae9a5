#!/usr/bin/env python
"""BASELINE config 5 as the reference runs it (inputs/blast_3d_amr.in, refinement = adaptive,
GLM-MHD WENOZ + HLLD, nghost = 4, 3 levels; the blast of tests/golden/make_amr_goldens.py) on one
GPU, default FMA mode: zone-cycles per wall second including the per-cycle tagging and remeshing,
beside the figure the unmodified serial reference prints for the same deck
(tests/golden/reference_traces_amr.json was produced by that run: 12 cycles, 435 -> 526 blocks)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import athenapk_b200 as apk  # noqa: E402
from trace_util import ADAPTIVE_DECKS  # noqa: E402

res = {}
for name, ncyc in (("adaptive_3d_wenoz_hlld", 40), ("adaptive_3d_plm_hlle", 60)):
    kw, pgen, pars, tlim = ADAPTIVE_DECKS[name]
    s = apk.HydroSim(fp_mode="fast", **kw)
    s.pgen(pgen, pars)
    t0 = time.time()
    s.initialize(tlim)
    s.synchronize()
    t_init = time.time() - t0
    nb0 = s.nblocks
    s.execute(tlim, 3)
    s.synchronize()
    n_pre, ms_pre = s.amr_timing()
    zones = 0
    t0 = time.time()
    for c in range(3, 3 + ncyc):
        zones += s.nblocks * kw["mb"][0] * kw["mb"][1] * kw["mb"][2]
        s.execute(tlim, c + 1)
    s.synchronize()
    dt = time.time() - t0
    created, destroyed, per_level = s.amr_counts()
    n_amr, ms_amr = s.amr_timing()
    res[name] = {"zone_cycles_per_s": zones / dt, "ms_per_cycle": 1e3 * dt / ncyc, "cycles": ncyc,
                 "blocks_start": nb0, "blocks_end": s.nblocks, "created": created,
                 "destroyed": destroyed, "per_level": per_level[:4], "init_s": t_init,
                 "remeshes": n_amr - n_pre,
                 "amr_ms_per_cycle": {k: (ms_amr[k] - ms_pre[k]) / ncyc for k in ms_amr}}
    # device time per task (CUDA events around every task; perturbs the wall clock, hence after it)
    s.enable_task_timing(True)
    c0 = s.ncycle
    s.execute(tlim, c0 + 10)
    s.synchronize()
    tasks = {}
    for t in ("init_u1", "calc_flux", "flux_correction", "update", "source_unsplit", "send", "set",
              "prolongate_bounds", "fill_derived", "new_dt", "tag_refine"):
        ms, cnt = s.task_timing(t)
        if cnt:
            tasks[t] = ms / 10.0
    res[name]["task_ms_per_cycle"] = tasks
    s.close()
print(json.dumps(res, indent=1))
