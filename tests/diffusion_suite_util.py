"""The reference's regression suite case tst/regression/test_suites/diffusion (diffusion.py): viscous
diffusion of a 1-D Gaussian velocity profile (problem generator `diffusion`, iprob = 30), unsplit and
RKL2 integrators, outflow walls, no hyperbolic fluxes (riemann = none), N = 256 / 512 / 1024.
  * the history files the UNMODIFIED reference wrote at N = 256 (tests/golden/
    reference_diffusion_visc_256_{unsplit,rkl2}.hst, output every 0.1: the kinetic energy of the
    spreading profile at 6 digits) must be reproduced row by row;
  * the case's acceptance criterion (diffusion.py:108-180): L1 error of v2 against the analytic
    Gaussian at t = 2 converges with N at a rate <= -1.95.
The Ohmic half of the suite (iprob = 40) is not pinned: the reference's serial build segfaults on
it in this container.  Command line: diffusion.py:43-90 (its hydro/cfl, hydro/integrator entries
are in the wrong input block and therefore inactive: rk1, cfl = 0.3 of inputs/diffusion.in)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIFF_COEFF, TLIM, T0, AMP = 0.25, 2.0, 0.5, 1e-6


def deck(res, integrator):
    return dict(fluid="euler", nx=(res, 1, 1), mb=(64, 1, 1), ng=2, xmin=(-6.0, -1.0, -1.0),
                xmax=(6.0, 1.0, 1.0), gamma=1.4, cfl=0.3, recon="dc", riemann="none",
                integrator="rk1", bc=["outflow"] * 2 + ["periodic"] * 4, diffint=integrator,
                viscosity="isotropic", mom_diff_coeff=DIFF_COEFF, rkl2_max_dt_ratio=200.0), \
        "diffusion", [30, 0.0, 0.0, T0, AMP, DIFF_COEFF]


def reference_rows(integrator):
    f = os.path.join(GOLDEN, "reference_diffusion_visc_256_%s.hst" % integrator)
    return [l.split() for l in open(f).read().splitlines()[2:]]


def history_rows(engine, advance, history, nblocks, dt_out=0.1):
    def row():
        return ["%12.5e" % engine.time, "%12.5e" % engine.dt, "%d" % engine.ncycle, "%d" % nblocks] + \
               ["%12.5e" % v for v in list(history(engine))[:6]]
    rows, nxt = [row()], dt_out
    while engine.time < TLIM:
        advance(engine, TLIM, engine.ncycle + 1)
        if engine.time < TLIM and engine.time >= nxt:
            rows.append(row())
            nxt += dt_out
    rows.append(row())
    return [[x.strip() for x in r] for r in rows]


def check_rows(rows, ref):
    assert len(rows) == len(ref)
    for g, r in zip(rows, ref):
        assert g[:4] == r[:4], (g, r)           # time dt cycle nbtotal as printed
        for c in (4, 6, 8, 9):                   # mass 2-mom KE tot-E: last printed digit
            a, b = float(g[c]), float(r[c])
            assert abs(a - b) <= 1.01e-5 * 10.0 ** np.floor(np.log10(abs(b))), (c, g, r)


def l1_error(x, v2):
    ref = AMP / np.sqrt(4.0 * np.pi * DIFF_COEFF * (T0 + TLIM)) * \
        np.exp(-(x ** 2.0) / (4.0 * DIFF_COEFF * (T0 + TLIM)))
    return float(np.average(np.abs(v2 - ref)))


def convergence_rate(res_list, l1_list):
    """diffusion.py:150-163 (least-squares slope of log L1 against log N)"""
    return float(np.polyfit(np.log(res_list), np.log(l1_list), 1)[0])
