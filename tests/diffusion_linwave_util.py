"""The reference's own regression case tst/regression/test_suites/diffusion_linwave3d (a decaying
slow magnetosonic wave with viscosity, conduction and resistivity, unsplit diffusion integrator,
VL2 + PLM + HLLD, three wave periods) for the oracle and the CUDA product:
  * the history file the UNMODIFIED reference wrote for it (tests/golden/
    reference_diffusion_linwave3d_{16,32}.hst: 201 rows, 427 / 1093 cycles) must be reproduced row
    by row -- time, dt, cycle, mass, KE, tot-E, ME and MaxAbsV2 at every printed digit (the momentum
    sums and relDivB are round-off noise that depends on the summation order);
  * the acceptance criterion of the test itself (diffusion_linwave3d.py:95-181): the decay rate of
    max |v2| fitted by weighted least squares must be within 22 % (N = 16) / 5 % (N = 32) of the
    analytic slow-mode decay rate of Ryu et al.
Generated with the command line of diffusion_linwave3d.py:49-86 (parthenon/time/tlim=3.0 is three
wave periods in the linear-wave problem generators)."""
import os

import numpy as np
from numpy.polynomial import Polynomial

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NU, KAPPA, ETA = 0.01, 0.02, 0.02
REL_TOL = {16: 0.22, 32: 0.05}  # diffusion_linwave3d.py:26


def deck(res):
    return dict(fluid="glmmhd", nx=(2 * res, res, res), ng=2, xmin=(0.0, 0.0, 0.0),
                xmax=(3.0, 1.5, 1.5), gamma=float("1.666666666666667"), cfl=0.3, recon="plm",
                riemann="hlld", integrator="vl2", diffint="unsplit", conduction="isotropic",
                thermal_diff_coeff=KAPPA, viscosity="isotropic", mom_diff_coeff=NU,
                resistivity="ohmic", ohm_diff_coeff=ETA), "linear_wave_mhd", [2, 1e-4, 0.0]


def reference_rows(res):
    f = os.path.join(GOLDEN, "reference_diffusion_linwave3d_%d.hst" % res)
    return [l.split() for l in open(f).read().splitlines()[2:]]


def history_rows(engine, tlim, advance, history, max_abs_v2, dt_out=0.03):
    """what Parthenon's history output writes (outputs.cpp:445-485, history.cpp:172-209): a row at
    cycle 0, whenever time >= next_time while the run keeps going, and the final one"""
    def row():
        vals = list(history(engine)) + [max_abs_v2(engine)]
        return ["%12.5e" % engine.time, "%12.5e" % engine.dt, "%d" % engine.ncycle, "1"] + \
               ["%12.5e" % v for v in vals]
    rows, nxt = [row()], dt_out
    while engine.time < tlim:
        advance(engine, tlim, engine.ncycle + 1)
        if engine.time < tlim and engine.time >= nxt:
            rows.append(row())
            nxt += dt_out
    rows.append(row())
    return [[x.strip() for x in r] for r in rows]


def check_rows(rows, ref, sums_last_digit=False):
    """sums_last_digit: the volume sums may differ by one unit of the last printed digit -- the
    device reducers add in another order than the reference's serial loop, and tot-E of this deck
    (17.04375...) sits on a rounding boundary of the 6-digit format.  time, dt, cycle and MaxAbsV2
    (a maximum: order independent) are always compared as printed."""
    assert len(rows) == len(ref)
    exact = [0, 1, 2, 3, 12] if sums_last_digit else [0, 1, 2, 3, 4, 8, 9, 10, 12]
    sums = [4, 8, 9, 10] if sums_last_digit else []  # mass KE tot-E ME
    for g, r in zip(rows, ref):
        assert [g[c] for c in exact] == [r[c] for c in exact], (g, r)
        for c in sums:
            a, b = float(g[c]), float(r[c])
            assert abs(a - b) <= 1.01e-5 * 10.0 ** np.floor(np.log10(abs(b))), (c, g, r)


def decay_rate_error(rows):
    """diffusion_linwave3d.py:95-150"""
    ksqr = (2.0 * np.pi / 1.0) ** 2
    slow_mode_rate = (4.0 * NU + 3.0 * ETA / 4.0 + KAPPA * 4.0 / 5.0) * (2.0 / 15.0) * ksqr
    tt = np.array([float(r[0]) for r in rows])
    max_vy = np.array([float(r[12]) for r in rows])
    yy = np.log(np.abs(max_vy))
    p = Polynomial.fit(tt, yy, 1, w=np.sqrt(max_vy))
    fit_rate = -p.convert(domain=(-1, 1)).coef[-1]
    return abs(slow_mode_rate / fit_rate - 1.0)
