"""Exact solution of the 1-D Riemann problem for the Euler equations with an ideal-gas EOS
(Toro, Riemann Solvers and Numerical Methods for Fluid Dynamics, ch. 4: two-rarefaction / two-shock
pressure functions + Newton iteration, then sampling).  Independent of both the reference and the
CUDA path: used to pin the shock-tube decks, which the reference's own riemann_hydro suite only
plots (tst/regression/test_suites/riemann_hydro/riemann_hydro.py:116,163: test_success = True)."""
import numpy as np


def _f(p, rho, pk, g):
    a = np.sqrt(g * pk / rho)
    if p > pk:  # shock
        A, B = 2.0 / ((g + 1.0) * rho), (g - 1.0) / (g + 1.0) * pk
        return (p - pk) * np.sqrt(A / (p + B)), (1.0 - 0.5 * (p - pk) / (B + p)) * np.sqrt(A / (p + B))
    pr = p / pk  # rarefaction
    return 2.0 * a / (g - 1.0) * (pr ** ((g - 1.0) / (2.0 * g)) - 1.0), pr ** (-(g + 1.0) / (2.0 * g)) / (rho * a)


def star_state(rl, ul, pl, rr, ur, pr, g):
    p = max(1e-8, 0.5 * (pl + pr))
    for _ in range(100):
        fl, dl = _f(p, rl, pl, g)
        fr, dr = _f(p, rr, pr, g)
        pn = max(1e-10, p - (fl + fr + ur - ul) / (dl + dr))
        if abs(pn - p) / (0.5 * (pn + p)) < 1e-14:
            p = pn
            break
        p = pn
    fl, _ = _f(p, rl, pl, g)
    fr, _ = _f(p, rr, pr, g)
    return p, 0.5 * (ul + ur) + 0.5 * (fr - fl)


def sample(xi, rl, ul, pl, rr, ur, pr, g):
    """(rho, u, p) at similarity coordinate xi = (x - x0) / t"""
    pm, um = star_state(rl, ul, pl, rr, ur, pr, g)
    gm, gp = g - 1.0, g + 1.0
    out = np.empty((3,) + np.shape(xi))
    for idx, s in np.ndenumerate(xi):
        if s <= um:  # left of the contact
            a = np.sqrt(g * pl / rl)
            if pm > pl:  # left shock
                sl = ul - a * np.sqrt(gp / (2 * g) * pm / pl + gm / (2 * g))
                w = (rl, ul, pl) if s <= sl else (rl * (pm / pl + gm / gp) / (gm / gp * pm / pl + 1.0), um, pm)
            else:
                am = a * (pm / pl) ** (gm / (2 * g))
                if s <= ul - a:
                    w = (rl, ul, pl)
                elif s > um - am:
                    w = (rl * (pm / pl) ** (1.0 / g), um, pm)
                else:
                    c = 2.0 / gp + gm / (gp * a) * (ul - s)
                    w = (rl * c ** (2.0 / gm), 2.0 / gp * (a + gm / 2.0 * ul + s), pl * c ** (2.0 * g / gm))
        else:
            a = np.sqrt(g * pr / rr)
            if pm > pr:  # right shock
                sr = ur + a * np.sqrt(gp / (2 * g) * pm / pr + gm / (2 * g))
                w = (rr, ur, pr) if s >= sr else (rr * (pm / pr + gm / gp) / (gm / gp * pm / pr + 1.0), um, pm)
            else:
                am = a * (pm / pr) ** (gm / (2 * g))
                if s >= ur + a:
                    w = (rr, ur, pr)
                elif s <= um + am:
                    w = (rr * (pm / pr) ** (1.0 / g), um, pm)
                else:
                    c = 2.0 / gp - gm / (gp * a) * (ur - s)
                    w = (rr * c ** (2.0 / gm), 2.0 / gp * (-a + gm / 2.0 * ur + s), pr * c ** (2.0 * g / gm))
        out[(slice(None),) + idx] = w
    return out
