"""The reference's regression suite case tst/regression/test_suites/aniso_therm_cond_ring_multid:
the ring-conduction problem (problem generator `diffusion`, iprob = 20 / 21 / 22; see
ring_cond_suite_util.py) with the UNSPLIT integrator on a 2-D mesh and on 3-D meshes with the ring
in the x1-x2, x2-x3 and x3-x1 plane (8 cells along the third direction).  Acceptance criteria
(aniso_therm_cond_ring_multid.py:107-158): no temperature below the background, the 2-D and 3-D
errors agree to 4 decimals, the three 3-D orientations agree to 14 decimals (the kernels treat the
directions alike).  The case runs N = 64 to t = 200 (27 000 - 41 000 cycles)."""
import numpy as np


def deck(res, plane):
    """plane: '2d', 'x1x2', 'x2x3', 'x3x1'"""
    h = res // 2
    nx, mb, iprob = {"2d": ((res, res, 1), (h, h, 1), 20), "x1x2": ((res, res, 8), (h, h, 8), 20),
                     "x2x3": ((8, res, res), (8, h, h), 21), "x3x1": ((res, 8, res), (h, 8, h), 22)}[plane]
    return dict(fluid="glmmhd", nx=nx, mb=mb, ng=2, xmin=(-1.0, -1.0, -1.0), xmax=(1.0, 1.0, 1.0),
                gamma=2.0, cfl=0.3, recon="dc", riemann="none", integrator="rk1", diffint="unsplit",
                conduction="anisotropic", thermal_diff_coeff=0.01), \
        "diffusion", [iprob, 1.0, 0.0, 0.5, 1e-6, 0.01], iprob


def errors(blocks, nx, mb, iprob):
    """blocks: [(xmin_incl_ghosts, dx, prim)] -> (min T, L1, L2) against the fully diffused ring"""
    T, R = [], []
    g = [2 if nx[d] > 1 else 0 for d in range(3)]
    for xmin, dx, prim in blocks:
        xs = [xmin[d] + (np.arange(g[d], g[d] + mb[d]) + 0.5) * dx[d] for d in range(3)]
        Z, Y, X = np.meshgrid(xs[2], xs[1], xs[0], indexing="ij")
        a, b = {20: (X, Y), 21: (Y, Z), 22: (Z, X)}[iprob]
        R.append(np.sqrt(a ** 2 + b ** 2).ravel())
        T.append(prim[4][g[2]:g[2] + mb[2], g[1]:g[1] + mb[1], g[0]:g[0] + mb[0]].ravel())
    T, r = np.concatenate(T), np.concatenate(R)
    ref = np.where(np.abs(r - 0.6) < 0.1, 10.1667, 10.0)
    return float(T.min()), float(np.mean(np.abs(T - ref))), float(np.sqrt(np.mean(np.abs(T - ref) ** 2.0)))


def check(errs):
    """errs: {plane: (min T, L1, L2)} -- aniso_therm_cond_ring_multid.py:136-157"""
    for plane, (tmin, l1, l2) in errs.items():
        assert tmin >= 10.0, (plane, tmin)
    for q in (1, 2):
        assert abs(errs["2d"][q] - errs["x1x2"][q]) < 1.5e-4, errs
        assert abs(errs["x1x2"][q] - errs["x2x3"][q]) < 1.5e-14, errs
        assert abs(errs["x2x3"][q] - errs["x3x1"][q]) < 1.5e-14, errs
