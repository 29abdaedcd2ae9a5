"""The reference's regression suite case tst/regression/test_suites/aniso_therm_cond_ring_conv:
anisotropic conduction of a hot patch around a ring of circular field lines (problem generator
`diffusion`, iprob = 20; inputs/diffusion.in: GLM-MHD, no hyperbolic fluxes, gamma = 2, periodic
[-1, 1]^2, RKL2 with rkl2_max_dt_ratio = 100, t = 200), N = 32 / 64 / 128 / 256 in 2 x 2
meshblocks.  Acceptance criteria (aniso_therm_cond_ring_conv.py:80-150): no temperature below the
background (the limited transverse gradients keep the scheme monotone), L2 error against the
fully diffused ring at N = 256 <= 0.0264, L1 error converging at a rate <= -0.53 and L2 at <= -0.35
(the L1 errors of Balsara, Tilley & Howk 2007 are 0.0407, 0.0326, 0.0237, 0.0150)."""
import numpy as np

TLIM = 200.0


def deck(res):
    return dict(fluid="glmmhd", nx=(res, res, 1), mb=(res // 2, res // 2, 1), ng=2,
                xmin=(-1.0, -1.0, -1.0), xmax=(1.0, 1.0, 1.0), gamma=2.0, cfl=0.3, recon="dc",
                riemann="none", integrator="rk1", diffint="rkl2", conduction="anisotropic",
                thermal_diff_coeff=0.01, rkl2_max_dt_ratio=100.0), \
        "diffusion", [20, 1.0, 0.0, 0.5, 1e-6, 0.01]


def errors(blocks, res):
    """blocks: [(xmin_incl_ghosts, dx, prim)] per meshblock -> (min T, L1, L2)"""
    h, T, R = res // 2, [], []
    for xmin, dx, prim in blocks:
        i = np.arange(2, 2 + h)
        X, Y = np.meshgrid(xmin[0] + (i + 0.5) * dx[0], xmin[1] + (i + 0.5) * dx[1])
        R.append(np.sqrt(X ** 2 + Y ** 2).ravel())
        T.append(prim[4][0, 2:2 + h, 2:2 + h].ravel())  # p = e = T (gamma = 2, rho = 1)
    T, r = np.concatenate(T), np.concatenate(R)
    ref = np.where(np.abs(r - 0.6) < 0.1, 10.1667, 10.0)
    return float(T.min()), float(np.mean(np.abs(T - ref))), float(np.sqrt(np.mean(np.abs(T - ref) ** 2.0)))


def rate(res_list, errs):
    return float(np.polyfit(np.log(res_list), np.log(errs), 1)[0])
