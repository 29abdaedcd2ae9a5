"""The reference's regression suite case tst/regression/test_suites/aniso_therm_cond_gauss_conv:
thermal conduction of a 1-D Gaussian temperature profile on a 2-D mesh (problem generator
`diffusion`, iprob = 10; inputs/diffusion.in: GLM-MHD, no hyperbolic fluxes, gamma = 2, periodic),
isotropic (`none`) or anisotropic along a field that is aligned with x1, at 45 degrees, or
perpendicular; unsplit (rk2) and RKL2 (rk1) integrators; N = 128 / 256 / 512 by 32 cells.
Acceptance criterion of the case (aniso_therm_cond_gauss_conv.py:127-199): the L1 error of the
temperature (= pressure: rho = 1, gamma = 2) against the analytic Gaussian with the effective
coefficient 0.25 Bx^2 converges at a rate <= -1.98; with a perpendicular field nothing diffuses."""
import numpy as np

TLIM, T0, AMP, COEFF = 2.0, 0.5, 1e-6, 0.25
FIELDS = {"none": (0.0, 0.0), "aligned": (1.0, 0.0), "angle": (1 / np.sqrt(2), 1 / np.sqrt(2)),
          "perp": (0.0, 1.0)}


def deck(res, field, integrator):
    bx, by = FIELDS[field]
    # the reference formats the field components with %f (6 decimals) on its command line
    bx, by = float("%f" % bx), float("%f" % by)
    return dict(fluid="glmmhd", nx=(res, 32, 1), mb=(64, 32, 1), ng=2, xmin=(-6.0, -1.0, -1.0),
                xmax=(6.0, 1.0, 1.0), gamma=2.0, cfl=0.3, recon="dc", riemann="none",
                integrator="rk2" if integrator == "unsplit" else "rk1", diffint=integrator,
                conduction="isotropic" if field == "none" else "anisotropic",
                thermal_diff_coeff=COEFF, rkl2_max_dt_ratio=100.0), \
        "diffusion", [10, bx, by, T0, AMP, COEFF], bx


def l1_error(x, temp, bx, field):
    eff = COEFF if bx == 0.0 else COEFF * bx * bx
    t = 0.0 if field == "perp" else TLIM
    ref = 1.0 + AMP / (np.sqrt(4 * np.pi * eff * (T0 + t)) / np.exp(-(x ** 2) / (4.0 * eff * (T0 + t))))
    return float(np.average(np.abs(temp - ref)))


def rate(res_list, l1_list):
    return float(np.polyfit(np.log(res_list), np.log(l1_list), 1)[0])
