"""Shock tubes against the EXACT Riemann solution (BASELINE config 2: "flux correctness vs exact
solution"; Toro tests 1, 6, 7 with the initial conditions of the reference's riemann_hydro suite,
tst/regression/test_suites/riemann_hydro/riemann_hydro.py:40-53, which itself asserts nothing).
CPU: the oracle; GPU: the CUDA path (strict mode must in addition equal the oracle bit for bit)."""
import numpy as np
import pytest

from oracle_lib import Oracle
from riemann_exact import sample, star_state

GAMMA = 1.4
# rho_l, u_l, p_l, rho_r, u_r, p_r, x0, t_end
TUBES = {
    "toro1": (1.0, 0.75, 1.0, 0.125, 0.0, 0.1, 0.5, 0.2),
    "toro6": (1.4, 0.0, 1.0, 1.0, 0.0, 1.0, 0.5, 2.0),
    "toro7": (1.4, 0.1, 1.0, 1.0, 0.1, 1.0, 0.5, 2.0),
    "sod": (1.0, 0.0, 1.0, 0.125, 0.0, 0.1, 0.5, 0.4),  # inputs/sod.in
}
METHODS = [("vl2", "plm", "hllc", 2), ("vl2", "plm", "hlle", 2), ("rk1", "dc", "hllc", 2),
           ("rk3", "ppm", "hllc", 3), ("rk3", "wenoz", "hlle", 3)]
# L1 error bounds of rho at nx1 = 256 (first-order convergence at discontinuities: generous but
# far below what a wrong wave speed / flux would give, ~1e-1)
BOUND = {"toro1": 3.5e-3, "sod": 3.0e-3, "toro6": 6.0e-3, "toro7": 6.0e-3}
BOUND_FIRST_ORDER = 1.2e-2  # rk1 + donor cell


def deck(name, integrator, recon, riemann, ng, nx=256):
    rl, ul, pl, rr, ur, pr, x0, tend = TUBES[name]
    kw = dict(fluid="euler", nx=(nx, 1, 1), mb=(64, 1, 1), ng=ng,
              bc=["outflow", "outflow", "periodic", "periodic", "periodic", "periodic"],
              xmin=(0.0, -0.5, -0.5), xmax=(1.0, 0.5, 0.5), gamma=GAMMA, cfl=0.5, recon=recon,
              riemann=riemann, integrator=integrator)
    return kw, [rl, pl, ul, rr, pr, ur, x0], tend


def l1_errors(prim_blocks, ng, name, nx=256):
    rl, ul, pl, rr, ur, pr, x0, tend = TUBES[name]
    w = np.concatenate([b[:, 0, 0, ng:-ng] for b in prim_blocks], axis=1)  # [nv][nx]
    x = (np.arange(nx) + 0.5) / nx
    ex = sample((x - x0) / tend, rl, ul, pl, rr, ur, pr, GAMMA)
    return [float(np.mean(np.abs(w[i] - ex[j]))) for i, j in ((0, 0), (1, 1), (4, 2))]


def test_exact_solver_known_star_states():
    # Toro table 4.2 (tests 1 of table 4.1 = Sod, and table 10.1 test 1)
    p, u = star_state(1.0, 0.0, 1.0, 0.125, 0.0, 0.1, 1.4)
    assert abs(p - 0.30313) < 1e-5 and abs(u - 0.92745) < 1e-5
    # Toro table 4.2 test 2 (123 problem: two rarefactions) and test 3 (strong left blast)
    p, u = star_state(1.0, -2.0, 0.4, 1.0, 2.0, 0.4, 1.4)
    assert abs(p - 0.00189) < 1e-5 and abs(u) < 1e-12
    p, u = star_state(1.0, 0.0, 1000.0, 1.0, 0.0, 0.01, 1.4)
    assert abs(p - 460.894) < 1e-3 and abs(u - 19.5975) < 1e-4


@pytest.mark.parametrize("name", list(TUBES))
@pytest.mark.parametrize("method", METHODS, ids=["-".join(m[:3]) for m in METHODS])
def test_oracle_shock_tube_vs_exact(name, method):
    integrator, recon, riemann, ng = method
    kw, pars, tend = deck(name, integrator, recon, riemann, ng)
    o = Oracle(**kw)
    o.pgen("sod", pars)
    o.initialize(tend)
    o.run(tend, 100000)
    assert abs(o.time - tend) < 1e-14
    err = l1_errors([o.prim(b) for b in range(4)], ng, name)
    assert err[0] < (BOUND_FIRST_ORDER if recon == "dc" else BOUND[name]), (name, method, err)
    if name in ("toro6", "toro7"):
        # isolated contacts: velocity and pressure stay uniform to rounding for every scheme
        assert err[1] < 1e-13 and err[2] < 1e-13, (name, method, err)
    if name == "toro6" and riemann == "hllc":
        # HLLC resolves a stationary contact exactly (the reason the suite exists)
        assert err[0] == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["toro1", "sod", "toro7"])
@pytest.mark.parametrize("fused", [True, False])
def test_cuda_shock_tube_vs_exact_and_oracle(name, fused):
    import athenapk_b200 as apk
    kw, pars, tend = deck(name, "vl2", "plm", "hllc", 2)
    o = Oracle(**kw)
    o.pgen("sod", pars)
    o.initialize(tend)
    o.run(tend, 100000)
    for fp_mode in ("strict", "fast"):
        s = apk.HydroSim(fp_mode=fp_mode, fused=fused, **kw)
        s.pgen("sod", pars)
        s.initialize(tend)
        s.execute(tend, 100000)
        assert s.ncycle == o.ncycle
        prim = [s.get("prim", b) for b in range(s.nblocks)]
        order = np.argsort([s.block_info(b)[1][0] for b in range(s.nblocks)])
        err = l1_errors([prim[b] for b in order], 2, name)
        assert err[0] < BOUND[name], (name, fp_mode, err)
        if fp_mode == "strict":
            for b in range(s.nblocks):
                ob = s.block_info(b)[1][0]
                assert np.array_equal(prim[b], o.prim(ob)), (name, b)
            assert s.dt == o.dt
        s.close()
