"""Reference-run golden traces (tests/golden/reference_traces.json, produced by running the
unmodified reference: tests/golden/make_reference_goldens.py) and the engine set-ups that reproduce
each case with the CPU oracle (oracle_lib.Oracle) or the CUDA product (athenapk_b200.HydroSim)."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_traces.json")
TRACES = json.load(open(GOLDEN))
GAMMA_DECK = float("1.666666666666667")  # the decks' literal; differs from 5/3 in the last bit


def lw_kwargs(case):
    N = case["N"] if "N" in case else 16
    return dict(nx=(2 * N, N, N), fluid=case.get("fluid", "glmmhd"),
                integrator=case.get("integrator", "vl2"), recon=case.get("recon", "ppm"),
                riemann=case.get("riemann", "hlld"), ng=case.get("ng", 3), gamma=GAMMA_DECK,
                xmin=(0.0, 0.0, 0.0), xmax=(3.0, 1.5, 1.5), cfl=0.3)


def lw_pgen(case):
    return case.get("problem", "linear_wave_mhd"), [case.get("wave_flag", 0), 1e-6, 0.0]


DECKS = {
    "sod": (dict(nx=(256, 1, 1), mb=(64, 1, 1), ng=2,
                 bc=["outflow", "outflow", "periodic", "periodic", "periodic", "periodic"],
                 xmin=(0.0, -0.5, -0.5), xmax=(1.0, 0.5, 0.5), gamma=1.4, cfl=0.5, recon="plm",
                 riemann="hllc", integrator="vl2"), "sod", [], 0.4),
    "orszag_tang_50": (dict(fluid="glmmhd", nx=(256, 256, 1), mb=(64, 256, 1), ng=3,
                            xmin=(-0.5,) * 3, xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.4,
                            recon="ppm", riemann="hlld", integrator="vl2", fofc=True),
                       "orszag_tang", [], 1.0),
    # FirstOrderFluxCorrect triggered: see make_reference_goldens.py
    "blast_fofc": (dict(fluid="euler", nx=(64, 64, 1), mb=(32, 32, 1), ng=3, xmin=(-0.5,) * 3,
                        xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.9, recon="ppm", riemann="hllc",
                        integrator="vl2", fofc=True, dfloor=1e-8, pfloor=1e-8),
                   "blast", [0.1, 0.001, 1.6e8, 1.0], 0.1),
    "uniform_mhd": (dict(fluid="glmmhd", nx=(64, 64, 64), mb=(32, 32, 32), ng=3,
                         xmin=(0.0,) * 3, xmax=(1.0,) * 3, gamma=1.0001, cfl=0.3, recon="ppm",
                         riemann="hlld", integrator="vl2"), "uniform", [], 1000.0),
    "orszag_tang_3d": (dict(fluid="glmmhd", nx=(64, 64, 64), mb=(32, 32, 32), ng=3,
                            xmin=(-0.5,) * 3, xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.3,
                            recon="ppm", riemann="hlld", integrator="vl2"),
                       "orszag_tang", [], 1000.0),
}


# <diffusion> decks (tests/golden/make_diffusion_goldens.py -> reference_traces_diffusion.json)
DIFF_TRACES = json.load(open(os.path.join(os.path.dirname(GOLDEN), "reference_traces_diffusion.json")))
_OT2 = dict(fluid="glmmhd", nx=(64, 64, 1), mb=(32, 32, 1), ng=3, xmin=(-0.5,) * 3, xmax=(0.5,) * 3,
            gamma=GAMMA_DECK, cfl=0.4, recon="ppm", riemann="hlld", integrator="vl2")
_OT3 = dict(fluid="glmmhd", nx=(32, 32, 32), mb=(16, 16, 16), ng=3, xmin=(-0.5,) * 3,
            xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.3, recon="ppm", riemann="hlld", integrator="vl2")


def _diff(integrator, cond=None, visc=None, ohm=None, ratio=-1.0, cfl=-1.0):
    kw = dict(diffint=integrator, rkl2_max_dt_ratio=ratio, cfl_diff=cfl)
    if cond:
        kw.update(conduction=cond[0], thermal_diff_coeff=cond[1])
    if visc:
        kw.update(viscosity="isotropic", mom_diff_coeff=visc)
    if ohm:
        kw.update(resistivity="ohmic", ohm_diff_coeff=ohm)
    return kw


DIFF_DECKS = {
    "ot2d_unsplit": (dict(_OT2, **_diff("unsplit", ("isotropic", 0.002), 0.002, 0.002)), "orszag_tang", [], 1.0),
    "ot2d_unsplit_stiff": (dict(_OT2, **_diff("unsplit", ("isotropic", 0.03), 0.05, 0.04)), "orszag_tang", [], 1.0),
    "ot2d_rkl2": (dict(_OT2, **_diff("rkl2", ("isotropic", 0.05), 0.05, 0.05)), "orszag_tang", [], 1.0),
    "ot2d_rkl2_aniso": (dict(_OT2, **_diff("rkl2", ("anisotropic", 0.2), 0.02, 0.03, ratio=6.0, cfl=0.25)),
                        "orszag_tang", [], 1.0),
    "ot3d_unsplit": (dict(_OT3, **_diff("unsplit", ("isotropic", 0.004), 0.003, 0.005)), "orszag_tang", [], 1000.0),
    "ot3d_rkl2_aniso": (dict(_OT3, **_diff("rkl2", ("anisotropic", 0.1), 0.04, 0.05)), "orszag_tang", [], 1000.0),
    "blast2d_rkl2_hydro": (dict(fluid="euler", nx=(64, 64, 1), mb=(32, 32, 1), ng=2, xmin=(-0.5,) * 3,
                                xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.3, recon="plm", riemann="hllc",
                                integrator="rk3", **_diff("rkl2", ("isotropic", 0.02), 0.03)),
                           "blast", [0.1, 0.001, 1.6e8, 1.0], 0.1),
    "sod_unsplit": (dict(nx=(256, 1, 1), mb=(64, 1, 1), ng=2,
                         bc=["outflow", "outflow", "periodic", "periodic", "periodic", "periodic"],
                         xmin=(0.0, -0.5, -0.5), xmax=(1.0, 0.5, 0.5), gamma=1.4, cfl=0.5, recon="plm",
                         riemann="hllc", integrator="vl2", **_diff("unsplit", ("isotropic", 0.001), 0.002)),
                    "sod", [], 0.4),
}


def _amr(regions, fluid, recon, rs, ng, **extra):
    return (dict(fluid=fluid, nx=(32, 32, 32), mb=(8, 8, 8), ng=ng, xmin=(-0.5,) * 3,
                 xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.3, recon=recon, riemann=rs,
                 integrator="vl2",
                 refine=[(lev, (lo,) * 3, (hi,) * 3) for lev, lo, hi in regions], **extra),
            "blast", [0.12, 0.1, 100.0, 1.0, 0.22, 0.1, -0.05], 0.1)


# diffusion on refined meshes (make_diffusion_goldens.py): diffusive fluxes through the flux
# correction, RKL2 sub-steps through the multilevel ghost exchange
DIFF_DECKS.update({
    "amr_2level_rkl2_hydro": _amr([(1, -0.2, 0.2)], "euler", "plm", "hlle", 2,
                                  **_diff("rkl2", ("isotropic", 0.02), 0.03)),
    "amr_3level_unsplit_mhd": _amr([(1, -0.2, 0.2), (2, -0.1, 0.1)], "glmmhd", "plm", "hlld", 2,
                                   **_diff("unsplit", ("isotropic", 0.0002), 0.0003, 0.0002)),
    "amr_3level_rkl2_mhd_aniso": _amr([(1, -0.2, 0.2), (2, -0.1, 0.1)], "glmmhd", "wenoz", "hlld", 4,
                                      **_diff("rkl2", ("anisotropic", 0.01), 0.004, 0.005)),
})


# static mesh refinement (see make_reference_goldens.py): 2 and 3 levels
AMR_DECKS = {
    "amr_2level_plm_hlle": _amr([(1, -0.2, 0.2)], "euler", "plm", "hlle", 2),
    "amr_3level_plm_hlle": _amr([(1, -0.2, 0.2), (2, -0.1, 0.1)], "euler", "plm", "hlle", 2),
    "amr_3level_wenoz_hlld": _amr([(1, -0.2, 0.2), (2, -0.1, 0.1)], "glmmhd", "wenoz", "hlld", 4),
}


# adaptive mesh refinement (tests/golden/make_amr_goldens.py -> reference_traces_amr.json):
# inputs/blast_3d_amr.in with refinement = adaptive, <refinement> type = pressure_gradient
ADAPTIVE_TRACES = json.load(open(os.path.join(os.path.dirname(GOLDEN), "reference_traces_amr.json")))
_BLAST_OFF = [0.12, 0.1, 100.0, 1.0, 0.22, 0.1, -0.05]


def _adaptive(fluid, recon, rs, ng, numlevel, threshold=0.1, derefine_count=10, nx=(32, 32, 32),
              mb=(8, 8, 8), **criterion):
    return (dict(fluid=fluid, nx=nx, mb=mb, ng=ng, xmin=(-0.5,) * 3, xmax=(0.5,) * 3,
                 gamma=GAMMA_DECK, cfl=0.3, recon=recon, riemann=rs, integrator="vl2",
                 adaptive=dict(numlevel=numlevel, threshold=threshold, derefine_count=derefine_count,
                               **criterion)),
            "blast", _BLAST_OFF, 0.1)


ADAPTIVE_DECKS = {
    "adaptive_3d_plm_hlle": _adaptive("euler", "plm", "hlle", 2, 3),
    "adaptive_3d_wenoz_hlld": _adaptive("glmmhd", "wenoz", "hlld", 4, 3),
    "adaptive_2d_plm_hllc": _adaptive("euler", "plm", "hllc", 2, 4, threshold=0.2, derefine_count=3,
                                      nx=(64, 64, 1), mb=(8, 8, 1)),
    # the other <refinement> criteria: x1-x2 velocity gradient, maximum density
    "adaptive_2d_xyvelocity": _adaptive("euler", "plm", "hlle", 2, 3, threshold=0.05, derefine_count=5,
                                        nx=(64, 64, 1), mb=(8, 8, 1), type="xyvelocity_gradient"),
    "adaptive_2d_maxdensity": _adaptive("euler", "plm", "hllc", 2, 3, threshold=1.3, derefine_count=5,
                                        nx=(64, 64, 1), mb=(8, 8, 1), type="maxdensity",
                                        threshold2=1.1),
}


# GLM-MHD with B != 0 on an adaptive mesh: inputs/orszag_tang.in (FOFC on) with refinement = adaptive
ADAPTIVE_DECKS["adaptive_2d_orszag_tang_mhd"] = (
    dict(fluid="glmmhd", nx=(64, 64, 1), mb=(8, 8, 1), ng=4, xmin=(-0.5,) * 3, xmax=(0.5,) * 3,
         gamma=GAMMA_DECK, cfl=0.4, recon="ppm", riemann="hlld", integrator="vl2", fofc=True,
         adaptive=dict(numlevel=3, threshold=0.05, derefine_count=5)), "orszag_tang", [], 1.0)

# a three-stage integrator and RKL2 super time stepping on adaptive meshes
_A2D = dict(fluid="euler", nx=(64, 64, 1), mb=(8, 8, 1), ng=2, xmin=(-0.5,) * 3, xmax=(0.5,) * 3,
            gamma=GAMMA_DECK, cfl=0.3, adaptive=dict(numlevel=3, threshold=0.1, derefine_count=4))
ADAPTIVE_DECKS["adaptive_2d_rk3_weno3_hllc"] = (
    dict(_A2D, recon="weno3", riemann="hllc", integrator="rk3"), "blast", _BLAST_OFF, 0.1)
ADAPTIVE_DECKS["adaptive_2d_rkl2_diffusion"] = (
    dict(_A2D, recon="plm", riemann="hlle", integrator="vl2", diffint="rkl2", conduction="isotropic",
         thermal_diff_coeff=0.02, viscosity="isotropic", mom_diff_coeff=0.03), "blast", _BLAST_OFF, 0.1)

# refined meshes on non-periodic domains (physical BCs on coarse buffers and blocks)
_OUT4 = ["outflow"] * 4 + ["periodic"] * 2
BOUNDARY_AMR_DECKS = {
    "static_2d_outflow": (dict(fluid="euler", nx=(64, 64, 1), mb=(8, 8, 1), ng=2, xmin=(-0.5,) * 3,
                               xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.3, recon="plm", riemann="hlle",
                               integrator="vl2", bc=_OUT4,
                               refine=[(2, (-0.5, -0.5, -0.5), (-0.2, -0.1, 0.5))]),
                          "blast", [0.12, 0.1, 100.0, 1.0, -0.42, -0.4, 0.0], 0.1),
    "adaptive_2d_reflect_outflow": (dict(fluid="euler", nx=(64, 64, 1), mb=(8, 8, 1), ng=2,
                                         xmin=(-0.5,) * 3, xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.3,
                                         recon="plm", riemann="hllc", integrator="vl2",
                                         bc=["reflecting", "outflow", "outflow", "reflecting",
                                             "periodic", "periodic"],
                                         adaptive=dict(numlevel=3, threshold=0.1, derefine_count=5)),
                                    "blast", [0.12, 0.1, 100.0, 1.0, -0.42, 0.41, 0.0], 0.1),
    "static_3d_outflow_wenoz_hlld": (dict(fluid="glmmhd", nx=(32, 32, 32), mb=(8, 8, 8), ng=4,
                                          xmin=(-0.5,) * 3, xmax=(0.5,) * 3, gamma=GAMMA_DECK, cfl=0.3,
                                          recon="wenoz", riemann="hlld", integrator="vl2",
                                          bc=["outflow"] * 6,
                                          refine=[(2, (0.1, -0.5, -0.2), (0.5, -0.2, 0.2))]),
                                     "blast", [0.12, 0.1, 100.0, 1.0, 0.4, -0.41, 0.05], 0.1),
}


def follow_adaptive_trace(engine, case, tlim, advance, counts, max_cycles=None):
    """like follow_trace, plus the number of meshblocks (total and per level) the reference prints
    every cycle; counts(engine) -> (created, destroyed, [blocks per level]).  max_cycles: follow
    only the first cycles of the trace (CPU tests of long 3-D decks; the GPU tests follow all)"""
    for c, (t, dt) in enumerate(case["trace"]):
        if max_cycles is not None and c > max_cycles:
            return
        if c > 0:
            advance(engine, tlim, c)
        got = ("%.16e" % engine.time, "%.16e" % engine.dt)
        assert got == (t, dt), "cycle %d: got %s, reference printed %s" % (c, got, (t, dt))
        created, destroyed, per_level = counts(engine)
        want = case["blocks"][c]
        assert [engine.nblocks] + per_level[:len(want) - 1] == want, (c, per_level, want)
    created, destroyed, _ = counts(engine)
    # Mesh::nbnew / nbdel also count the refinement during Mesh::Initialize (static meshes print
    # no totals)
    if case["final"]:
        assert [engine.nblocks, created, destroyed] == case["final"]


def follow_trace(engine, trace, tlim, advance, every=1):
    """advance(engine, tlim, ncycle) runs the engine up to cycle `ncycle`; compares (time, dt) with
    the reference's printout at all 17 digits.  Returns the number of cycles compared."""
    n = 0
    for c, (t, dt) in enumerate(trace):
        if c % every and c != len(trace) - 1:
            continue
        if c > 0:
            advance(engine, tlim, c)
        got = ("%.16e" % engine.time, "%.16e" % engine.dt)
        assert got == (t, dt), "cycle %d: got %s, reference printed %s" % (c, got, (t, dt))
        n += 1
    return n


def check_lw_row(engine, case):
    """the row the reference appended to linearwave-errors.dat: Nx1 Nx2 Nx3 Ncycle RMS L1[...] ..."""
    row = case["lw_row"]
    nerr = 8 if case.get("problem", "linear_wave_mhd") == "linear_wave_mhd" else 5
    e = engine.linear_wave_errors()
    assert engine.ncycle == int(row[3])
    got = ["%e" % x for x in e[:1 + nerr]]
    assert got == row[4:5 + nerr], (got, row[4:5 + nerr])
