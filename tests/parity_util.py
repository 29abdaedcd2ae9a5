"""Helpers shared by the GPU parity tests: run the same deck through the CUDA product
(athenapk_b200.HydroSim, via the C-ABI) and through the CPU oracle, and compare."""
import numpy as np

from oracle_lib import Oracle

SIM_ONLY = {"fp_mode", "fused", "rank", "nranks"}
ORACLE_ONLY = {"fofc_jacobi"}


def make_pair(apk, pgen="random", pgen_pars=(), tlim=float("inf"), **kw):
    okw = {k: v for k, v in kw.items() if k not in SIM_ONLY}
    skw = {k: v for k, v in kw.items() if k not in ORACLE_ONLY}
    o = Oracle(**okw)
    s = apk.HydroSim(**skw)
    o.pgen(pgen, pgen_pars)
    s.pgen(pgen, pgen_pars)
    if pgen == "linear_wave" and tlim is None:
        tlim = o.lw_tlim
    o.initialize(tlim)
    s.initialize(tlim)
    return s, o


def oracle_block(s, o, b):
    loc = s.block_info(b)[1]
    if o.p.adaptive:  # the mesh changes: index by (level, logical location), rebuilt per block count
        if getattr(o, "_leaf_index_n", None) != (o.nblocks, o.ncycle):
            o._leaf_index = {o.block_loc(q): q for q in range(o.nblocks)}
            o._leaf_index_n = (o.nblocks, o.ncycle)
        return o._leaf_index[(s.block_level(b), tuple(loc))]
    if o.p.nref > 0:  # refined mesh: match blocks by (level, logical location)
        if not hasattr(o, "_leaf_index"):
            o._leaf_index = {o.block_loc(q): q for q in range(o.nblocks)}
        return o._leaf_index[(s.block_level(b), tuple(loc))]
    nb = [o.p.nx[d] // o.p.mb[d] for d in range(3)]
    return loc[0] + nb[0] * (loc[1] + nb[1] * loc[2])


def max_abs_diff(s, o, fields=("cons", "prim")):
    worst = 0.0
    for b in range(s.nblocks):
        ob = oracle_block(s, o, b)
        for f in fields:
            ref = o.cons(ob) if f == "cons" else o.prim(ob)
            worst = max(worst, float(np.max(np.abs(s.get(f, b) - ref))))
    return worst


def rel_l1(s, o, field="cons"):
    """relative L1 difference of the state: max over variables of sum|got-ref| normalised by the
    largest per-variable L1 norm of the reference state (rounding-level differences are absolute
    at the scale of the O(1) fields; normalising a 1e-6-amplitude momentum component by its own
    norm would ask for 1e-18 absolute agreement, below what FMA vs non-FMA arithmetic -- the
    reference's own CUDA vs CPU builds -- can deliver)."""
    num = None
    den = None
    for b in range(s.nblocks):
        ob = oracle_block(s, o, b)
        ref = o.cons(ob) if field == "cons" else o.prim(ob)
        got = s.get(field, b)
        n = np.abs(got - ref).reshape(ref.shape[0], -1).sum(axis=1)
        d = np.abs(ref).reshape(ref.shape[0], -1).sum(axis=1)
        num = n if num is None else num + n
        den = d if den is None else den + d
    return float(np.max(num) / max(float(den.max()), 1e-300))


def rel_l1_per_variable(s, o, field="cons"):
    """strict form of rel_l1: every variable normalised by ITS OWN L1 norm, max over variables.
    For a wave of amplitude 1e-6 on a background at rest this asks the momentum perturbations to
    agree to the stated tolerance RELATIVE TO THE PERTURBATION (variables whose reference norm is
    below 1e-300 are skipped: identically zero components)."""
    num = None
    den = None
    for b in range(s.nblocks):
        ob = oracle_block(s, o, b)
        ref = o.cons(ob) if field == "cons" else o.prim(ob)
        got = s.get(field, b)
        n = np.abs(got - ref).reshape(ref.shape[0], -1).sum(axis=1)
        d = np.abs(ref).reshape(ref.shape[0], -1).sum(axis=1)
        num = n if num is None else num + n
        den = d if den is None else den + d
    ok = den > 1e-300
    return float(np.max(num[ok] / den[ok]))


def assert_bitexact(s, o, what=""):
    for b in range(s.nblocks):
        ob = oracle_block(s, o, b)
        for f in ("cons", "prim"):
            ref = o.cons(ob) if f == "cons" else o.prim(ob)
            got = s.get(f, b)
            if not np.array_equal(got, ref):
                bad = np.argwhere(got != ref)
                raise AssertionError("%s: %s differs in block %d at %d entries, first %s: %r vs %r"
                                     % (what, f, b, len(bad), tuple(bad[0]), got[tuple(bad[0])],
                                        ref[tuple(bad[0])]))
