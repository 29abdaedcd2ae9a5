"""GPU parity of the AMR building blocks through the C-ABI: restriction of cells
and face fluxes, AthenaPK's multi-D limited prolongation and the pressure-gradient tagger against
the CPU oracle (itself pinned to the reference's own structs, tests/test_oracle_amr.py).
Strict FP mode: bit for bit; FMA mode: <= 1e-14 relative."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from parity_util import assert_bitexact, make_pair
from trace_util import ADAPTIVE_DECKS, ADAPTIVE_TRACES, BOUNDARY_AMR_DECKS, follow_adaptive_trace

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def apk():
    import athenapk_b200 as m
    m.lib()
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_items(apk, pairs, ndim, ng, xmin, dx, los, his):
    items = (apk.AmrItem * len(pairs))()
    for it, (f, c), lo, hi in zip(items, pairs, los, his):
        it.fine, it.coarse = f.data_ptr(), c.data_ptr()
        it.nf[:] = list(f.shape[:0:-1])
        it.nc[:] = list(c.shape[:0:-1])
        it.lo[:] = lo
        it.hi[:] = hi
        it.xmin[:] = xmin
        it.dx[:] = dx
    raw = np.frombuffer(bytes(items), dtype=np.uint8).copy()
    return torch.from_numpy(raw).cuda()


def shapes(ndim, nx, ng):
    nf = [nx + 2 * ng if d < ndim else 1 for d in range(3)]
    nc = [nx // 2 + 2 * ng if d < ndim else 1 for d in range(3)]
    return nf, nc


@pytest.mark.parametrize("fp_mode", [1, 0])
@pytest.mark.parametrize("ndim,el", [(1, 0), (2, 0), (3, 0), (2, 1), (3, 1), (3, 2), (3, 3)])
def test_restrict_average(apk, ndim, el, fp_mode):
    L = apk.lib()
    rng = np.random.default_rng(ndim * 10 + el)
    nx, ng, nv, nitems = 16, 3, 9, 5
    nf, nc = shapes(ndim, nx, ng)
    xmin, dx = (0.1, -0.3, 0.7), (0.013, 0.017, 0.011)
    lo = [ng if d < ndim else 0 for d in range(3)]
    hi = [ng + nx // 2 - 1 if d < ndim else 0 for d in range(3)]
    if el:
        hi[el - 1] += 1
    fines = [rng.standard_normal((nv, nf[2], nf[1], nf[0])) for _ in range(nitems)]
    pairs = [(dev(f), dev(np.full((nv, nc[2], nc[1], nc[0]), np.nan))) for f in fines]
    items = make_items(apk, pairs, ndim, ng, xmin, dx, [lo] * nitems, [hi] * nitems)
    rc = L.apk_restrict_average(ndim, ng, nv, el, C.c_void_p(items.data_ptr()), nitems, fp_mode,
                                None)
    assert rc == 0, L.apk_last_error()
    torch.cuda.synchronize()
    for f, (_, c) in zip(fines, pairs):
        ref = np.full((nv, nc[2], nc[1], nc[0]), np.nan)
        oracle_lib.restrict_average(ndim, ng, el, xmin, dx, f, ref, lo, hi)
        got = c.cpu().numpy()
        if fp_mode == 1:
            assert np.array_equal(got, ref, equal_nan=True)
        else:
            assert np.allclose(got, ref, rtol=1e-14, atol=1e-14, equal_nan=True)


@pytest.mark.parametrize("fp_mode", [1, 0])
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_prolongate_minmod_multid(apk, ndim, fp_mode):
    L = apk.lib()
    rng = np.random.default_rng(ndim)
    nx, ng, nv, nitems = 16, 3, 9, 4
    nf, nc = shapes(ndim, nx, ng)
    xmin, dx = (0.25, 1.5, -2.0), (1.0 / 64, 1.0 / 48, 1.0 / 80)
    lo = [max(1, ng - ng // 2) if d < ndim else 0 for d in range(3)]
    hi = [min(nc[d] - 2, (nf[d] - 2 - ng) // 2 + ng) if d < ndim else 0 for d in range(3)]
    coarses = [np.exp(0.3 * rng.standard_normal((nv, nc[2], nc[1], nc[0]))) for _ in range(nitems)]
    coarses[1] = np.ascontiguousarray(np.cumsum(coarses[1], axis=3))  # monotone along x1
    pairs = [(dev(np.full((nv, nf[2], nf[1], nf[0]), np.nan)), dev(c)) for c in coarses]
    items = make_items(apk, pairs, ndim, ng, xmin, dx, [lo] * nitems, [hi] * nitems)
    rc = L.apk_prolongate_minmod_multid(ndim, ng, nv, C.c_void_p(items.data_ptr()), nitems, fp_mode,
                                        None)
    assert rc == 0, L.apk_last_error()
    torch.cuda.synchronize()
    for c, (f, _) in zip(coarses, pairs):
        ref = np.full((nv, nf[2], nf[1], nf[0]), np.nan)
        oracle_lib.prolongate_minmod_multid(ndim, ng, xmin, dx, ref, np.ascontiguousarray(c), lo, hi)
        got = f.cpu().numpy()
        if fp_mode == 1:
            assert np.array_equal(got, ref, equal_nan=True)
        else:
            assert np.allclose(got, ref, rtol=1e-13, atol=1e-13, equal_nan=True)


def test_restrict_prolongate_roundtrip_large(apk):
    """size-independent property at a production block size: restriction of the prolongation
    returns the coarse data (children average to the parent)."""
    L = apk.lib()
    ndim, nx, ng, nv = 3, 32, 3, 9
    nf, nc = shapes(ndim, nx, ng)
    rng = np.random.default_rng(5)
    coarse = rng.standard_normal((nv, nc[2], nc[1], nc[0]))
    lo, hi = [ng] * 3, [ng + nx // 2 - 1] * 3
    f, c = dev(np.zeros((nv, nf[2], nf[1], nf[0]))), dev(coarse)
    back = dev(np.zeros_like(coarse))
    xmin, dx = (0.0, 0.0, 0.0), (1.0 / 32,) * 3
    it1 = make_items(apk, [(f, c)], ndim, ng, xmin, dx, [lo], [hi])
    it2 = make_items(apk, [(f, back)], ndim, ng, xmin, dx, [lo], [hi])
    assert L.apk_prolongate_minmod_multid(ndim, ng, nv, C.c_void_p(it1.data_ptr()), 1, 1, None) == 0
    assert L.apk_restrict_average(ndim, ng, nv, 0, C.c_void_p(it2.data_ptr()), 1, 1, None) == 0
    torch.cuda.synchronize()
    sl = (slice(None),) + (slice(ng, ng + nx // 2),) * 3
    assert np.allclose(back.cpu().numpy()[sl], coarse[sl], rtol=0, atol=1e-13)


def test_pressure_gradient_tags(apk):
    kw = dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(32, 32, 32),
              mb=(16, 16, 16), ng=3, xmin=(0.0, 0.0, 0.0), xmax=(1.0, 1.0, 1.0), gamma=5.0 / 3.0,
              cfl=0.3, fp_mode="strict")
    s = apk.HydroSim(**kw)
    s.pgen("ot3d")
    s.initialize()
    s.execute(None, 3)
    L = apk.lib()
    nb = s.nblocks
    tags = torch.zeros(nb, dtype=torch.int32, device="cuda")
    eps = torch.zeros(nb, dtype=torch.float64, device="cuda")
    ref = [oracle_lib.pressure_gradient_tag(3, 3, s.get("prim", b), 4, 0.0) for b in range(nb)]
    thr = float(np.median([m for _, m in ref]))
    rc = L.apk_tag_pressure_gradient(C.c_void_p(L.apk_sim_pack_desc(s.h)),
                                     C.c_void_p(L.apk_sim_device_blocks(s.h)), C.c_double(thr),
                                     C.c_void_p(tags.data_ptr()), C.c_void_p(eps.data_ptr()),
                                     C.c_void_p(s.stream))
    assert rc == 0, L.apk_last_error()
    s.synchronize()
    got_eps, got_tags = eps.cpu().numpy(), tags.cpu().numpy()
    for b in range(nb):
        tag, m = oracle_lib.pressure_gradient_tag(3, 3, s.get("prim", b), 4, thr)
        assert got_eps[b] == m, (b, got_eps[b], m)
        assert got_tags[b] == tag
    assert set(got_tags.tolist()) >= {1}  # the threshold is the median: some blocks refine
    s.close()


# ---- adaptive mesh refinement end to end (BASELINE config 5 as the reference runs it) -------------
@pytest.mark.parametrize("name", sorted(ADAPTIVE_DECKS) + sorted(BOUNDARY_AMR_DECKS))
def test_adaptive_refinement_reference_traces(apk, name):
    """inputs/blast_3d_amr.in with refinement = adaptive on the GPU, strict mode: device-side
    pressure-gradient tags, host tree update, prolongation of new fine blocks / restriction into new
    coarse blocks on the device, rebuilt multilevel exchange.  Every cycle's (time, dt) equals the
    unmodified reference's printout at all 17 digits, the meshblock counts per level equal the ones
    it prints every cycle, the created / destroyed totals match, and the final state -- same
    meshblocks in the same order -- equals the oracle's bit for bit."""
    # (BOUNDARY_AMR_DECKS: static and adaptive refinement on NON-periodic domains -- outflow and
    # reflecting walls applied to the coarse buffers before the prolongation and to the blocks after)
    kw, pgen, pars, tlim = dict(ADAPTIVE_DECKS, **BOUNDARY_AMR_DECKS)[name]
    s, o = make_pair(apk, pgen, pars, tlim=tlim, fp_mode="strict", **kw)
    case = ADAPTIVE_TRACES[name]
    assert not s.fused_active
    follow_adaptive_trace(s, case, tlim, lambda s, t, c: s.execute(t, c), lambda s: s.amr_counts())
    assert s.ncycle == case["ncycle"]
    o.run(tlim, case["ncycle"])
    assert (s.time, s.dt, s.nblocks) == (o.time, o.dt, o.nblocks)
    leaves = [(s.block_level(b),) + tuple(s.block_info(b)[1]) for b in range(s.nblocks)]
    oleaves = [(o.block_loc(b)[0],) + o.block_loc(b)[1] for b in range(o.nblocks)]
    if kw.get("adaptive"):
        assert leaves == oleaves  # same meshblocks in the same (Parthenon) order
    else:
        assert sorted(leaves) == sorted(oleaves)
    assert_bitexact(s, o, name)
    s.close()


def test_adaptive_refinement_fast_mode_conserves(apk):
    """the benchmarked FP mode on the adaptive config-5 deck: restriction, prolongation and the flux
    correction are conservative, so total mass / momentum / energy stay put (1e-13 relative) while
    the mesh goes from 435 to ~580 meshblocks; the block count follows the reference's within 5 %."""
    kw, pgen, pars, tlim = ADAPTIVE_DECKS["adaptive_3d_wenoz_hlld"]
    s = apk.HydroSim(fp_mode="fast", **kw)
    s.pgen(pgen, pars)
    s.initialize(tlim)
    case = ADAPTIVE_TRACES["adaptive_3d_wenoz_hlld"]
    assert s.nblocks == case["blocks"][0][0]
    h0 = s.history()
    s.execute(tlim, case["ncycle"])
    h1 = s.history()
    assert abs(h1[0] - h0[0]) <= 1e-13 * abs(h0[0])          # mass
    assert abs(h1[5] - h0[5]) <= 1e-13 * abs(h0[5])          # total energy
    assert np.max(np.abs(h1[1:4] - h0[1:4])) <= 1e-13 * abs(h0[0])
    assert abs(s.nblocks - case["final"][0]) <= 0.05 * case["final"][0]
    assert abs(s.dt - float(case["trace"][-1][1])) <= 1e-6 * s.dt
    s.close()
