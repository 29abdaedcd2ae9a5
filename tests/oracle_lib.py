"""ctypes binding of the CPU oracle (oracle/libapk_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")

FLUID = {"euler": 0, "glmmhd": 1}
RECON = {"dc": 0, "plm": 1, "ppm": 2, "wenoz": 3, "weno3": 4, "limo3": 5}
RIEMANN = {"none": 0, "hlle": 1, "llf": 2, "hllc": 3, "hlld": 4}
INTEGRATOR = {"rk1": 0, "rk2": 1, "vl2": 2, "rk3": 3}
BC = {"periodic": 0, "outflow": 1, "reflecting": 2}
PGEN = {"linear_wave": 0, "sod": 1, "orszag_tang": 2, "uniform": 3, "ot3d": 4, "blast": 5,
        "random": 6, "linear_wave_mhd": 7, "lw_implode": 8, "diffusion": 9}


class Params(C.Structure):
    _fields_ = [
        ("fluid", C.c_int), ("recon", C.c_int), ("riemann", C.c_int), ("integrator", C.c_int),
        ("nx", C.c_int * 3), ("mb", C.c_int * 3), ("ng", C.c_int), ("bc", C.c_int * 6),
        ("xmin", C.c_double * 3), ("xmax", C.c_double * 3),
        ("gamma", C.c_double), ("cfl", C.c_double),
        ("dfloor", C.c_double), ("pfloor", C.c_double), ("efloor", C.c_double),
        ("vceil", C.c_double), ("eceil", C.c_double),
        ("fofc", C.c_int), ("fofc_jacobi", C.c_int), ("dedner_extended", C.c_int),
        ("glmmhd_alpha", C.c_double), ("nscalars", C.c_int),
        ("nref", C.c_int), ("ref_level", C.c_int * 4),
        ("ref_xmin", (C.c_double * 3) * 4), ("ref_xmax", (C.c_double * 3) * 4),
        ("diffint", C.c_int), ("conduction", C.c_int), ("viscosity", C.c_int),
        ("resistivity", C.c_int),
        ("thermal_diff_coeff", C.c_double), ("mom_diff_coeff", C.c_double),
        ("ohm_diff_coeff", C.c_double), ("cfl_diff", C.c_double),
        ("rkl2_max_dt_ratio", C.c_double), ("conduction_sat_phi", C.c_double),
        ("adaptive", C.c_int), ("numlevel", C.c_int), ("derefine_count", C.c_int),
        ("refine_threshold", C.c_double),
        ("refine_type", C.c_int), ("refine_threshold2", C.c_double),
    ]


DIFFINT = {"none": 0, "unsplit": 1, "rkl2": 2}
# <refinement> type; threshold = threshold_pressure_gradient / threshold_xyvelocity_gradient /
# maxdensity_refine_above, threshold2 = maxdensity_deref_below
REFINE_TYPE = {"pressure_gradient": 0, "xyvelocity_gradient": 1, "maxdensity": 2}
CONDUCTION = {"none": 0, "isotropic": 1, "anisotropic": 2}


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libapk_oracle.so"])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ORACLE_DIR, "libapk_oracle.so")
    src = os.path.join(ORACLE_DIR, "apk_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build()
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    L.apko_default_params.argtypes = [C.POINTER(Params)]
    L.apko_create.argtypes = [C.POINTER(Params)]
    L.apko_create.restype = C.c_void_p
    L.apko_destroy.argtypes = [C.c_void_p]
    L.apko_pgen.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
    L.apko_initialize.argtypes = [C.c_void_p, C.c_double]
    L.apko_run.argtypes = [C.c_void_p, C.c_double, C.c_int]
    L.apko_cycle.argtypes = [C.c_void_p]
    L.apko_num_blocks.argtypes = [C.c_void_p]
    L.apko_amr_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.apko_amr_tag.argtypes = [C.c_void_p]
    L.apko_amr_remesh.argtypes = [C.c_void_p]
    L.apko_amr_flag.argtypes = [C.c_void_p, C.c_int]
    L.apko_amr_last_input.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.apko_amr_last_input.restype = C.POINTER(C.c_int)
    L.apko_block_loc.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.apko_block_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.apko_nvar.argtypes = [C.c_void_p]
    for f in ("apko_cons", "apko_prim"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        getattr(L, f).restype = dp
    L.apko_flux.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.apko_flux.restype = dp
    L.apko_block_coords.argtypes = [C.c_void_p, C.c_int, dp, dp]
    for f in ("apko_time", "apko_dt", "apko_c_h", "apko_lw_tlim", "apko_estimate_dt"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = C.c_double
    L.apko_ncycle.argtypes = [C.c_void_p]
    L.apko_fofc_total.argtypes = [C.c_void_p]
    L.apko_fofc_total.restype = C.c_longlong
    L.apko_linear_wave_errors.argtypes = [C.c_void_p, dp]
    L.apko_history.argtypes = [C.c_void_p, dp]
    L.apko_calc_fluxes.argtypes = [C.c_void_p, C.c_int]
    L.apko_fofc.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    L.apko_fofc.restype = C.c_longlong
    L.apko_flux_div_update.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    L.apko_dedner.argtypes = [C.c_void_p, C.c_double]
    L.apko_exchange.argtypes = [C.c_void_p]
    L.apko_cons_to_prim.argtypes = [C.c_void_p]
    L.apko_save_u1.argtypes = [C.c_void_p]
    L.apko_set_c_h.argtypes = [C.c_void_p, C.c_double]
    L.apko_reset_fluxes.argtypes = [C.c_void_p]
    L.apko_calc_diff_fluxes.argtypes = [C.c_void_p]
    L.apko_estimate_dt_diff.argtypes = [C.c_void_p]
    L.apko_estimate_dt_diff.restype = C.c_double
    L.apko_sts.argtypes = [C.c_void_p, C.c_double]
    L.apko_last_s_rkl.argtypes = [C.c_void_p]
    L.apko_recon_scalar.argtypes = [C.c_int, dp, C.c_double, C.c_int, dp, dp]
    L.apko_riemann_face.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_double, C.c_double, dp]
    L.apko_c2p_cell.argtypes = [C.c_int, C.c_int, dp, dp] + [C.c_double] * 6
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_params(fluid="euler", recon="plm", riemann="hlle", integrator="vl2", nx=(64, 32, 32),
                mb=None, ng=2, bc="periodic", xmin=(0.0, 0.0, 0.0), xmax=(3.0, 1.5, 1.5),
                gamma=5.0 / 3.0, cfl=0.3, fofc=False, fofc_jacobi=False, dedner_extended=False,
                glmmhd_alpha=0.1, nscalars=0, dfloor=-1.0, pfloor=-1.0, efloor=-1.0,
                vceil=float("inf"), eceil=float("inf"), refine=(), diffint="none",
                conduction="none", thermal_diff_coeff=0.0, viscosity="none", mom_diff_coeff=0.0,
                resistivity="none", ohm_diff_coeff=0.0, cfl_diff=-1.0, rkl2_max_dt_ratio=-1.0,
                adaptive=None):
    """refine: static refinement regions [(level, (x1min, x2min, x3min), (x1max, x2max, x3max)), ...]
    = the <parthenon/static_refinementN> blocks of a deck.
    adaptive: dict(numlevel=, threshold=, derefine_count=10) = parthenon/mesh/refinement = adaptive
    with <refinement> type = pressure_gradient"""
    p = Params()
    lib().apko_default_params(C.byref(p))
    p.fluid, p.recon, p.riemann = FLUID[fluid], RECON[recon], RIEMANN[riemann]
    p.integrator = INTEGRATOR[integrator]
    mb = mb or nx
    for d in range(3):
        p.nx[d], p.mb[d] = nx[d], mb[d]
        p.xmin[d], p.xmax[d] = xmin[d], xmax[d]
    p.ng = ng
    bcs = [bc] * 6 if isinstance(bc, str) else list(bc)
    for f in range(6):
        p.bc[f] = BC[bcs[f]]
    p.gamma, p.cfl = gamma, cfl
    p.fofc, p.fofc_jacobi, p.dedner_extended = int(fofc), int(fofc_jacobi), int(dedner_extended)
    p.glmmhd_alpha, p.nscalars = glmmhd_alpha, nscalars
    p.dfloor, p.pfloor, p.efloor, p.vceil, p.eceil = dfloor, pfloor, efloor, vceil, eceil
    p.diffint, p.conduction = DIFFINT[diffint], CONDUCTION[conduction]
    p.viscosity, p.resistivity = int(viscosity == "isotropic"), int(resistivity == "ohmic")
    p.thermal_diff_coeff, p.mom_diff_coeff, p.ohm_diff_coeff = thermal_diff_coeff, mom_diff_coeff, ohm_diff_coeff
    p.cfl_diff, p.rkl2_max_dt_ratio = cfl_diff, rkl2_max_dt_ratio
    if adaptive:
        p.adaptive, p.numlevel = 1, adaptive["numlevel"]
        p.derefine_count = adaptive.get("derefine_count", 10)
        p.refine_threshold = adaptive["threshold"]
        p.refine_type = REFINE_TYPE[adaptive.get("type", "pressure_gradient")]
        p.refine_threshold2 = adaptive.get("threshold2", 0.0)
    p.nref = len(refine)
    for r, (lev, lo, hi) in enumerate(refine):
        p.ref_level[r] = lev
        for d in range(3):
            p.ref_xmin[r][d], p.ref_xmax[r][d] = lo[d], hi[d]
    return p


class Oracle:
    """Thin object wrapper around an apko_engine."""

    def __init__(self, **kw):
        self.L = lib()
        self.p = make_params(**kw)
        self.h = self.L.apko_create(C.byref(self.p))
        if not self.h:
            raise ValueError("apko_create failed (mesh not divisible by meshblock?)")
        n = (C.c_int * 3)()
        self.L.apko_block_dims(self.h, n)
        self.n = tuple(n)  # (n1, n2, n3)
        self.nvar = self.L.apko_nvar(self.h)

    @property
    def nblocks(self):
        return self.L.apko_num_blocks(self.h)

    def amr_counts(self):
        """(blocks created, blocks destroyed, [blocks per level])"""
        a, b, lv = C.c_int(), C.c_int(), (C.c_int * 8)()
        self.L.apko_amr_counts(self.h, C.byref(a), C.byref(b), lv)
        return a.value, b.value, list(lv)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.apko_destroy(self.h)
            self.h = None

    def amr_last_input(self):
        """(leaves [(level, lx1, lx2, lx3)], flags) the last remesh started from"""
        n = C.c_int()
        p = self.L.apko_amr_last_input(self.h, C.byref(n))
        a = np.ctypeslib.as_array(p, shape=(n.value, 5)).copy() if n.value else np.zeros((0, 5), int)
        return [tuple(int(v) for v in r[:4]) for r in a], [int(r[4]) for r in a]

    def block_loc(self, b):
        """(refinement level, (lx1, lx2, lx3)) of block b"""
        lev = C.c_int()
        lx = (C.c_int * 3)()
        self.L.apko_block_loc(self.h, b, C.byref(lev), lx)
        return lev.value, tuple(lx)

    def pgen(self, name, pars=()):
        a = np.asarray(pars, dtype=np.float64)
        rc = self.L.apko_pgen(self.h, PGEN[name], _dp(a) if a.size else None, a.size)
        if rc != 0:
            raise ValueError("pgen failed")

    def initialize(self, tlim=float("inf")):
        self.L.apko_initialize(self.h, 1.0e300 if tlim == float("inf") else tlim)

    def run(self, tlim=None, nlim=-1):
        t = self.L.apko_time(self.h) + 1e300 if tlim is None else tlim
        return self.L.apko_run(self.h, t, nlim)

    def cycle(self):
        self.L.apko_cycle(self.h)

    def _arr(self, ptr):
        shape = (self.nvar, self.n[2], self.n[1], self.n[0])
        return np.ctypeslib.as_array(ptr, shape=shape)

    def cons(self, b):
        return self._arr(self.L.apko_cons(self.h, b))

    def prim(self, b):
        return self._arr(self.L.apko_prim(self.h, b))

    def flux(self, b, d):
        return self._arr(self.L.apko_flux(self.h, b, d))

    def all_cons(self):
        return np.stack([self.cons(b).copy() for b in range(self.nblocks)])

    def all_prim(self):
        return np.stack([self.prim(b).copy() for b in range(self.nblocks)])

    def set_cons(self, arr):
        for b in range(self.nblocks):
            self.cons(b)[...] = arr[b]

    @property
    def time(self):
        return self.L.apko_time(self.h)

    @property
    def dt(self):
        return self.L.apko_dt(self.h)

    @property
    def ncycle(self):
        return self.L.apko_ncycle(self.h)

    @property
    def fofc_total(self):
        return self.L.apko_fofc_total(self.h)

    @property
    def c_h(self):
        return self.L.apko_c_h(self.h)

    @property
    def lw_tlim(self):
        return self.L.apko_lw_tlim(self.h)

    def linear_wave_errors(self):
        """[RMS-L1, per-variable L1...]: 5 variables (linear_wave) or 8 (linear_wave_mhd)"""
        out = np.zeros(9)
        self.L.apko_linear_wave_errors(self.h, _dp(out))
        return out

    def history(self):
        out = np.zeros(8)
        self.L.apko_history(self.h, _dp(out))
        return out


def recon_scalar(kind, q, dx=1.0, positivity=False):
    q = np.ascontiguousarray(q, dtype=np.float64)
    ql, qr = C.c_double(), C.c_double()
    lib().apko_recon_scalar(RECON[kind], _dp(q), dx, int(positivity), C.byref(ql), C.byref(qr))
    return ql.value, qr.value


def riemann_face(fluid, solver, ivx, wl, wr, gamma, c_h=1.0):
    wl = np.ascontiguousarray(wl, dtype=np.float64)
    wr = np.ascontiguousarray(wr, dtype=np.float64)
    out = np.zeros(9 if fluid == "glmmhd" else 5)
    lib().apko_riemann_face(FLUID[fluid], RIEMANN[solver], ivx, _dp(wl), _dp(wr), gamma, c_h,
                            _dp(out))
    return out


# ---- AMR operators: arrays are [nv][n3][n2][n1]; lo/hi inclusive coarse (i,j,k) ----
def _amr_args(ndim, ng, xmin_block, dx, fine, coarse, lo, hi):
    assert fine.dtype == np.float64 and coarse.dtype == np.float64
    assert fine.flags.c_contiguous and coarse.flags.c_contiguous
    ip = C.POINTER(C.c_int)
    nf = np.asarray(fine.shape[:0:-1], dtype=np.int32)
    nc = np.asarray(coarse.shape[:0:-1], dtype=np.int32)
    x = np.asarray(xmin_block, dtype=np.float64)
    d = np.asarray(dx, dtype=np.float64)
    l = np.asarray(lo, dtype=np.int32)
    h = np.asarray(hi, dtype=np.int32)
    keep = (nf, nc, x, d, l, h)
    return keep, (_dp(x), _dp(d), fine.shape[0], nf.ctypes.data_as(ip), _dp(fine),
                  nc.ctypes.data_as(ip), _dp(coarse), l.ctypes.data_as(ip), h.ctypes.data_as(ip))


def restrict_average(ndim, ng, el, xmin_block, dx, fine, coarse, lo, hi):
    keep, a = _amr_args(ndim, ng, xmin_block, dx, fine, coarse, lo, hi)
    getattr(lib(), 'apko_' + "restrict_average")(ndim, ng, el, *a)


def prolongate_minmod_multid(ndim, ng, xmin_block, dx, fine, coarse, lo, hi):
    keep, a = _amr_args(ndim, ng, xmin_block, dx, fine, coarse, lo, hi)
    getattr(lib(), 'apko_' + "prolongate_minmod_multid")(ndim, ng, *a)


def pressure_gradient_tag(ndim, ng, prim, ipr, threshold):
    """(tag, maxeps) of refinement::gradient::PressureGradient for one block"""
    prim = np.ascontiguousarray(prim, dtype=np.float64)
    n = np.asarray(prim.shape[:0:-1], dtype=np.int32)
    m = C.c_double()
    f = lib().apko_pressure_gradient_tag
    f.restype = C.c_int
    tag = f(ndim, ng, n.ctypes.data_as(C.POINTER(C.c_int)), _dp(prim), ipr, C.c_double(threshold),
            C.byref(m))
    return tag, m.value
