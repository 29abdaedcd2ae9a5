"""ctypes binding of oracle/_ref/libapk_ref.so: the reference's own hot-path headers
(src/recon, src/hydro/rsolvers, src/eos) compiled unmodified against a serial Parthenon
stand-in (oracle/ref_shim).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle_lib import FLUID, ORACLE_DIR, RECON, RIEMANN

_lib = None


def available():
    return lib() is not None


def lib():
    global _lib
    if _lib is not None:
        return _lib or None
    path = os.path.join(ORACLE_DIR, "_ref", "libapk_ref.so")
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])
    if not os.path.exists(path):
        _lib = False
        return None
    L = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    L.apk_ref_recon_scalar.argtypes = [C.c_int, dp, C.c_double, C.c_int, dp, dp]
    L.apk_ref_recon_pencil.argtypes = [C.c_int, C.c_int, dp] + [C.c_int] * 8 + [dp, dp, dp]
    L.apk_ref_riemann_face.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_double,
                                       C.c_double, dp]
    L.apk_ref_c2p_cell.argtypes = [C.c_int, C.c_int, dp, dp] + [C.c_double] * 6
    L.apk_ref_c2p_cell.restype = C.c_int
    L.apk_ref_fast_speed.argtypes = [C.c_double] * 6
    L.apk_ref_fast_speed.restype = C.c_double
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def recon_scalar(kind, q, dx=1.0, positivity=False):
    q = np.ascontiguousarray(q, dtype=np.float64)
    ql, qr = C.c_double(), C.c_double()
    lib().apk_ref_recon_scalar(RECON[kind], _dp(q), dx, int(positivity), C.byref(ql),
                               C.byref(qr))
    return ql.value, qr.value


def recon_pencil(kind, direction, prim, k, j, il, iu, dx3=(1.0, 1.0, 1.0)):
    prim = np.ascontiguousarray(prim, dtype=np.float64)
    nv, nk, nj, ni = prim.shape
    ql = np.zeros((nv, ni))
    qr = np.zeros((nv, ni))
    dx = np.asarray(dx3, dtype=np.float64)
    lib().apk_ref_recon_pencil(RECON[kind], direction, _dp(prim), nv, nk, nj, ni, k, j, il, iu,
                               _dp(dx), _dp(ql), _dp(qr))
    return ql, qr


def riemann_face(fluid, solver, ivx, wl, wr, gamma, c_h=1.0):
    wl = np.ascontiguousarray(wl, dtype=np.float64)
    wr = np.ascontiguousarray(wr, dtype=np.float64)
    out = np.zeros(9 if fluid == "glmmhd" else 5)
    lib().apk_ref_riemann_face(FLUID[fluid], RIEMANN[solver], ivx, _dp(wl), _dp(wr), gamma,
                               c_h, _dp(out))
    return out


def c2p_cell(fluid, u, gamma, nscalars=0, dfloor=-1.0, pfloor=-1.0, efloor=-1.0,
             vceil=float("inf"), eceil=float("inf")):
    u = np.array(u, dtype=np.float64)
    w = np.zeros_like(u)
    err = lib().apk_ref_c2p_cell(FLUID[fluid], nscalars, _dp(u), _dp(w), gamma, dfloor, pfloor,
                                 efloor, vceil, eceil)
    return u, w, err


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
    getattr(lib(), 'apk_ref_' + "restrict_average")(ndim, ng, el, *a)


def prolongate_minmod_multid(ndim, ng, xmin_block, dx, fine, coarse, lo, hi):
    keep, a = _amr_args(ndim, ng, xmin_block, dx, fine, coarse, lo, hi)
    getattr(lib(), 'apk_ref_' + "prolongate_minmod_multid")(ndim, ng, *a)
