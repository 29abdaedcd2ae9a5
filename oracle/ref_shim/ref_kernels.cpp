// ref_kernels.cpp -- ORACLE BUILD SUPPORT (test infrastructure, NOT product code).
//
// Compiles the reference's OWN hot-path headers, unmodified, from /root/reference (include
// path), against the serial Parthenon stand-in of apk_ref_stub.hpp, and exposes them through
// a flat C interface so tests can pin oracle/apk_oracle.c against the real reference code:
//   src/recon/{plm,ppm,wenoz,weno3,limo3}_simple.hpp  scalar reconstruction functions
//   src/hydro/rsolvers/*.hpp                           Riemann<fluid,solver>::Solve
//   src/eos/adiabatic_{hydro,glmmhd}.hpp               ConsToPrim, sound/fast speeds
// Output goes to oracle/_ref/libapk_ref.so (git-ignored). No reference source is copied.
#include "apk_ref_stub.hpp"

extern "C" int apk_ref_require_failed;
int apk_ref_require_failed = 0;

#include "main.hpp" // reference src/main.hpp (enums IDN.., Fluid, Reconstruction, ...)


#include "eos/adiabatic_glmmhd.hpp"
#include "eos/adiabatic_hydro.hpp"
#include "recon/dc_simple.hpp"
#include "recon/limo3_simple.hpp"
#include "recon/plm_simple.hpp"
#include "recon/ppm_simple.hpp"
#include "recon/weno3_simple.hpp"
#include "recon/wenoz_simple.hpp"
#include "hydro/rsolvers/rsolvers.hpp"
#include "hydro/prolongation/custom_ops.hpp"  // + Parthenon's own prolong_restrict/pr_ops.hpp

// the pack-level drivers live in .cpp files we do not build; give the vtable a home
void AdiabaticHydroEOS::ConservedToPrimitive(parthenon::MeshData<Real> *) const {}
void AdiabaticGLMMHDEOS::ConservedToPrimitive(parthenon::MeshData<Real> *) const {}

// Full Reconstruct<recon, DIR> wrapper on a [nv][nk][nj][ni] prim array (tests the index
// conventions ql(n,i+1)/qr(n,i) in X1 and ql(n,i)/qr(n,i) in X2/X3).
// ql/qr are [nv][ni] outputs for the pencil (k, j, il..iu).
template <Reconstruction R>
static void recon_pencil_t(int dir, const double *prim, int nv, int nk, int nj, int ni, int k,
                           int j, int il, int iu, const double *dx3, double *ql, double *qr) {
  parthenon::VariablePack<Real> q(nv, nk, nj, ni);
  for (int d = 0; d < 3; ++d) q.coords.dx[d] = dx3[d];
  for (int n = 0; n < nv; ++n)
    for (int kk = 0; kk < nk; ++kk)
      for (int jj = 0; jj < nj; ++jj)
        for (int ii = 0; ii < ni; ++ii)
          q(n, kk, jj, ii) = prim[(((size_t)n * nk + kk) * nj + jj) * ni + ii];
  parthenon::ScratchPad2D<Real> wl(nv, ni + 1), wr(nv, ni + 1);
  for (int n = 0; n < nv; ++n)
    for (int i = 0; i < ni + 1; ++i) wl(n, i) = wr(n, i) = NAN;
  parthenon::team_mbr_t m;
  if (dir == 1) Reconstruct<R, parthenon::X1DIR>(m, k, j, il, iu, q, wl, wr);
  if (dir == 2) Reconstruct<R, parthenon::X2DIR>(m, k, j, il, iu, q, wl, wr);
  if (dir == 3) Reconstruct<R, parthenon::X3DIR>(m, k, j, il, iu, q, wl, wr);
  for (int n = 0; n < nv; ++n)
    for (int i = 0; i < ni; ++i) {
      ql[n * ni + i] = wl(n, i);
      qr[n * ni + i] = wr(n, i);
    }
}

extern "C" {

// kind: 0 dc, 1 plm, 2 ppm, 3 wenoz, 4 weno3, 5 limo3 ; q = q_{i-2..i+2}
void apk_ref_recon_scalar(int kind, const double *q, double dx, int positivity, double *ql_ip1,
                          double *qr_i) {
  switch (kind) {
  case 0: *ql_ip1 = *qr_i = q[2]; break;
  case 1: PLM(q[1], q[2], q[3], *ql_ip1, *qr_i); break;
  case 2: PPM(q[0], q[1], q[2], q[3], q[4], *ql_ip1, *qr_i); break;
  case 3: WENOZ(q[0], q[1], q[2], q[3], q[4], *ql_ip1, *qr_i); break;
  case 4: WENO3(q[1], q[2], q[3], *ql_ip1, *qr_i, dx * dx); break;
  case 5: LimO3(q[1], q[2], q[3], *ql_ip1, *qr_i, dx, positivity != 0); break;
  default: *ql_ip1 = *qr_i = NAN;
  }
}

void apk_ref_recon_pencil(int kind, int dir, const double *prim, int nv, int nk, int nj, int ni,
                          int k, int j, int il, int iu, const double *dx3, double *ql,
                          double *qr) {
#define GO(R) recon_pencil_t<Reconstruction::R>(dir, prim, nv, nk, nj, ni, k, j, il, iu, dx3, ql, qr)
  switch (kind) {
  case 0: GO(dc); break;
  case 1: GO(plm); break;
  case 2: GO(ppm); break;
  case 3: GO(wenoz); break;
  case 4: GO(weno3); break;
  case 5: GO(limo3); break;
  }
#undef GO
}

// fluid: 0 euler, 1 glmmhd ; solver: 0 none, 1 hlle, 2 llf, 3 hllc, 4 hlld ; ivx in {1,2,3}
// wl/wr/flux are full state vectors in the unpermuted variable layout.
void apk_ref_riemann_face(int fluid, int solver, int ivx, const double *wl_in,
                          const double *wr_in, double gamma, double c_h, double *flux) {
  const int nv = fluid ? 9 : 5;
  parthenon::team_mbr_t m;
  parthenon::ScratchPad2D<Real> wl(nv, 2), wr(nv, 2);
  for (int n = 0; n < nv; ++n) {
    wl(n, 1) = wl_in[n];
    wr(n, 1) = wr_in[n];
    wl(n, 0) = wr(n, 0) = NAN;
  }
  // (k,j,i) layout 2x2x2 so that i-1/j-1/k-1 exist for the cell-wise LLF solver
  VariableFluxPack<Real> cons(nv, 2, 2, 2);
  VariablePack<Real> prim(nv, 2, 2, 2);
  for (int n = 0; n < nv; ++n) {
    prim(n, 1, 1, 1) = wr_in[n];
    prim(n, 1, 1, 0) = prim(n, 1, 0, 1) = prim(n, 0, 1, 1) = NAN;
    if (ivx == 1) prim(n, 1, 1, 0) = wl_in[n];
    if (ivx == 2) prim(n, 1, 0, 1) = wl_in[n];
    if (ivx == 3) prim(n, 0, 1, 1) = wl_in[n];
  }
  const Real inf = std::numeric_limits<Real>::infinity();
  if (fluid == 0) {
    AdiabaticHydroEOS eos(-1.0, -1.0, -1.0, inf, inf, gamma);
    switch (solver) {
    case 0: Riemann<Fluid::euler, RiemannSolver::none>::Solve(m, 1, 1, 1, 1, ivx, wl, wr, cons, eos, c_h); break;
    case 1: Riemann<Fluid::euler, RiemannSolver::hlle>::Solve(m, 1, 1, 1, 1, ivx, wl, wr, cons, eos, c_h); break;
    case 2: Riemann<Fluid::euler, RiemannSolver::llf>::Solve(eos, 1, 1, 1, ivx, prim, cons, c_h); break;
    case 3: Riemann<Fluid::euler, RiemannSolver::hllc>::Solve(m, 1, 1, 1, 1, ivx, wl, wr, cons, eos, c_h); break;
    default: for (int n = 0; n < nv; ++n) cons.flux(ivx, n, 1, 1, 1) = NAN;
    }
  } else {
    AdiabaticGLMMHDEOS eos(-1.0, -1.0, -1.0, inf, inf, gamma);
    switch (solver) {
    case 0: Riemann<Fluid::glmmhd, RiemannSolver::none>::Solve(m, 1, 1, 1, 1, ivx, wl, wr, cons, eos, c_h); break;
    case 1: Riemann<Fluid::glmmhd, RiemannSolver::hlle>::Solve(m, 1, 1, 1, 1, ivx, wl, wr, cons, eos, c_h); break;
    case 2: Riemann<Fluid::glmmhd, RiemannSolver::llf>::Solve(eos, 1, 1, 1, ivx, prim, cons, c_h); break;
    case 4: Riemann<Fluid::glmmhd, RiemannSolver::hlld>::Solve(m, 1, 1, 1, 1, ivx, wl, wr, cons, eos, c_h); break;
    default: for (int n = 0; n < nv; ++n) cons.flux(ivx, n, 1, 1, 1) = NAN;
    }
  }
  for (int n = 0; n < nv; ++n) flux[n] = cons.flux(ivx, n, 1, 1, 1);
}

// one-cell ConsToPrim; u is updated in place (floors/ceilings), w receives the primitives.
// returns 1 if a PARTHENON_REQUIRE (negative density/pressure without floor) fired.
int apk_ref_c2p_cell(int fluid, int nscalars, double *u, double *w, double gamma, double dfloor,
                     double pfloor, double efloor, double vceil, double eceil) {
  const int nh = fluid ? 9 : 5;
  const int nv = nh + nscalars;
  VariablePack<Real> cons(nv, 1, 1, 1), prim(nv, 1, 1, 1);
  for (int n = 0; n < nv; ++n) cons(n, 0, 0, 0) = u[n];
  apk_ref_require_failed = 0;
  if (fluid == 0) {
    AdiabaticHydroEOS eos(pfloor, dfloor, efloor, vceil, eceil, gamma);
    eos.ConsToPrim(cons, prim, nh, nscalars, 0, 0, 0);
  } else {
    AdiabaticGLMMHDEOS eos(pfloor, dfloor, efloor, vceil, eceil, gamma);
    eos.ConsToPrim(cons, prim, nh, nscalars, 0, 0, 0);
  }
  for (int n = 0; n < nv; ++n) {
    u[n] = cons(n, 0, 0, 0);
    w[n] = prim(n, 0, 0, 0);
  }
  return apk_ref_require_failed;
}

double apk_ref_fast_speed(double gamma, double d, double p, double bx, double by, double bz) {
  const Real inf = std::numeric_limits<Real>::infinity();
  AdiabaticGLMMHDEOS eos(-1.0, -1.0, -1.0, inf, inf, gamma);
  return eos.FastMagnetosonicSpeed(d, p, bx, by, bz);
}
} // extern "C"


// ---- AMR operators: the reference's own structs on caller-owned arrays --------
// fine: [nv][n3][n2][n1], coarse: [nv][c3][c2][c1]; coarse index range [lo, hi] per dimension;
// fine/coarse interior start indices as in Parthenon (ng in active dimensions, 0 otherwise).
namespace {
struct AmrGeom {
  parthenon::Coordinates_t coords, coarse_coords;
  IndexRange cib, cjb, ckb, ib, jb, kb;
};
AmrGeom amr_geom(int ndim, int ng, const double *xmin_block, const double *dx) {
  int istart[3] = {ng, ndim >= 2 ? ng : 0, ndim >= 3 ? ng : 0};
  AmrGeom g;
  g.coords = parthenon::Coordinates_t(xmin_block, dx, istart);
  g.coarse_coords = parthenon::Coordinates_t(g.coords, 2);
  g.ib.s = g.cib.s = istart[0];
  g.jb.s = g.cjb.s = istart[1];
  g.kb.s = g.ckb.s = istart[2];
  return g;
}
template <class Op, parthenon::TE el, class... A>
void amr_dispatch(int ndim, A... a) {
  if (ndim == 1) Op::template Do<1, el>(a...);
  else if (ndim == 2) Op::template Do<2, el>(a...);
  else Op::template Do<3, el>(a...);
}
}  // namespace

extern "C" {
// el: 0 = cell centred (RestrictAverage on cons), 1/2/3 = x1/x2/x3 face flux (flux correction)
void apk_ref_restrict_average(int ndim, int ng, int el, const double *xmin_block, const double *dx,
                              int nv, const int *nfine, double *fine, const int *ncoarse,
                              double *coarse, const int *lo, const int *hi) {
  AmrGeom g = amr_geom(ndim, ng, xmin_block, dx);
  ParArrayND<Real, parthenon::VariableState> f(fine, nv, nfine[2], nfine[1], nfine[0]);
  ParArrayND<Real, parthenon::VariableState> c(coarse, nv, ncoarse[2], ncoarse[1], ncoarse[0]);
  using Op = parthenon::refinement_ops::RestrictAverage;
  for (int n = 0; n < nv; ++n)
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
#define APK_REF_CALL(E)                                                                        \
  amr_dispatch<Op, E>(ndim, 0, 0, n, k, j, i, g.ckb, g.cjb, g.cib, g.kb, g.jb, g.ib, g.coords, \
                      g.coarse_coords, &c, &f)
          if (el == 0) APK_REF_CALL(parthenon::TE::CC);
          else if (el == 1) APK_REF_CALL(parthenon::TE::F1);
          else if (el == 2) APK_REF_CALL(parthenon::TE::F2);
          else APK_REF_CALL(parthenon::TE::F3);
#undef APK_REF_CALL
        }
}

void apk_ref_prolongate_minmod_multid(int ndim, int ng, const double *xmin_block,
                                      const double *dx, int nv, const int *nfine, double *fine,
                                      const int *ncoarse, double *coarse, const int *lo,
                                      const int *hi) {
  AmrGeom g = amr_geom(ndim, ng, xmin_block, dx);
  ParArrayND<Real, parthenon::VariableState> f(fine, nv, nfine[2], nfine[1], nfine[0]);
  ParArrayND<Real, parthenon::VariableState> c(coarse, nv, ncoarse[2], ncoarse[1], ncoarse[0]);
  using Op = Hydro::refinement_ops::ProlongateCellMinModMultiD;
  for (int n = 0; n < nv; ++n)
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i)
          amr_dispatch<Op, parthenon::TE::CC>(ndim, 0, 0, n, k, j, i, g.ckb, g.cjb, g.cib, g.kb,
                                              g.jb, g.ib, g.coords, g.coarse_coords, &c, &f);
}
}
