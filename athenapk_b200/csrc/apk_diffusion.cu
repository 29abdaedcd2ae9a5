// apk_diffusion.cu -- the <diffusion> block of AthenaPK on the GPU: diffusive
// fluxes with fixed coefficients that ADD to the flux arrays (diffusion/diffusion.cpp:19-53 ->
// conduction.cpp ThermalFluxIsoFixed :193-260 / ThermalFluxGeneral :265-550 for anisotropic
// conduction with saturation, viscosity.cpp MomentumDiffFluxIsoFixed :98-291, resistivity.cpp
// OhmicDiffFluxIsoFixed :89-232), ResetFluxes (hydro_driver.cpp:36-91), the two RKL2 register
// updates of the super-time-stepping integrator (RKL2StepFirst :93-126 fused with
// Update::FluxDivergence update.cpp:63-85 and the Y0 copy, RKL2StepOther :128-166) and the
// cell-dependent time-step estimate of anisotropic conduction (conduction.cpp:91-186).
// One thread = one cell and its (up to three) lower faces; a warp covers 32 consecutive x1 cells
// of one row, so every load / store is a coalesced row segment.  Operation order follows the
// reference expression by expression: bit-identical to the CPU oracle when compiled with
// -fmad=false (APK_NS = apk_strict).
#include <cstdint>

#include "apk_launch.h"
#include "apk_math.cuh"

namespace APK_NS {

namespace {

constexpr int kThreads = 128;

struct Box {
  int il, iu, jl, ju, kl, ku;
};

__device__ __forceinline__ bool decode(long t, const Box &r, int nb, int &b, int &k, int &j,
                                       int &i) {
  const int ni = r.iu - r.il + 1, nj = r.ju - r.jl + 1, nk = r.ku - r.kl + 1;
  const long per = (long)ni * nj * nk;
  if (t >= per * nb) return false;
  b = (int)(t / per);
  long rem = t - (long)b * per;
  k = r.kl + (int)(rem / ((long)ni * nj));
  rem -= (long)(k - r.kl) * ni * nj;
  j = r.jl + (int)(rem / ni);
  i = r.il + (int)(rem - (long)(j - r.jl) * ni);
  return true;
}

// cells is..ie+1 x js..je+1 x ks..ke+1 (active dimensions): every cell whose lower face lies in a
// flux loop range of the reference
__host__ __device__ inline Box face_box(const ApkGeom &g) {
  return Box{g.is, g.ie + 1, g.js, g.je + (g.ndim >= 2), g.ks, g.ke + (g.ndim >= 3)};
}

enum { IDN = 0, IV1 = 1, IV2 = 2, IV3 = 3, IPR = 4, IEN = 4, IM1 = 1, IM2 = 2, IM3 = 3, IB1 = 5,
       IB2 = 6, IB3 = 7 };

// hydro_driver.cpp:36-91
__global__ void __launch_bounds__(kThreads)
k_reset_fluxes(ApkGeom g, const apk_block_t *__restrict__ blocks) {
  int b, k, j, i;
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, face_box(g), g.nb, b, k, j, i)) return;
  const apk_block_t &B = blocks[b];
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  const bool in[3] = {j <= g.je && k <= g.ke, i <= g.ie && k <= g.ke, i <= g.ie && j <= g.je};
  for (int d = 0; d < g.ndim; ++d) {
    if (!in[d]) continue;
    for (int n = 0; n < g.nvar; ++n) B.flux[d][(long)n * g.ncell + c] = 0.0;
  }
}

struct DiffPars {
  int conduction, viscosity, resistivity;
  double chi, nu, eta, sat_prefac;
};

// diffusion.hpp:37-66
APK_DEV double lim_minmod(double A, double B) {
  if (A * B > 0) return (A > 0) ? dmin(A, B) : dmax(A, B);
  return 0.0;
}
APK_DEV double lim_mc(double A, double B) {
  return lim_minmod(2.0 * lim_minmod(A, B), fdiv(A + B, 2.0));
}
APK_DEV double lim4(double A, double B, double C, double D) {
  return lim_mc(lim_mc(A, B), lim_mc(C, D));
}

__global__ void __launch_bounds__(kThreads)
k_diff_fluxes(ApkGeom g, const apk_block_t *__restrict__ blocks, DiffPars p) {
  int b, k, j, i;
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, face_box(g), g.nb, b, k, j, i)) return;
  const apk_block_t &B = blocks[b];
  const long sj = g.n1, sk = (long)g.n1 * g.n2, nc = g.ncell;
  const long c = (long)k * sk + (long)j * sj + i;
  const double *__restrict__ W = B.prim;
  const int ndim = g.ndim;
  const long st[3] = {1, sj, sk};
  const bool in[3] = {j <= g.je && k <= g.ke, i <= g.ie && k <= g.ke, i <= g.ie && j <= g.je};
  const double dx[3] = {B.dx[0], B.dx[1], B.dx[2]};
#define PW(n, off) W[(long)(n)*nc + c + (off)]
#define TC(off) fdiv(PW(IPR, off), PW(IDN, off))

  // ---- thermal conduction --------------------------------------------------------------------
  if (p.conduction == 1) {  // ThermalFluxIsoFixed
    for (int d = 0; d < ndim; ++d) {
      if (!in[d]) continue;
      const double T_i = TC(0), T_im1 = TC(-st[d]);
      const double dTdx = fdiv(T_i - T_im1, dx[d]);
      const double denf = 0.5 * (PW(IDN, 0) + PW(IDN, -st[d]));
      B.flux[d][(long)IEN * nc + c] -= p.chi * denf * dTdx;
    }
  } else if (p.conduction == 2) {  // ThermalFluxGeneral, anisotropic, fixed coefficient
    for (int d = 0; d < ndim; ++d) {
      if (!in[d]) continue;
      const long m = -st[d];
      double dT[3] = {0.0, 0.0, 0.0};
      for (int t = 0; t < 3; ++t) {
        if (t == d) {
          dT[t] = fdiv(TC(0) - TC(m), dx[d]);
        } else if (t < 2 || ndim >= 3) {
          const long o = st[t];
          dT[t] = fdiv(lim4(TC(o) - TC(0), TC(0) - TC(-o), TC(m + o) - TC(m), TC(m) - TC(m - o)),
                       dx[t]);
        }
      }
      const double denf = 0.5 * (PW(IDN, 0) + PW(IDN, m));
      const double tdf = 0.5 * (p.chi + p.chi);
      const double Bx = 0.5 * (PW(IB1, m) + PW(IB1, 0));
      const double By = 0.5 * (PW(IB2, m) + PW(IB2, 0));
      const double Bz = (d == 2 || ndim >= 3) ? 0.5 * (PW(IB3, m) + PW(IB3, 0)) : 0.0;
      double Bmag = fsqrt(sqr(Bx) + sqr(By) + sqr(Bz));
      Bmag = dmax(Bmag, kTiny);
      const double bd = fdiv(d == 0 ? Bx : (d == 1 ? By : Bz), Bmag);
      const double bDotGradT = fdiv(Bx * dT[0] + By * dT[1] + Bz * dT[2], Bmag);
      const double flux_classic = -tdf * denf * bDotGradT * bd;
      const double flux_classic_mag = fabs(tdf * denf * bDotGradT);
      double flux_sat;
      if (flux_classic > 0.0) {
        flux_sat = p.sat_prefac * fsqrt(fdiv(PW(IPR, m), denf)) * PW(IPR, m);
      } else if (flux_classic < 0.0) {
        flux_sat = p.sat_prefac * fsqrt(fdiv(PW(IPR, 0), denf)) * PW(IPR, 0);
      } else {
        const double presf = 0.5 * (PW(IPR, 0) + PW(IPR, m));
        flux_sat = p.sat_prefac * fsqrt(fdiv(presf, denf)) * presf;
      }
      B.flux[d][(long)IEN * nc + c] += fdiv(flux_sat, flux_sat + flux_classic_mag) * flux_classic;
    }
  }

  // ---- isotropic viscosity (MomentumDiffFluxIsoFixed) -------------------------------------------
  if (p.viscosity == 1) {
    const double dx1 = dx[0], dx2 = dx[1], dx3 = dx[2];
    if (in[0]) {
      double fvx = fdiv(4.0 * (PW(IV1, 0) - PW(IV1, -1)), 3.0 * dx1);
      double fvy = fdiv(PW(IV2, 0) - PW(IV2, -1), dx1);
      double fvz = fdiv(PW(IV3, 0) - PW(IV3, -1), dx1);
      if (ndim > 1) {
        fvx -= fdiv((PW(IV2, sj) + PW(IV2, sj - 1)) - (PW(IV2, -sj) + PW(IV2, -sj - 1)), 6.0 * dx2);
        fvy += fdiv((PW(IV1, sj) + PW(IV1, sj - 1)) - (PW(IV1, -sj) + PW(IV1, -sj - 1)), 4.0 * dx2);
      }
      if (ndim > 2) {
        fvx -= fdiv((PW(IV3, sk) + PW(IV3, sk - 1)) - (PW(IV3, -sk) + PW(IV3, -sk - 1)), 6.0 * dx3);
        fvz += fdiv((PW(IV1, sk) + PW(IV1, sk - 1)) - (PW(IV1, -sk) + PW(IV1, -sk - 1)), 4.0 * dx3);
      }
      const double nud = 0.5 * p.nu * (PW(IDN, 0) + PW(IDN, -1));
      B.flux[0][(long)IM1 * nc + c] -= nud * fvx;
      B.flux[0][(long)IM2 * nc + c] -= nud * fvy;
      B.flux[0][(long)IM3 * nc + c] -= nud * fvz;
      B.flux[0][(long)IEN * nc + c] -= 0.5 * nud * ((PW(IV1, -1) + PW(IV1, 0)) * fvx +
                                                    (PW(IV2, -1) + PW(IV2, 0)) * fvy +
                                                    (PW(IV3, -1) + PW(IV3, 0)) * fvz);
    }
    if (ndim >= 2 && in[1]) {
      double fvx = fdiv(PW(IV1, 0) - PW(IV1, -sj), dx2) +
                   fdiv((PW(IV2, 1) + PW(IV2, -sj + 1)) - (PW(IV2, -1) + PW(IV2, -sj - 1)), 4.0 * dx1);
      double fvy = fdiv((PW(IV2, 0) - PW(IV2, -sj)) * 4.0, 3.0 * dx2) -
                   fdiv((PW(IV1, 1) + PW(IV1, -sj + 1)) - (PW(IV1, -1) + PW(IV1, -sj - 1)), 6.0 * dx1);
      double fvz = fdiv(PW(IV3, 0) - PW(IV3, -sj), dx2);
      if (ndim > 2) {
        fvy -= fdiv((PW(IV3, sk) + PW(IV3, sk - sj)) - (PW(IV3, -sk) + PW(IV3, -sk - sj)), 6.0 * dx3);
        fvz += fdiv((PW(IV2, sk) + PW(IV2, sk - sj)) - (PW(IV2, -sk) + PW(IV2, -sk - sj)), 4.0 * dx3);
      }
      const double nud = 0.5 * p.nu * (PW(IDN, 0) + PW(IDN, -sj));
      B.flux[1][(long)IM1 * nc + c] -= nud * fvx;
      B.flux[1][(long)IM2 * nc + c] -= nud * fvy;
      B.flux[1][(long)IM3 * nc + c] -= nud * fvz;
      B.flux[1][(long)IEN * nc + c] -= 0.5 * nud * ((PW(IV1, -sj) + PW(IV1, 0)) * fvx +
                                                    (PW(IV2, -sj) + PW(IV2, 0)) * fvy +
                                                    (PW(IV3, -sj) + PW(IV3, 0)) * fvz);
    }
    if (ndim >= 3 && in[2]) {
      const double fvx = fdiv(PW(IV1, 0) - PW(IV1, -sk), dx3) +
                         fdiv((PW(IV3, 1) + PW(IV3, -sk + 1)) - (PW(IV3, -1) + PW(IV3, -sk - 1)), 4.0 * dx1);
      const double fvy = fdiv(PW(IV2, 0) - PW(IV2, -sk), dx3) +
                         fdiv((PW(IV3, sj) + PW(IV3, -sk + sj)) - (PW(IV3, -sj) + PW(IV3, -sk - sj)), 4.0 * dx2);
      const double fvz = fdiv((PW(IV3, 0) - PW(IV3, -sk)) * 4.0, 3.0 * dx3) -
                         fdiv((PW(IV1, 1) + PW(IV1, -sk + 1)) - (PW(IV1, -1) + PW(IV1, -sk - 1)), 6.0 * dx1) -
                         fdiv((PW(IV2, sj) + PW(IV2, -sk + sj)) - (PW(IV2, -sj) + PW(IV2, -sk - sj)), 6.0 * dx2);
      const double nud = 0.5 * p.nu * (PW(IDN, 0) + PW(IDN, -sk));
      B.flux[2][(long)IM1 * nc + c] -= nud * fvx;
      B.flux[2][(long)IM2 * nc + c] -= nud * fvy;
      B.flux[2][(long)IM3 * nc + c] -= nud * fvz;
      B.flux[2][(long)IEN * nc + c] -= 0.5 * nud * ((PW(IV1, -sk) + PW(IV1, 0)) * fvx +
                                                    (PW(IV2, -sk) + PW(IV2, 0)) * fvy +
                                                    (PW(IV3, -sk) + PW(IV3, 0)) * fvz);
    }
  }

  // ---- Ohmic resistivity (OhmicDiffFluxIsoFixed) -------------------------------------------------
  if (p.resistivity == 1) {
    const double dx1 = dx[0], dx2 = dx[1], dx3 = dx[2];
    // coords.Xf<d, face != d>(idx) is the cell-centre coordinate xmin + (idx + 0.5) dx
    // (uniform_coordinates.hpp:117-157): the 2 dx spacings are differences of positions
    const double s1 = (B.xmin[0] + ((i + 1) + 0.5) * dx1) - (B.xmin[0] + ((i - 1) + 0.5) * dx1);
    const double s2 = (B.xmin[1] + ((j + 1) + 0.5) * dx2) - (B.xmin[1] + ((j - 1) + 0.5) * dx2);
    const double s3 = (B.xmin[2] + ((k + 1) + 0.5) * dx3) - (B.xmin[2] + ((k - 1) + 0.5) * dx3);
    if (in[0]) {
      const double d3B1 = ndim > 2 ? fdiv(0.5 * (PW(IB1, sk - 1) + PW(IB1, sk)) -
                                              0.5 * (PW(IB1, -sk - 1) + PW(IB1, -sk)), s3)
                                   : 0.0;
      const double d1B3 = fdiv(PW(IB3, 0) - PW(IB3, -1), dx1);
      const double j2 = d3B1 - d1B3;
      const double d1B2 = fdiv(PW(IB2, 0) - PW(IB2, -1), dx1);
      const double d2B1 = ndim > 1 ? fdiv(0.5 * (PW(IB1, sj - 1) + PW(IB1, sj)) -
                                              0.5 * (PW(IB1, -sj - 1) + PW(IB1, -sj)), s2)
                                   : 0.0;
      const double j3 = d1B2 - d2B1;
      B.flux[0][(long)IB2 * nc + c] += -p.eta * j3;
      B.flux[0][(long)IB3 * nc + c] += p.eta * j2;
      B.flux[0][(long)IEN * nc + c] +=
          0.5 * p.eta * ((PW(IB3, -1) + PW(IB3, 0)) * j2 - (PW(IB2, -1) + PW(IB2, 0)) * j3);
    }
    if (ndim >= 2 && in[1]) {
      const double d1B2 = fdiv(0.5 * (PW(IB2, -sj + 1) + PW(IB2, 1)) -
                                   0.5 * (PW(IB2, -sj - 1) + PW(IB2, -1)), s1);
      const double d2B1 = fdiv(PW(IB1, 0) - PW(IB1, -sj), dx2);
      const double j3 = d1B2 - d2B1;
      const double d2B3 = fdiv(PW(IB3, 0) - PW(IB3, -sj), dx2);
      const double d3B2 = ndim > 2 ? fdiv(0.5 * (PW(IB2, sk - sj) + PW(IB2, sk)) -
                                              0.5 * (PW(IB2, -sk - sj) + PW(IB2, -sk)), s3)
                                   : 0.0;
      const double j1 = d2B3 - d3B2;
      B.flux[1][(long)IB1 * nc + c] += p.eta * j3;
      B.flux[1][(long)IB3 * nc + c] += -p.eta * j1;
      B.flux[1][(long)IEN * nc + c] +=
          0.5 * p.eta * ((PW(IB1, -sj) + PW(IB1, 0)) * j3 - (PW(IB3, -sj) + PW(IB3, 0)) * j1);
    }
    if (ndim >= 3 && in[2]) {
      const double d2B3 = fdiv(0.5 * (PW(IB3, -sk + sj) + PW(IB3, sj)) -
                                   0.5 * (PW(IB3, -sk - sj) + PW(IB3, -sj)), s2);
      const double d3B2 = fdiv(PW(IB2, 0) - PW(IB2, -sk), dx3);
      const double j1 = d2B3 - d3B2;
      const double d3B1 = fdiv(PW(IB1, 0) - PW(IB1, -sk), dx3);
      const double d1B3 = fdiv(0.5 * (PW(IB3, -sk + 1) + PW(IB3, 1)) -
                                   0.5 * (PW(IB3, -sk - 1) + PW(IB3, -1)), s1);
      const double j2 = d3B1 - d1B3;
      B.flux[2][(long)IB1 * nc + c] += -p.eta * j2;
      B.flux[2][(long)IB2 * nc + c] += p.eta * j1;
      B.flux[2][(long)IEN * nc + c] +=
          0.5 * p.eta * ((PW(IB2, -sk) + PW(IB2, 0)) * j1 - (PW(IB1, -sk) + PW(IB1, 0)) * j2);
    }
  }
#undef TC
#undef PW
}

// FluxDivHelper (update.hpp:42-56)
__device__ __forceinline__ double flux_div(const ApkGeom &g, const apk_block_t &B, int l, long c,
                                           double a1, double a2, double a3, double vol) {
  const long o = (long)l * g.ncell + c;
  double du = (a1 * B.flux[0][o + 1] - a1 * B.flux[0][o]);
  if (g.ndim >= 2) du += (a2 * B.flux[1][o + g.n1] - a2 * B.flux[1][o]);
  if (g.ndim == 3) du += (a3 * B.flux[2][o + (long)g.n1 * g.n2] - a3 * B.flux[2][o]);
  return fdiv(-du, vol);
}

// Y0 <- base (hydro_driver.cpp:196-208), MY0 = FluxDivergence(base) (:255), RKL2StepFirst (:93-126):
// base = Y0 + mu_tilde_1 tau MY0, Yjm2 = Y0.  mt = mu_tilde_1 * tau (host product, same rounding)
__global__ void __launch_bounds__(kThreads)
k_rkl2_first(ApkGeom g, const apk_block_t *__restrict__ blocks, double *const *__restrict__ my0,
             double *const *__restrict__ yjm2, double mt) {
  int b, k, j, i;
  const Box r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, r, g.nb, b, k, j, i)) return;
  const apk_block_t &B = blocks[b];
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  const double a1 = B.dx[1] * B.dx[2], a2 = B.dx[0] * B.dx[2], a3 = B.dx[0] * B.dx[1];
  const double vol = B.dx[0] * B.dx[1] * B.dx[2];
  double *__restrict__ M = my0[b], *__restrict__ Y2 = yjm2[b];
  for (int l = 0; l < g.nvar; ++l) {
    const long o = (long)l * g.ncell + c;
    const double y0 = B.cons[o];
    const double m = flux_div(g, B, l, c, a1, a2, a3, vol);
    B.cons_u1[o] = y0;
    M[o] = m;
    B.cons[o] = y0 + mt * m;
    Y2[o] = y0;
  }
}

// RKL2StepOther (hydro_driver.cpp:128-166); c0 = 1 - mu_j - nu_j, mtt = mu_tilde_j tau,
// gtt = gamma_tilde_j tau (host products in the reference's association)
__global__ void __launch_bounds__(kThreads)
k_rkl2_other(ApkGeom g, const apk_block_t *__restrict__ blocks, double *const *__restrict__ my0,
             double *const *__restrict__ yjm2, double mu_j, double nu_j, double c0, double mtt,
             double gtt) {
  int b, k, j, i;
  const Box r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, r, g.nb, b, k, j, i)) return;
  const apk_block_t &B = blocks[b];
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  const double a1 = B.dx[1] * B.dx[2], a2 = B.dx[0] * B.dx[2], a3 = B.dx[0] * B.dx[1];
  const double vol = B.dx[0] * B.dx[1] * B.dx[2];
  const double *__restrict__ M = my0[b];
  double *__restrict__ Y2 = yjm2[b];
  for (int l = 0; l < g.nvar; ++l) {
    const long o = (long)l * g.ncell + c;
    const double MYjm1 = flux_div(g, B, l, c, a1, a2, a3, vol);
    const double yjm1 = B.cons[o];
    const double Yj = mu_j * yjm1 + nu_j * Y2[o] + c0 * B.cons_u1[o] + mtt * MYjm1 + gtt * M[o];
    Y2[o] = yjm1;
    B.cons[o] = Yj;
  }
}

// EstimateConductionTimestep, general branch for anisotropic conduction with a fixed coefficient
// (conduction.cpp:91-186): min over interior cells of dx^2 / (chi |B_d| / |B| cos(theta) + TINY)
__global__ void __launch_bounds__(kThreads)
k_min_dt_cond(ApkGeom g, const apk_block_t *__restrict__ blocks, double chi, double sat_prefac,
              unsigned long long *out) {
  int b, k, j, i;
  const Box r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  double m = DBL_MAX;
  if (decode((long)blockIdx.x * blockDim.x + threadIdx.x, r, g.nb, b, k, j, i)) {
    const apk_block_t &B = blocks[b];
    const long sj = g.n1, sk = (long)g.n1 * g.n2, nc = g.ncell;
    const long c = (long)k * sk + (long)j * sj + i;
    const double *__restrict__ W = B.prim;
    const int ndim = g.ndim;
#define PW(n, off) W[(long)(n)*nc + c + (off)]
#define TC(off) fdiv(PW(IPR, off), PW(IDN, off))
    const double rho = PW(IDN, 0), p = PW(IPR, 0);
    const double dTdx = fdiv(0.5 * (TC(1) - TC(-1)), B.dx[0]);
    const double dTdy = ndim >= 2 ? fdiv(0.5 * (TC(sj) - TC(-sj)), B.dx[1]) : 0.0;
    const double dTdz = ndim >= 3 ? fdiv(0.5 * (TC(sk) - TC(-sk)), B.dx[2]) : 0.0;
    const double gradTmag = fsqrt(sqr(dTdx) + sqr(dTdy) + sqr(dTdz));
    const double Bx = PW(IB1, 0), By = PW(IB2, 0), Bz = PW(IB3, 0);
    const double Bmag = fsqrt(sqr(Bx) + sqr(By) + sqr(Bz));
    if (gradTmag != 0.0 && Bmag != 0.0) {
      const double flux_sat = sat_prefac * fsqrt(fdiv(p, rho)) * p;
      const double flux_classic = chi * rho * gradTmag;
      if (!(fdiv(flux_classic, flux_sat) > 100.)) {
        const double costheta = fdiv(fabs(Bx * dTdx + By * dTdy + Bz * dTdz), Bmag * gradTmag);
        m = fmin(m, fdiv(sqr(B.dx[0]), fdiv(chi * fabs(Bx), Bmag) * costheta + kTiny));
        if (ndim >= 2) m = fmin(m, fdiv(sqr(B.dx[1]), fdiv(chi * fabs(By), Bmag) * costheta + kTiny));
        if (ndim >= 3) m = fmin(m, fdiv(sqr(B.dx[2]), fdiv(chi * fabs(Bz), Bmag) * costheta + kTiny));
      }
    }
#undef TC
#undef PW
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m < DBL_MAX)
    atomicMin(out, (unsigned long long)__double_as_longlong(m));
}

__global__ void k_set_u64(unsigned long long *p, unsigned long long v) { *p = v; }

inline unsigned grid_for(long n) { return (unsigned)((n + kThreads - 1) / kThreads); }
inline long box_cells(const Box &r, int nb) {
  return (long)(r.iu - r.il + 1) * (r.ju - r.jl + 1) * (r.ku - r.kl + 1) * nb;
}

}  // namespace

int launch_reset_fluxes(const apk_pack_desc_t &d, const apk_block_t *blocks, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  k_reset_fluxes<<<grid_for(box_cells(face_box(g), g.nb)), kThreads, 0, s>>>(g, blocks);
  apk_common::count_launch();
  return 0;
}

int launch_diff_fluxes(const apk_pack_desc_t &d, const apk_block_t *blocks,
                       const apk_diffusion_t &df, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  if (df.conduction == APK_CONDUCTION_ANISOTROPIC && g.ndim < 2) {
    apk_common::set_error("apk_calc_diff_fluxes: anisotropic conduction needs a 2-D or 3-D mesh");
    return 1;
  }
  if ((df.conduction == APK_CONDUCTION_ANISOTROPIC || df.resistivity) &&
      d.fluid != APK_FLUID_GLMMHD) {
    apk_common::set_error("apk_calc_diff_fluxes: anisotropic conduction / resistivity need fluid = glmmhd");
    return 1;
  }
  const DiffPars p = {df.conduction, df.viscosity, df.resistivity, df.thermal_diff_coeff,
                      df.mom_diff_coeff, df.ohm_diff_coeff, df.conduction_sat_prefac};
  k_diff_fluxes<<<grid_for(box_cells(face_box(g), g.nb)), kThreads, 0, s>>>(g, blocks, p);
  apk_common::count_launch();
  return 0;
}

int launch_rkl2_first(const apk_pack_desc_t &d, const apk_block_t *blocks, double *const *my0,
                      double *const *yjm2, double mt, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  const Box r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  k_rkl2_first<<<grid_for(box_cells(r, g.nb)), kThreads, 0, s>>>(g, blocks, my0, yjm2, mt);
  apk_common::count_launch();
  return 0;
}

int launch_rkl2_other(const apk_pack_desc_t &d, const apk_block_t *blocks, double *const *my0,
                      double *const *yjm2, double mu_j, double nu_j, double c0, double mtt,
                      double gtt, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  const Box r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  k_rkl2_other<<<grid_for(box_cells(r, g.nb)), kThreads, 0, s>>>(g, blocks, my0, yjm2, mu_j, nu_j,
                                                                c0, mtt, gtt);
  apk_common::count_launch();
  return 0;
}

int launch_min_dt_conduction(const apk_pack_desc_t &d, const apk_block_t *blocks,
                             const apk_diffusion_t &df, double *d_out, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  const Box r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  unsigned long long *out = reinterpret_cast<unsigned long long *>(d_out);
  const double big = DBL_MAX;
  unsigned long long bits;
  memcpy(&bits, &big, sizeof(bits));
  k_set_u64<<<1, 1, 0, s>>>(out, bits);
  k_min_dt_cond<<<grid_for(box_cells(r, g.nb)), kThreads, 0, s>>>(
      g, blocks, df.thermal_diff_coeff, df.conduction_sat_prefac, out);
  apk_common::count_launch(2);
  return 0;
}

}  // namespace APK_NS
