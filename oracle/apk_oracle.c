/* apk_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See apk_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no FMA contraction, like the reference's
 * Kokkos-Serial build with g++ -O3 and no -march).
 * All file:line cites are relative to /root/reference.
 */
#include "apk_oracle.h"

#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* external/parthenon/src/config.hpp.in:86-89 */
#define TINY_NUMBER 1.0e-20
#define SQR(x) ((x) * (x))
#define SIGN(x) (((x) < 0.0) ? -1.0 : 1.0)
/* std::min / std::max semantics (first argument wins ties) */
static inline double smin(double a, double b) { return (b < a) ? b : a; }
static inline double smax(double a, double b) { return (a < b) ? b : a; }

/* variable indices, src/main.hpp:19-33 */
enum { IDN = 0, IM1 = 1, IM2 = 2, IM3 = 3, IEN = 4, NHYDRO = 5, IB1 = 5, IB2 = 6, IB3 = 7, IPS = 8 };
enum { IV1 = 1, IV2 = 2, IV3 = 3, IPR = 4 };
#define NGLMMHD 9
#define MAXVAR 16

/* ===================================================================================== */
/* Reconstruction scalars                                                                */
/* ===================================================================================== */

/* src/recon/plm_simple.hpp:20-37 */
static void rc_plm(double qm1, double q0, double qp1, double *ql_ip1, double *qr_i) {
  double dql = q0 - qm1;
  double dqr = qp1 - q0;
  double dq2 = dql * dqr;
  double dqm = 0.0;
  if (dq2 > 0.0) dqm = dq2 / (dql + dqr);
  *ql_ip1 = q0 + dqm;
  *qr_i = q0 - dqm;
}

/* src/recon/ppm_simple.hpp:38-162 */
static void rc_ppm(double qm2, double qm1, double q0, double qp1, double qp2, double *ql_ip1,
                   double *qr_i) {
  const double C2 = 1.25;
  /* step 1: interface averages (CW 1.6) */
  double qa = q0 - qm1;
  double qb = qp1 - q0;
  const double dd_m = 0.5 * qa + 0.5 * (qm1 - qm2);
  const double dd_0 = 0.5 * qb + 0.5 * qa;
  const double dd_p = 0.5 * (qp2 - qp1) + 0.5 * qb;
  double fm = 0.5 * (qm1 + q0) + (dd_m - dd_0) / 6.0;  /* face i-1/2 */
  double fp = 0.5 * (q0 + qp1) + (dd_0 - dd_p) / 6.0;  /* face i+1/2 */

  /* step 2a: second-derivative limiter at the interfaces (CD 4.3.1) */
  const double d2m = qm2 + q0 - 2.0 * qm1;
  const double d2c = qm1 + qp1 - 2.0 * q0;
  const double d2p = q0 + qp2 - 2.0 * qp1;

  double ta = fm - qm1;
  double tb = q0 - fm;
  qa = 3.0 * (qm1 + q0 - 2.0 * fm);
  qb = d2m;
  double qc = d2c;
  double qd = 0.0;
  if (SIGN(qa) == SIGN(qb) && SIGN(qa) == SIGN(qc)) {
    qd = SIGN(qa) * smin(C2 * fabs(qb), smin(C2 * fabs(qc), fabs(qa)));
  }
  double fm_lim = 0.5 * (qm1 + q0) - qd / 6.0;
  if (ta * tb < 0.0) fm = fm_lim;

  ta = fp - q0;
  tb = qp1 - fp;
  qa = 3.0 * (q0 + qp1 - 2.0 * fp);
  qb = d2c;
  qc = d2p;
  qd = 0.0;
  if (SIGN(qa) == SIGN(qb) && SIGN(qa) == SIGN(qc)) {
    qd = SIGN(qa) * smin(C2 * fabs(qb), smin(C2 * fabs(qc), fabs(qa)));
  }
  double fp_lim = 0.5 * (q0 + qp1) - qd / 6.0;
  if (ta * tb < 0.0) fp = fp_lim;

  const double d2f = 6.0 * (fm + fp - 2.0 * q0);

  double r_i = fm;
  double l_ip1 = fp;

  /* step 3 */
  const double dq_minus = q0 - r_i;
  const double dq_plus = l_ip1 - q0;

  /* step 4: CS limiters on the parabola */
  ta = dq_minus * dq_plus;
  tb = (qp1 - q0) * (q0 - qm1);

  qa = d2m;
  qb = d2c;
  qc = d2p;
  qd = d2f;
  double qe = 0.0;
  if (SIGN(qa) == SIGN(qb) && SIGN(qa) == SIGN(qc) && SIGN(qa) == SIGN(qd)) {
    qe = SIGN(qd) * smin(smin(C2 * fabs(qa), C2 * fabs(qb)), smin(C2 * fabs(qc), fabs(qd)));
  }
  qa = smax(fabs(qm1), fabs(qm2));
  qb = smax(smax(fabs(q0), fabs(qp1)), fabs(qp2));

  double rho = 0.0;
  if (fabs(qd) > (1.0e-12) * smax(qa, qb)) rho = qe / qd;

  double tmp_m = q0 - rho * dq_minus;
  double tmp_p = q0 + rho * dq_plus;
  double tmp2_m = q0 - 2.0 * dq_plus;
  double tmp2_p = q0 + 2.0 * dq_minus;

  if (ta <= 0.0 || tb <= 0.0) {
    if (rho <= (1.0 - (1.0e-12))) {
      r_i = tmp_m;
      l_ip1 = tmp_p;
    }
  } else {
    if (fabs(dq_minus) >= 2.0 * fabs(dq_plus)) r_i = tmp2_m;
    if (fabs(dq_plus) >= 2.0 * fabs(dq_minus)) l_ip1 = tmp2_p;
  }
  *qr_i = r_i;
  *ql_ip1 = l_ip1;
}

/* src/recon/wenoz_simple.hpp:27-81 */
static void rc_wenoz(double qm2, double qm1, double q0, double qp1, double qp2, double *ql_ip1,
                     double *qr_i) {
  const double c0 = 13. / 12., c1 = 0.25;
  double beta[3];
  beta[0] = c0 * SQR(qm2 + q0 - 2.0 * qm1) + c1 * SQR(qm2 + 3.0 * q0 - 4.0 * qm1);
  beta[1] = c0 * SQR(qm1 + qp1 - 2.0 * q0) + c1 * SQR(qm1 - qp1);
  beta[2] = c0 * SQR(qp2 + q0 - 2.0 * qp1) + c1 * SQR(qp2 + 3.0 * q0 - 4.0 * qp1);
  const double epsL = 1.0e-42;
  const double tau5 = fabs(beta[0] - beta[2]);
  double ind[3];
  ind[0] = tau5 / (beta[0] + epsL);
  ind[1] = tau5 / (beta[1] + epsL);
  ind[2] = tau5 / (beta[2] + epsL);

  double f[3], al[3], asum;
  f[0] = (2.0 * qm2 - 7.0 * qm1 + 11.0 * q0);
  f[1] = (-1.0 * qm1 + 5.0 * q0 + 2.0 * qp1);
  f[2] = (2.0 * q0 + 5.0 * qp1 - qp2);
  al[0] = 0.1 * (1.0 + SQR(ind[0]));
  al[1] = 0.6 * (1.0 + SQR(ind[1]));
  al[2] = 0.3 * (1.0 + SQR(ind[2]));
  asum = 6.0 * (al[0] + al[1] + al[2]);
  *ql_ip1 = (f[0] * al[0] + f[1] * al[1] + f[2] * al[2]) / asum;

  f[0] = (2.0 * qp2 - 7.0 * qp1 + 11.0 * q0);
  f[1] = (-1.0 * qp1 + 5.0 * q0 + 2.0 * qm1);
  f[2] = (2.0 * q0 + 5.0 * qm1 - qm2);
  al[0] = 0.1 * (1.0 + SQR(ind[2]));
  al[1] = 0.6 * (1.0 + SQR(ind[1]));
  al[2] = 0.3 * (1.0 + SQR(ind[0]));
  asum = 6.0 * (al[0] + al[1] + al[2]);
  *qr_i = (f[0] * al[0] + f[1] * al[1] + f[2] * al[2]) / asum;
}

/* src/recon/weno3_simple.hpp:26-63; dx2 = dx^2 of the sweep direction (:82-83) */
static void rc_weno3(double qm1, double q0, double qp1, double dx2, double *ql_ip1,
                     double *qr_i) {
  double beta[2], ind[2], f[2], al[2], asum;
  beta[0] = SQR(qp1 - q0);
  beta[1] = SQR(q0 - qm1);
  const double tau = SQR(qp1 - 2.0 * q0 + qm1);
  ind[0] = tau / (beta[0] + dx2);
  ind[1] = tau / (beta[1] + dx2);
  f[0] = q0 + qp1;
  f[1] = -qm1 + 3.0 * q0;
  al[0] = (1.0 + ind[0]) * 2.0 / 3.0;
  al[1] = (1.0 + ind[1]) / 3.0;
  asum = 2.0 * (al[0] + al[1]);
  *ql_ip1 = (al[0] * f[0] + al[1] * f[1]) / asum;
  f[0] = q0 + qm1;
  f[1] = -qp1 + 3.0 * q0;
  al[0] = (1.0 + ind[1]) * 2.0 / 3.0;
  al[1] = (1.0 + ind[0]) / 3.0;
  asum = 2.0 * (al[0] + al[1]);
  *qr_i = (al[0] * f[0] + al[1] * f[1]) / asum;
}

/* src/hydro/diffusion/diffusion.hpp:37-47 */
static double lim_minmod(double A, double B) {
  if (A * B > 0) {
    if (A > 0) return smin(A, B);
    return smax(A, B);
  }
  return 0.0;
}

/* src/recon/limo3_simple.hpp:27-55 */
static double limo3_limiter(double dvp, double dvm, double dx) {
  const double r = 0.1;
  const double eps = 10.0 * DBL_EPSILON;
  const double theta = dvm / (dvp + TINY_NUMBER);
  const double q = (2.0 + theta) / 3.0;
  const double phi =
      smax(0.0, smin(q, smax(-0.5 * theta, smin(2.0 * theta, smin(q, 1.6)))));
  double eta = r * dx;
  eta = (dvm * dvm + dvp * dvp) / (eta * eta);
  if (eta <= 1.0 - eps) {
    return q;
  } else if (eta >= 1.0 + eps) {
    return phi;
  }
  return 0.5 * ((1.0 - (eta - 1.0) / eps) * q + (1.0 + (eta - 1.0) / eps) * phi);
}

/* src/recon/limo3_simple.hpp:57-78 */
static void rc_limo3(double qm1, double q0, double qp1, double dx, int positivity,
                     double *ql_ip1, double *qr_i) {
  const double dqp = qp1 - q0;
  const double dqm = q0 - qm1;
  *ql_ip1 = q0 + 0.5 * dqp * limo3_limiter(dqp, dqm, dx);
  *qr_i = q0 - 0.5 * dqm * limo3_limiter(dqm, dqp, dx);
  if (positivity && (*ql_ip1 <= 0.0 || *qr_i <= 0.0)) {
    double dqmm = lim_minmod(dqp, dqm);
    *ql_ip1 = q0 + 0.5 * dqmm;
    *qr_i = q0 - 0.5 * dqmm;
  }
}

/* q[0..4] = q_{i-2..i+2}; dx = cell width of the sweep direction */
void apko_recon_scalar(int recon, const double q[5], double dx, int positivity, double *ql_ip1,
                       double *qr_i) {
  switch (recon) {
  case APKO_RECON_DC: /* src/recon/dc_simple.hpp:33-36 */
    *ql_ip1 = q[2];
    *qr_i = q[2];
    break;
  case APKO_RECON_PLM:
    rc_plm(q[1], q[2], q[3], ql_ip1, qr_i);
    break;
  case APKO_RECON_PPM:
    rc_ppm(q[0], q[1], q[2], q[3], q[4], ql_ip1, qr_i);
    break;
  case APKO_RECON_WENOZ:
    rc_wenoz(q[0], q[1], q[2], q[3], q[4], ql_ip1, qr_i);
    break;
  case APKO_RECON_WENO3:
    rc_weno3(q[1], q[2], q[3], dx * dx, ql_ip1, qr_i);
    break;
  case APKO_RECON_LIMO3:
    rc_limo3(q[1], q[2], q[3], dx, positivity, ql_ip1, qr_i);
    break;
  default:
    *ql_ip1 = *qr_i = NAN;
  }
}

/* ===================================================================================== */
/* EOS helpers                                                                           */
/* ===================================================================================== */

/* src/eos/adiabatic_hydro.hpp:38-41 */
static double eos_cs(double gamma, double d, double p) { return sqrt(gamma * p / d); }

/* src/eos/adiabatic_glmmhd.hpp:46-54 */
static double eos_cf(double gamma, double d, double p, double bx, double by, double bz) {
  double asq = gamma * p;
  double ct2 = by * by + bz * bz;
  double qsq = bx * bx + ct2 + asq;
  double tmp = bx * bx + ct2 - asq;
  return sqrt(0.5 * (qsq + sqrt(tmp * tmp + 4.0 * asq * ct2)) / d);
}

/* src/eos/adiabatic_hydro.hpp:51-142 and adiabatic_glmmhd.hpp:61-167 (one cell).
 * u (cons) is mutated by the floors/ceilings exactly like the reference. Returns 0 ok,
 * 1 = negative density w/o floor, 2 = negative pressure w/o floor (PARTHENON_REQUIRE). */
static int c2p_cell(int fluid, int nscalars, double *u, double *w, int stride, double gamma,
                    double dfloor, double pfloor, double efloor, double vceil, double eceil) {
  const int nhydro = (fluid == APKO_FLUID_GLMMHD) ? NGLMMHD : NHYDRO;
  const double gm1 = gamma - 1.0;
  int err = 0;
#define U(n) u[(size_t)(n) * stride]
#define W(n) w[(size_t)(n) * stride]
  if (!(U(IDN) > 0.0 || dfloor > 0.0)) err = 1;
  U(IDN) = (U(IDN) > dfloor) ? U(IDN) : dfloor;
  W(IDN) = U(IDN);
  double di = 1.0 / U(IDN);
  W(IV1) = U(IM1) * di;
  W(IV2) = U(IM2) * di;
  W(IV3) = U(IM3) * di;
  double e_B = 0.0;
  double e_k = 0.5 * di * (SQR(U(IM1)) + SQR(U(IM2)) + SQR(U(IM3)));
  if (fluid == APKO_FLUID_GLMMHD) {
    W(IB1) = U(IB1);
    W(IB2) = U(IB2);
    W(IB3) = U(IB3);
    W(IPS) = U(IPS);
    e_B = 0.5 * (SQR(U(IB1)) + SQR(U(IB2)) + SQR(U(IB3)));
    W(IPR) = gm1 * (U(IEN) - e_k - e_B);
  } else {
    W(IPR) = gm1 * (U(IEN) - e_k);
  }
  const double v2 = SQR(W(IV1)) + SQR(W(IV2)) + SQR(W(IV3));
  if (v2 > SQR(vceil)) {
    const double v = sqrt(v2);
    W(IV1) *= vceil / v;
    W(IV2) *= vceil / v;
    W(IV3) *= vceil / v;
    U(IM1) *= vceil / v;
    U(IM2) *= vceil / v;
    U(IM3) *= vceil / v;
    double e_k_new = 0.5 * U(IDN) * SQR(vceil);
    U(IEN) -= e_k - e_k_new;
    e_k = e_k_new;
  }
  if (!(W(IPR) > 0.0 || pfloor > 0.0 || efloor > 0.0)) err = err ? err : 2;
  /* note: hydro adds e_k only; MHD adds e_k + e_B, in that order */
  if ((pfloor > 0.0) && (W(IPR) < pfloor)) {
    U(IEN) = (fluid == APKO_FLUID_GLMMHD) ? (pfloor / gm1) + e_k + e_B : (pfloor / gm1) + e_k;
    W(IPR) = pfloor;
  }
  const double eff_pfloor = gm1 * U(IDN) * efloor;
  if (W(IPR) < eff_pfloor) {
    U(IEN) = (fluid == APKO_FLUID_GLMMHD) ? (U(IDN) * efloor) + e_k + e_B
                                          : (U(IDN) * efloor) + e_k;
    W(IPR) = eff_pfloor;
  }
  const double eff_pceil = gm1 * U(IDN) * eceil;
  if (W(IPR) > eff_pceil) {
    U(IEN) = (fluid == APKO_FLUID_GLMMHD) ? (U(IDN) * eceil) + e_k + e_B
                                          : (U(IDN) * eceil) + e_k;
    W(IPR) = eff_pceil;
  }
  for (int n = nhydro; n < nhydro + nscalars; ++n) W(n) = U(n) * di;
#undef U
#undef W
  return err;
}

void apko_c2p_cell(int fluid, int nscalars, double *u, double *w, double gamma, double dfloor,
                   double pfloor, double efloor, double vceil, double eceil) {
  c2p_cell(fluid, nscalars, u, w, 1, gamma, dfloor, pfloor, efloor, vceil, eceil);
}

/* ===================================================================================== */
/* Riemann solvers: one face. wl/wr/flux are in the *unpermuted* variable layout.        */
/* ===================================================================================== */

/* src/hydro/rsolvers/hydro_hlle.hpp:39-138 */
static void rs_hydro_hlle(int ivx, const double *wl, const double *wr, double gamma,
                          double *flx) {
  const int ivy = IV1 + ((ivx - IV1) + 1) % 3;
  const int ivz = IV1 + ((ivx - IV1) + 2) % 3;
  const double gm1 = gamma - 1.0;
  const double igm1 = 1.0 / gm1;
  double L[5] = {wl[IDN], wl[ivx], wl[ivy], wl[ivz], wl[IPR]};
  double R[5] = {wr[IDN], wr[ivx], wr[ivy], wr[ivz], wr[IPR]};
  double roe[5], fl[5], fr[5], f[5];

  double sdl = sqrt(L[IDN]);
  double sdr = sqrt(R[IDN]);
  double isum = 1.0 / (sdl + sdr);
  roe[IDN] = sdl * sdr;
  roe[IV1] = (sdl * L[IV1] + sdr * R[IV1]) * isum;
  roe[IV2] = (sdl * L[IV2] + sdr * R[IV2]) * isum;
  roe[IV3] = (sdl * L[IV3] + sdr * R[IV3]) * isum;
  const double el = L[IPR] * igm1 + 0.5 * L[IDN] * (SQR(L[IV1]) + SQR(L[IV2]) + SQR(L[IV3]));
  const double er = R[IPR] * igm1 + 0.5 * R[IDN] * (SQR(R[IV1]) + SQR(R[IV2]) + SQR(R[IV3]));
  const double hroe = ((el + L[IPR]) / sdl + (er + R[IPR]) / sdr) * isum;

  const double cl = eos_cs(gamma, L[IDN], L[IPR]);
  const double cr = eos_cs(gamma, R[IDN], R[IPR]);
  double q = hroe - 0.5 * (SQR(roe[IV1]) + SQR(roe[IV2]) + SQR(roe[IV3]));
  const double a = (q < 0.0) ? 0.0 : sqrt(gm1 * q);

  const double al = smin((roe[IV1] - a), (L[IV1] - cl));
  const double ar = smax((roe[IV1] + a), (R[IV1] + cr));
  double bp = ar > 0.0 ? ar : TINY_NUMBER;
  double bm = al < 0.0 ? al : TINY_NUMBER; /* sic: +TINY (hydro_hlle.hpp:98) */

  double vxl = L[IV1] - bm;
  double vxr = R[IV1] - bp;
  fl[IDN] = L[IDN] * vxl;
  fr[IDN] = R[IDN] * vxr;
  fl[IV1] = L[IDN] * L[IV1] * vxl;
  fr[IV1] = R[IDN] * R[IV1] * vxr;
  fl[IV2] = L[IDN] * L[IV2] * vxl;
  fr[IV2] = R[IDN] * R[IV2] * vxr;
  fl[IV3] = L[IDN] * L[IV3] * vxl;
  fr[IV3] = R[IDN] * R[IV3] * vxr;
  fl[IV1] += L[IPR];
  fr[IV1] += R[IPR];
  fl[IEN] = el * vxl + L[IPR] * L[IV1];
  fr[IEN] = er * vxr + R[IPR] * R[IV1];

  double tmp = 0.0;
  if (bp != bm) tmp = 0.5 * (bp + bm) / (bp - bm);
  for (int n = 0; n < 5; ++n) f[n] = 0.5 * (fl[n] + fr[n]) + (fl[n] - fr[n]) * tmp;

  flx[IDN] = f[IDN];
  flx[ivx] = f[IV1];
  flx[ivy] = f[IV2];
  flx[ivz] = f[IV3];
  flx[IEN] = f[IEN];
}

/* src/hydro/rsolvers/hydro_hllc.hpp:31-157 */
static void rs_hydro_hllc(int ivx, const double *wl, const double *wr, double gamma,
                          double *flx) {
  const int ivy = IV1 + ((ivx - IV1) + 1) % 3;
  const int ivz = IV1 + ((ivx - IV1) + 2) % 3;
  const double gm1 = gamma - 1.0;
  const double igm1 = 1.0 / gm1;
  double L[5] = {wl[IDN], wl[ivx], wl[ivy], wl[ivz], wl[IPR]};
  double R[5] = {wr[IDN], wr[ivx], wr[ivy], wr[ivz], wr[IPR]};
  double fl[5], fr[5], f[5];

  double cl = eos_cs(gamma, L[IDN], L[IPR]);
  double cr = eos_cs(gamma, R[IDN], R[IPR]);
  double el = L[IPR] * igm1 + 0.5 * L[IDN] * (SQR(L[IV1]) + SQR(L[IV2]) + SQR(L[IV3]));
  double er = R[IPR] * igm1 + 0.5 * R[IDN] * (SQR(R[IV1]) + SQR(R[IV2]) + SQR(R[IV3]));
  double rhoa = .5 * (L[IDN] + R[IDN]);
  double ca = .5 * (cl + cr);
  double pmid = .5 * (L[IPR] + R[IPR] + (L[IV1] - R[IV1]) * rhoa * ca);

  double ql = (pmid <= L[IPR]) ? 1.0
                               : sqrt(1.0 + (gamma + 1) / (2 * gamma) * (pmid / L[IPR] - 1.0));
  double qr = (pmid <= R[IPR]) ? 1.0
                               : sqrt(1.0 + (gamma + 1) / (2 * gamma) * (pmid / R[IPR] - 1.0));

  double al = L[IV1] - cl * ql;
  double ar = R[IV1] + cr * qr;
  double bp = ar > 0.0 ? ar : (TINY_NUMBER);
  double bm = al < 0.0 ? al : -(TINY_NUMBER);

  double vxl = L[IV1] - al;
  double vxr = R[IV1] - ar;
  double tl = L[IPR] + vxl * L[IDN] * L[IV1];
  double tr = R[IPR] + vxr * R[IDN] * R[IV1];
  double ml = L[IDN] * vxl;
  double mr = -(R[IDN] * vxr);
  double am = (tl - tr) / (ml + mr);
  double cp = (ml * tr + mr * tl) / (ml + mr);
  cp = cp > 0.0 ? cp : 0.0;

  vxl = L[IV1] - bm;
  vxr = R[IV1] - bp;
  fl[IDN] = L[IDN] * vxl;
  fr[IDN] = R[IDN] * vxr;
  fl[IV1] = L[IDN] * L[IV1] * vxl + L[IPR];
  fr[IV1] = R[IDN] * R[IV1] * vxr + R[IPR];
  fl[IV2] = L[IDN] * L[IV2] * vxl;
  fr[IV2] = R[IDN] * R[IV2] * vxr;
  fl[IV3] = L[IDN] * L[IV3] * vxl;
  fr[IV3] = R[IDN] * R[IV3] * vxr;
  fl[IEN] = el * vxl + L[IPR] * L[IV1];
  fr[IEN] = er * vxr + R[IPR] * R[IV1];

  double sl, sr, sm;
  if (am >= 0.0) {
    sl = am / (am - bm);
    sr = 0.0;
    sm = -bm / (am - bm);
  } else {
    sl = 0.0;
    sr = -am / (bp - am);
    sm = bp / (bp - am);
  }
  f[IDN] = sl * fl[IDN] + sr * fr[IDN];
  f[IV1] = sl * fl[IV1] + sr * fr[IV1] + sm * cp;
  f[IV2] = sl * fl[IV2] + sr * fr[IV2];
  f[IV3] = sl * fl[IV3] + sr * fr[IV3];
  f[IEN] = sl * fl[IEN] + sr * fr[IEN] + sm * cp * am;

  flx[IDN] = f[IDN];
  flx[ivx] = f[IV1];
  flx[ivy] = f[IV2];
  flx[ivz] = f[IV3];
  flx[IEN] = f[IEN];
}

/* src/hydro/rsolvers/hydro_dc_llf.hpp:42-118 (face part; scalars handled by caller) */
static void rs_hydro_llf(int ivx, const double *wl, const double *wr, double gamma,
                         double *flx) {
  const int ivy = IV1 + ((ivx - IV1) + 1) % 3;
  const int ivz = IV1 + ((ivx - IV1) + 2) % 3;
  const double igm1 = 1.0 / (gamma - 1.0);
  double L[5] = {wl[IDN], wl[ivx], wl[ivy], wl[ivz], wl[IPR]};
  double R[5] = {wr[IDN], wr[ivx], wr[ivy], wr[ivz], wr[IPR]};
  double qa = L[IDN] * L[IV1];
  double qb = R[IDN] * R[IV1];
  double fs_d = qa + qb;
  double fs_mx = qa * L[IV1] + qb * R[IV1];
  double fs_my = qa * L[IV2] + qb * R[IV2];
  double fs_mz = qa * L[IV3] + qb * R[IV3];
  double el = L[IPR] * igm1 + 0.5 * L[IDN] * (SQR(L[IV1]) + SQR(L[IV2]) + SQR(L[IV3]));
  double er = R[IPR] * igm1 + 0.5 * R[IDN] * (SQR(R[IV1]) + SQR(R[IV2]) + SQR(R[IV3]));
  fs_mx += (L[IPR] + R[IPR]);
  double fs_e = (el + L[IPR]) * L[IV1] + (er + R[IPR]) * R[IV1];
  qa = eos_cs(gamma, L[IDN], L[IPR]);
  qb = eos_cs(gamma, R[IDN], R[IPR]);
  double a = fmax((fabs(L[IV1]) + qa), (fabs(R[IV1]) + qb));
  double du_d = a * (R[IDN] - L[IDN]);
  double du_mx = a * (R[IDN] * R[IV1] - L[IDN] * L[IV1]);
  double du_my = a * (R[IDN] * R[IV2] - L[IDN] * L[IV2]);
  double du_mz = a * (R[IDN] * R[IV3] - L[IDN] * L[IV3]);
  double du_e = a * (er - el);
  flx[IDN] = 0.5 * (fs_d - du_d);
  flx[ivx] = 0.5 * (fs_mx - du_mx);
  flx[ivy] = 0.5 * (fs_my - du_my);
  flx[ivz] = 0.5 * (fs_mz - du_mz);
  flx[IEN] = 0.5 * (fs_e - du_e);
}

/* src/hydro/rsolvers/glmmhd_dc_llf.hpp:45-155 */
static void rs_mhd_llf(int ivx, const double *wl, const double *wr, double gamma, double c_h,
                       double *flx) {
  const int ivy = IV1 + ((ivx - IV1) + 1) % 3;
  const int ivz = IV1 + ((ivx - IV1) + 2) % 3;
  const int iBx = ivx - 1 + NHYDRO, iBy = ivy - 1 + NHYDRO, iBz = ivz - 1 + NHYDRO;
  const double igm1 = 1.0 / (gamma - 1.0);
  double L[9] = {wl[IDN], wl[ivx], wl[ivy], wl[ivz], wl[IPR], wl[iBx], wl[iBy], wl[iBz], wl[IPS]};
  double R[9] = {wr[IDN], wr[ivx], wr[ivy], wr[ivz], wr[IPR], wr[iBx], wr[iBy], wr[iBz], wr[IPS]};

  double bxi = 0.5 * (L[IB1] + R[IB1]) - 0.5 / c_h * (R[IPS] - L[IPS]);
  double psii = 0.5 * (L[IPS] + R[IPS]) - 0.5 * c_h * (R[IB1] - L[IB1]);

  double qa = L[IDN] * L[IV1];
  double qb = R[IDN] * R[IV1];
  double qc = 0.5 * (SQR(L[IB2]) + SQR(L[IB3]) - SQR(bxi));
  double qd = 0.5 * (SQR(R[IB2]) + SQR(R[IB3]) - SQR(bxi));

  double fs_d = qa + qb;
  double fs_mx = qa * L[IV1] + qb * R[IV1] + qc + qd;
  double fs_my = qa * L[IV2] + qb * R[IV2] - bxi * (L[IB2] + R[IB2]);
  double fs_mz = qa * L[IV3] + qb * R[IV3] - bxi * (L[IB3] + R[IB3]);
  double fs_by = L[IB2] * L[IV1] + R[IB2] * R[IV1] - bxi * (L[IV2] + R[IV2]);
  double fs_bz = L[IB3] * L[IV1] + R[IB3] * R[IV1] - bxi * (L[IV3] + R[IV3]);

  double el = L[IPR] * igm1 + 0.5 * L[IDN] * (SQR(L[IV1]) + SQR(L[IV2]) + SQR(L[IV3])) + qc +
              SQR(bxi);
  double er = R[IPR] * igm1 + 0.5 * R[IDN] * (SQR(R[IV1]) + SQR(R[IV2]) + SQR(R[IV3])) + qd +
              SQR(bxi);
  fs_mx += (L[IPR] + R[IPR]);
  double fs_e = (el + L[IPR] + qc) * L[IV1] + (er + R[IPR] + qd) * R[IV1];
  fs_e -= bxi * (L[IB2] * L[IV2] + L[IB3] * L[IV3]);
  fs_e -= bxi * (R[IB2] * R[IV2] + R[IB3] * R[IV3]);

  qa = eos_cf(gamma, L[IDN], L[IPR], L[IB1], L[IB2], L[IB3]);
  qb = eos_cf(gamma, R[IDN], R[IPR], R[IB1], R[IB2], R[IB3]);
  double a = fmax((fabs(L[IV1]) + qa), (fabs(R[IV1]) + qb));

  double du_d = a * (R[IDN] - L[IDN]);
  double du_mx = a * (R[IDN] * R[IV1] - L[IDN] * L[IV1]);
  double du_my = a * (R[IDN] * R[IV2] - L[IDN] * L[IV2]);
  double du_mz = a * (R[IDN] * R[IV3] - L[IDN] * L[IV3]);
  double du_e = a * (er - el);
  double du_by = a * (R[IB2] - L[IB2]);
  double du_bz = a * (R[IB3] - L[IB3]);

  flx[IDN] = 0.5 * (fs_d - du_d);
  flx[ivx] = 0.5 * (fs_mx - du_mx);
  flx[ivy] = 0.5 * (fs_my - du_my);
  flx[ivz] = 0.5 * (fs_mz - du_mz);
  flx[IEN] = 0.5 * (fs_e - du_e);
  flx[iBx] = psii;
  flx[iBy] = 0.5 * (fs_by - du_by);
  flx[iBz] = 0.5 * (fs_bz - du_bz);
  flx[IPS] = SQR(c_h) * bxi;
}

/* src/hydro/rsolvers/glmmhd_hlle.hpp:27-192 */
static void rs_mhd_hlle(int ivx, const double *wl, const double *wr, double gamma, double c_h,
                        double *flx) {
  const int ivy = IV1 + ((ivx - IV1) + 1) % 3;
  const int ivz = IV1 + ((ivx - IV1) + 2) % 3;
  const int iBx = ivx - 1 + NHYDRO, iBy = ivy - 1 + NHYDRO, iBz = ivz - 1 + NHYDRO;
  const double gm1 = gamma - 1.0;
  double L[9] = {wl[IDN], wl[ivx], wl[ivy], wl[ivz], wl[IPR], wl[iBx], wl[iBy], wl[iBz], wl[IPS]};
  double R[9] = {wr[IDN], wr[ivx], wr[ivy], wr[ivz], wr[IPR], wr[iBx], wr[iBy], wr[iBz], wr[IPS]};
  double roe[9], fl[9], fr[9], f[9];

  double bxi = 0.5 * (L[IB1] + R[IB1]) - 0.5 / c_h * (R[IPS] - L[IPS]);
  double psii = 0.5 * (L[IPS] + R[IPS]) - 0.5 * c_h * (R[IB1] - L[IB1]);
  f[IB1] = psii;
  f[IPS] = SQR(c_h) * bxi;

  double sdl = sqrt(L[IDN]);
  double sdr = sqrt(R[IDN]);
  double isum = 1.0 / (sdl + sdr);
  roe[IDN] = sdl * sdr;
  roe[IV1] = (sdl * L[IV1] + sdr * R[IV1]) * isum;
  roe[IV2] = (sdl * L[IV2] + sdr * R[IV2]) * isum;
  roe[IV3] = (sdl * L[IV3] + sdr * R[IV3]) * isum;
  roe[IB2] = (sdr * L[IB2] + sdl * R[IB2]) * isum;
  roe[IB3] = (sdr * L[IB3] + sdl * R[IB3]) * isum;
  double x = 0.5 * (SQR(L[IB2] - R[IB2]) + SQR(L[IB3] - R[IB3])) / (SQR(sdl + sdr));
  double y = 0.5 * (L[IDN] + R[IDN]) / roe[IDN];

  double pbl = 0.5 * (bxi * bxi + SQR(L[IB2]) + SQR(L[IB3]));
  double pbr = 0.5 * (bxi * bxi + SQR(R[IB2]) + SQR(R[IB3]));
  double el = L[IPR] / gm1 + 0.5 * L[IDN] * (SQR(L[IV1]) + SQR(L[IV2]) + SQR(L[IV3])) + pbl;
  double er = R[IPR] / gm1 + 0.5 * R[IDN] * (SQR(R[IV1]) + SQR(R[IV2]) + SQR(R[IV3])) + pbr;
  double hroe = ((el + L[IPR] + pbl) / sdl + (er + R[IPR] + pbr) / sdr) * isum;

  double cl = eos_cf(gamma, L[IDN], L[IPR], L[IB1], L[IB2], L[IB3]);
  double cr = eos_cf(gamma, R[IDN], R[IPR], R[IB1], R[IB2], R[IB3]);

  double btsq = SQR(roe[IB2]) + SQR(roe[IB3]);
  double vaxsq = bxi * bxi / roe[IDN];
  double bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq;
  double hp = hroe - (vaxsq + btsq / roe[IDN]);
  double vsq = SQR(roe[IV1]) + SQR(roe[IV2]) + SQR(roe[IV3]);
  double twid_asq = smax((gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x), 0.0);
  double ct2 = bt_starsq / roe[IDN];
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = sqrt(tdif * tdif + 4.0 * twid_asq * ct2);
  double cfsq = 0.5 * (tsum + cf2_cs2);
  double a = sqrt(cfsq);

  double al = smin((roe[IV1] - a), (L[IV1] - cl));
  double ar = smax((roe[IV1] + a), (R[IV1] + cr));
  double bp = ar > 0.0 ? ar : 0.0;
  double bm = al < 0.0 ? al : 0.0;

  double vxl = L[IV1] - bm;
  double vxr = R[IV1] - bp;
  fl[IDN] = L[IDN] * vxl;
  fr[IDN] = R[IDN] * vxr;
  fl[IV1] = L[IDN] * L[IV1] * vxl + pbl - SQR(bxi);
  fr[IV1] = R[IDN] * R[IV1] * vxr + pbr - SQR(bxi);
  fl[IV2] = L[IDN] * L[IV2] * vxl - bxi * L[IB2];
  fr[IV2] = R[IDN] * R[IV2] * vxr - bxi * R[IB2];
  fl[IV3] = L[IDN] * L[IV3] * vxl - bxi * L[IB3];
  fr[IV3] = R[IDN] * R[IV3] * vxr - bxi * R[IB3];
  fl[IV1] += L[IPR];
  fr[IV1] += R[IPR];
  fl[IEN] = el * vxl + L[IV1] * (L[IPR] + pbl - bxi * bxi);
  fr[IEN] = er * vxr + R[IV1] * (R[IPR] + pbr - bxi * bxi);
  fl[IEN] -= bxi * (L[IB2] * L[IV2] + L[IB3] * L[IV3]);
  fr[IEN] -= bxi * (R[IB2] * R[IV2] + R[IB3] * R[IV3]);
  fl[IB2] = L[IB2] * vxl - bxi * L[IV2];
  fr[IB2] = R[IB2] * vxr - bxi * R[IV2];
  fl[IB3] = L[IB3] * vxl - bxi * L[IV3];
  fr[IB3] = R[IB3] * vxr - bxi * R[IV3];

  double tmp = 0.0;
  if (bp != bm) tmp = 0.5 * (bp + bm) / (bp - bm);
  f[IDN] = 0.5 * (fl[IDN] + fr[IDN]) + (fl[IDN] - fr[IDN]) * tmp;
  f[IV1] = 0.5 * (fl[IV1] + fr[IV1]) + (fl[IV1] - fr[IV1]) * tmp;
  f[IV2] = 0.5 * (fl[IV2] + fr[IV2]) + (fl[IV2] - fr[IV2]) * tmp;
  f[IV3] = 0.5 * (fl[IV3] + fr[IV3]) + (fl[IV3] - fr[IV3]) * tmp;
  f[IEN] = 0.5 * (fl[IEN] + fr[IEN]) + (fl[IEN] - fr[IEN]) * tmp;
  f[IB2] = 0.5 * (fl[IB2] + fr[IB2]) + (fl[IB2] - fr[IB2]) * tmp;
  f[IB3] = 0.5 * (fl[IB3] + fr[IB3]) + (fl[IB3] - fr[IB3]) * tmp;

  flx[IDN] = f[IDN];
  flx[ivx] = f[IV1];
  flx[ivy] = f[IV2];
  flx[ivz] = f[IV3];
  flx[IEN] = f[IEN];
  flx[iBx] = f[IB1];
  flx[iBy] = f[IB2];
  flx[iBz] = f[IB3];
  flx[IPS] = f[IPS];
}

typedef struct {
  double d, mx, my, mz, e, by, bz;
} cons1d;

#define SMALL_NUMBER 1.0e-8

/* src/hydro/rsolvers/glmmhd_hlld.hpp:39-395 */
static void rs_mhd_hlld(int ivx, const double *wl, const double *wr, double gamma, double c_h,
                        double *flx) {
  const int ivy = IV1 + ((ivx - IV1) + 1) % 3;
  const int ivz = IV1 + ((ivx - IV1) + 2) % 3;
  const int iBx = ivx - 1 + NHYDRO, iBy = ivy - 1 + NHYDRO, iBz = ivz - 1 + NHYDRO;
  const double gm1 = gamma - 1.0;
  const double igm1 = 1.0 / gm1;
  double L[9] = {wl[IDN], wl[ivx], wl[ivy], wl[ivz], wl[IPR], wl[iBx], wl[iBy], wl[iBz], wl[IPS]};
  double R[9] = {wr[IDN], wr[ivx], wr[ivy], wr[ivz], wr[IPR], wr[iBx], wr[iBy], wr[iBz], wr[IPS]};
  double f[9], spd[5];
  cons1d ul, ur, ulst, uldst, urdst, urst, fl, fr;

  double bxi = 0.5 * (L[IB1] + R[IB1]) - 0.5 / c_h * (R[IPS] - L[IPS]);
  double psii = 0.5 * (L[IPS] + R[IPS]) - 0.5 * c_h * (R[IB1] - L[IB1]);
  f[IB1] = psii;
  f[IPS] = SQR(c_h) * bxi;

  double bxsq = bxi * bxi;
  double pbl = 0.5 * (bxsq + (SQR(L[IB2]) + SQR(L[IB3])));
  double pbr = 0.5 * (bxsq + (SQR(R[IB2]) + SQR(R[IB3])));
  double kel = 0.5 * L[IDN] * (SQR(L[IV1]) + (SQR(L[IV2]) + SQR(L[IV3])));
  double ker = 0.5 * R[IDN] * (SQR(R[IV1]) + (SQR(R[IV2]) + SQR(R[IV3])));

  ul.d = L[IDN];
  ul.mx = L[IV1] * ul.d;
  ul.my = L[IV2] * ul.d;
  ul.mz = L[IV3] * ul.d;
  ul.e = L[IPR] * igm1 + kel + pbl;
  ul.by = L[IB2];
  ul.bz = L[IB3];

  ur.d = R[IDN];
  ur.mx = R[IV1] * ur.d;
  ur.my = R[IV2] * ur.d;
  ur.mz = R[IV3] * ur.d;
  ur.e = R[IPR] * igm1 + ker + pbr;
  ur.by = R[IB2];
  ur.bz = R[IB3];

  const double cfl = eos_cf(gamma, L[IDN], L[IPR], L[IB1], L[IB2], L[IB3]);
  const double cfr = eos_cf(gamma, R[IDN], R[IPR], R[IB1], R[IB2], R[IB3]);
  spd[0] = smin(L[IV1] - cfl, R[IV1] - cfr);
  spd[4] = smax(L[IV1] + cfl, R[IV1] + cfr);

  double ptl = L[IPR] + pbl;
  double ptr = R[IPR] + pbr;

  fl.d = ul.mx;
  fl.mx = ul.mx * L[IV1] + ptl - bxsq;
  fl.my = ul.my * L[IV1] - bxi * ul.by;
  fl.mz = ul.mz * L[IV1] - bxi * ul.bz;
  fl.e = L[IV1] * (ul.e + ptl - bxsq) - bxi * (L[IV2] * ul.by + L[IV3] * ul.bz);
  fl.by = ul.by * L[IV1] - bxi * L[IV2];
  fl.bz = ul.bz * L[IV1] - bxi * L[IV3];

  fr.d = ur.mx;
  fr.mx = ur.mx * R[IV1] + ptr - bxsq;
  fr.my = ur.my * R[IV1] - bxi * ur.by;
  fr.mz = ur.mz * R[IV1] - bxi * ur.bz;
  fr.e = R[IV1] * (ur.e + ptr - bxsq) - bxi * (R[IV2] * ur.by + R[IV3] * ur.bz);
  fr.by = ur.by * R[IV1] - bxi * R[IV2];
  fr.bz = ur.bz * R[IV1] - bxi * R[IV3];

  double sdl = spd[0] - L[IV1];
  double sdr = spd[4] - R[IV1];
  spd[2] = (sdr * ur.mx - sdl * ul.mx + (ptl - ptr)) / (sdr * ur.d - sdl * ul.d);

  double sdml = spd[0] - spd[2];
  double sdmr = spd[4] - spd[2];
  double sdml_inv = 1.0 / sdml;
  double sdmr_inv = 1.0 / sdmr;
  ulst.d = ul.d * sdl * sdml_inv;
  urst.d = ur.d * sdr * sdmr_inv;
  double ulst_d_inv = 1.0 / ulst.d;
  double urst_d_inv = 1.0 / urst.d;
  double sqrtdl = sqrt(ulst.d);
  double sqrtdr = sqrt(urst.d);

  spd[1] = spd[2] - fabs(bxi) / sqrtdl;
  spd[3] = spd[2] + fabs(bxi) / sqrtdr;

  double ptstl = ptl + ul.d * sdl * (spd[2] - L[IV1]);
  double ptstr = ptr + ur.d * sdr * (spd[2] - R[IV1]);
  double ptst = 0.5 * (ptstr + ptstl);

  ulst.mx = ulst.d * spd[2];
  if (fabs(ul.d * sdl * sdml - bxsq) < (SMALL_NUMBER)*ptst) {
    ulst.my = ulst.d * L[IV2];
    ulst.mz = ulst.d * L[IV3];
    ulst.by = ul.by;
    ulst.bz = ul.bz;
  } else {
    double tmp = bxi * (sdl - sdml) / (ul.d * sdl * sdml - bxsq);
    ulst.my = ulst.d * (L[IV2] - ul.by * tmp);
    ulst.mz = ulst.d * (L[IV3] - ul.bz * tmp);
    tmp = (ul.d * SQR(sdl) - bxsq) / (ul.d * sdl * sdml - bxsq);
    ulst.by = ul.by * tmp;
    ulst.bz = ul.bz * tmp;
  }
  double vbstl = (ulst.mx * bxi + (ulst.my * ulst.by + ulst.mz * ulst.bz)) * ulst_d_inv;
  ulst.e = (sdl * ul.e - ptl * L[IV1] + ptst * spd[2] +
            bxi * (L[IV1] * bxi + (L[IV2] * ul.by + L[IV3] * ul.bz) - vbstl)) *
           sdml_inv;

  urst.mx = urst.d * spd[2];
  if (fabs(ur.d * sdr * sdmr - bxsq) < (SMALL_NUMBER)*ptst) {
    urst.my = urst.d * R[IV2];
    urst.mz = urst.d * R[IV3];
    urst.by = ur.by;
    urst.bz = ur.bz;
  } else {
    double tmp = bxi * (sdr - sdmr) / (ur.d * sdr * sdmr - bxsq);
    urst.my = urst.d * (R[IV2] - ur.by * tmp);
    urst.mz = urst.d * (R[IV3] - ur.bz * tmp);
    tmp = (ur.d * SQR(sdr) - bxsq) / (ur.d * sdr * sdmr - bxsq);
    urst.by = ur.by * tmp;
    urst.bz = ur.bz * tmp;
  }
  double vbstr = (urst.mx * bxi + (urst.my * urst.by + urst.mz * urst.bz)) * urst_d_inv;
  urst.e = (sdr * ur.e - ptr * R[IV1] + ptst * spd[2] +
            bxi * (R[IV1] * bxi + (R[IV2] * ur.by + R[IV3] * ur.bz) - vbstr)) *
           sdmr_inv;

  if (0.5 * bxsq < (SMALL_NUMBER)*ptst) {
    uldst = ulst;
    urdst = urst;
  } else {
    double invsumd = 1.0 / (sqrtdl + sqrtdr);
    double bxsig = (bxi > 0.0 ? 1.0 : -1.0);
    uldst.d = ulst.d;
    urdst.d = urst.d;
    uldst.mx = ulst.mx;
    urdst.mx = urst.mx;
    double tmp = invsumd * (sqrtdl * (ulst.my * ulst_d_inv) + sqrtdr * (urst.my * urst_d_inv) +
                            bxsig * (urst.by - ulst.by));
    uldst.my = uldst.d * tmp;
    urdst.my = urdst.d * tmp;
    tmp = invsumd * (sqrtdl * (ulst.mz * ulst_d_inv) + sqrtdr * (urst.mz * urst_d_inv) +
                     bxsig * (urst.bz - ulst.bz));
    uldst.mz = uldst.d * tmp;
    urdst.mz = urdst.d * tmp;
    tmp = invsumd * (sqrtdl * urst.by + sqrtdr * ulst.by +
                     bxsig * sqrtdl * sqrtdr * ((urst.my * urst_d_inv) - (ulst.my * ulst_d_inv)));
    uldst.by = urdst.by = tmp;
    tmp = invsumd * (sqrtdl * urst.bz + sqrtdr * ulst.bz +
                     bxsig * sqrtdl * sqrtdr * ((urst.mz * urst_d_inv) - (ulst.mz * ulst_d_inv)));
    uldst.bz = urdst.bz = tmp;
    tmp = spd[2] * bxi + (uldst.my * uldst.by + uldst.mz * uldst.bz) / uldst.d;
    uldst.e = ulst.e - sqrtdl * bxsig * (vbstl - tmp);
    urdst.e = urst.e + sqrtdr * bxsig * (vbstr - tmp);
  }

#define SCALE_DIFF(out, s, a, b)  \
  out.d = s * (a.d - b.d);        \
  out.mx = s * (a.mx - b.mx);     \
  out.my = s * (a.my - b.my);     \
  out.mz = s * (a.mz - b.mz);     \
  out.e = s * (a.e - b.e);        \
  out.by = s * (a.by - b.by);     \
  out.bz = s * (a.bz - b.bz);
  SCALE_DIFF(uldst, spd[1], uldst, ulst)
  SCALE_DIFF(ulst, spd[0], ulst, ul)
  SCALE_DIFF(urdst, spd[3], urdst, urst)
  SCALE_DIFF(urst, spd[4], urst, ur)
#undef SCALE_DIFF

  if (spd[0] >= 0.0) {
    f[IDN] = fl.d; f[IV1] = fl.mx; f[IV2] = fl.my; f[IV3] = fl.mz; f[IEN] = fl.e;
    f[IB2] = fl.by; f[IB3] = fl.bz;
  } else if (spd[4] <= 0.0) {
    f[IDN] = fr.d; f[IV1] = fr.mx; f[IV2] = fr.my; f[IV3] = fr.mz; f[IEN] = fr.e;
    f[IB2] = fr.by; f[IB3] = fr.bz;
  } else if (spd[1] >= 0.0) {
    f[IDN] = fl.d + ulst.d; f[IV1] = fl.mx + ulst.mx; f[IV2] = fl.my + ulst.my;
    f[IV3] = fl.mz + ulst.mz; f[IEN] = fl.e + ulst.e; f[IB2] = fl.by + ulst.by;
    f[IB3] = fl.bz + ulst.bz;
  } else if (spd[2] >= 0.0) {
    f[IDN] = fl.d + ulst.d + uldst.d; f[IV1] = fl.mx + ulst.mx + uldst.mx;
    f[IV2] = fl.my + ulst.my + uldst.my; f[IV3] = fl.mz + ulst.mz + uldst.mz;
    f[IEN] = fl.e + ulst.e + uldst.e; f[IB2] = fl.by + ulst.by + uldst.by;
    f[IB3] = fl.bz + ulst.bz + uldst.bz;
  } else if (spd[3] > 0.0) {
    f[IDN] = fr.d + urst.d + urdst.d; f[IV1] = fr.mx + urst.mx + urdst.mx;
    f[IV2] = fr.my + urst.my + urdst.my; f[IV3] = fr.mz + urst.mz + urdst.mz;
    f[IEN] = fr.e + urst.e + urdst.e; f[IB2] = fr.by + urst.by + urdst.by;
    f[IB3] = fr.bz + urst.bz + urdst.bz;
  } else {
    f[IDN] = fr.d + urst.d; f[IV1] = fr.mx + urst.mx; f[IV2] = fr.my + urst.my;
    f[IV3] = fr.mz + urst.mz; f[IEN] = fr.e + urst.e; f[IB2] = fr.by + urst.by;
    f[IB3] = fr.bz + urst.bz;
  }

  flx[IDN] = f[IDN];
  flx[ivx] = f[IV1];
  flx[ivy] = f[IV2];
  flx[ivz] = f[IV3];
  flx[IEN] = f[IEN];
  flx[iBx] = f[IB1];
  flx[iBy] = f[IB2];
  flx[iBz] = f[IB3];
  flx[IPS] = f[IPS];
}

static void riemann_face(int fluid, int solver, int ivx, const double *wl, const double *wr,
                         double gamma, double c_h, double *flx) {
  const int nh = (fluid == APKO_FLUID_GLMMHD) ? NGLMMHD : NHYDRO;
  if (solver == APKO_RS_NONE) { /* rsolvers.hpp:35-63 */
    for (int n = 0; n < nh; ++n) flx[n] = 0.0;
    return;
  }
  if (fluid == APKO_FLUID_EULER) {
    if (solver == APKO_RS_HLLE) rs_hydro_hlle(ivx, wl, wr, gamma, flx);
    else if (solver == APKO_RS_HLLC) rs_hydro_hllc(ivx, wl, wr, gamma, flx);
    else if (solver == APKO_RS_LLF) rs_hydro_llf(ivx, wl, wr, gamma, flx);
    else for (int n = 0; n < nh; ++n) flx[n] = NAN;
  } else {
    if (solver == APKO_RS_HLLE) rs_mhd_hlle(ivx, wl, wr, gamma, c_h, flx);
    else if (solver == APKO_RS_HLLD) rs_mhd_hlld(ivx, wl, wr, gamma, c_h, flx);
    else if (solver == APKO_RS_LLF) rs_mhd_llf(ivx, wl, wr, gamma, c_h, flx);
    else for (int n = 0; n < nh; ++n) flx[n] = NAN;
  }
}

void apko_riemann_face(int fluid, int solver, int ivx, const double *wl, const double *wr,
                       double gamma, double c_h, double *flux) {
  riemann_face(fluid, solver, ivx, wl, wr, gamma, c_h, flux);
}

/* ===================================================================================== */
/* Mesh / engine                                                                         */
/* ===================================================================================== */

typedef struct {
  double *cons, *prim, *u1, *flux[3];
  double *my0, *yjm2; /* RKL2 registers "MY0", "Yjm2" (hydro_driver.cpp:211-212) */
  double xmin[3]; /* coords xmin_ (incl. ghost offset), uniform_coordinates.hpp:32-36 */
  double dx[3];
  double bxmin[3], bxmax[3]; /* physical block bounds */
  int loc[3];
  /* multilevel meshes */
  int level;       /* refinement level (0 = root grid); loc[] is the location at this level */
  double *coarse;  /* coarse buffer [nvar][nc3][nc2][nc1], nc = mb/2 + 2 ng (Variable::coarse_s) */
  int has_coarser; /* some neighbour is one level coarser */
  /* per neighbour offset (o3+1)*9+(o2+1)*3+(o1+1): 0 none, 1 same level, 2 coarser, 3 finer */
  int ntype[27];
  int nidx[27][4]; /* block index: the neighbour (types 1, 2) or its up to 4 abutting children */
  int nchild[27];
  /* adaptive refinement: MeshRefinement::refine_flag_, deref_count_ (mesh_refinement.hpp) */
  int refine_flag, deref_count;
} block_t;

struct apko_engine {
  apko_params p;
  int ndim, nvar, nhydro;
  int nb[3], nblocks;
  int multilevel; /* static refinement active */
  int nc[3];      /* coarse buffer extents */
  int maxlevel;
  int n[3];      /* array extents incl. ghosts */
  int is[3], ie[3];
  size_t ncell;  /* n1*n2*n3 */
  block_t *blk;
  /* driver state (driver.cpp, hydro.cpp Params) */
  double time, dt, tlim;
  int ncycle;
  double c_h, mindx, dt_hyp;
  double new_dt; /* allowed dt voted by the last EstimateTimestep */
  long long fofc_total;
  double dt_diff; /* "dt_diff" Param (hydro.cpp:700, 961-962) */
  int last_s_rkl;
  /* adaptive refinement: the problem generator is re-run while Mesh::Initialize refines
   * (mesh.cpp:746-840); Mesh::nbnew / nbdel */
  int pgen_id, pgen_npars;
  double pgen_pars[16];
  int nbnew, nbdel;
  int *amr_last; /* input of the last apko_amr_remesh: {level, lx1, lx2, lx3, refine_flag} per block */
  int amr_last_n;
  /* linear wave pgen state (linear_wave.cpp:43-51) */
  struct {
    int set, wave_flag;
    double amp, vflow, sin_a2, cos_a2, sin_a3, cos_a3, lambda, k_par, d0, p0, u0, gam, gm1;
    double ev[5], rem[5][5];
    double tlim_period;
    /* magnetised wave (linear_wave_mhd.cpp:40-48) */
    int mhd;
    double bx0, by0, bz0, dby, dbz;
    double ev7[7], rem7[7][7];
  } lw;
};

void apko_default_params(apko_params *p) {
  memset(p, 0, sizeof(*p));
  p->fluid = APKO_FLUID_EULER;
  p->recon = APKO_RECON_PLM;
  p->riemann = APKO_RS_HLLE;
  p->integrator = APKO_INT_VL2;
  for (int d = 0; d < 3; ++d) {
    p->nx[d] = 1;
    p->mb[d] = 1;
    p->xmin[d] = -0.5;
    p->xmax[d] = 0.5;
  }
  p->ng = 2;
  p->gamma = 5.0 / 3.0;
  p->cfl = 0.3;
  p->dfloor = p->pfloor = p->efloor = -1.0; /* hydro.cpp:507-510 */
  p->vceil = p->eceil = INFINITY;           /* hydro.cpp:525-529 */
  p->glmmhd_alpha = 0.1;                    /* hydro.cpp:295 */
  p->cfl_diff = -1.0;
  p->rkl2_max_dt_ratio = -1.0;              /* hydro.cpp:688 */
  p->conduction_sat_phi = 0.3;              /* hydro.cpp:563-564 */
  p->numlevel = 1;
  p->derefine_count = 10;                   /* mesh_refinement.cpp:56 */
}

/* mesh/forest/logical_location.cpp:61-74 */
static double index_to_sym_coord(int index, int bloc, int nrange) {
  int64_t noffset = index - (nrange) / 2;
  int64_t noffset_ceil = index - (nrange + 1) / 2;
  return (double)(noffset + noffset_ceil + (int64_t)bloc) / (2.0 * nrange);
}
/* defs.hpp:106-109 */
static double sym_logical_to_actual(double u, double xmin, double xmax) {
  return 0.5 * (xmin + xmax) + (u * xmax - u * xmin);
}
static int max_pow2_divisor(int in) { return in & (~(in - 1)); }

/* Block physical bounds: forest.cpp:66-141 (trees) + tree.cpp:296-318 (block in tree) */
static void block_bounds_level(const apko_engine *e, int d, int level, int ib, double *bmin,
                               double *bmax);
__attribute__((unused)) static void block_bounds(const apko_engine *e, int d, int ib, double *bmin, double *bmax) {
  block_bounds_level(e, d, 0, ib, bmin, bmax);
}
/* a block at refinement `level` with location ib (in units of blocks of that level): the tree it
 * lives in is cut 2^(log2(mcp2) + level) times (tree.cpp:296-318) */
static void block_bounds_level(const apko_engine *e, int d, int level, int ib, double *bmin,
                               double *bmax) {
  const apko_params *p = &e->p;
  int any_sym = 0, mcp2 = 1 << 30;
  for (int dd = 0; dd < 3; ++dd) {
    if (p->nx[dd] == 1) {
      any_sym = 1;
      continue;
    }
    int m = max_pow2_divisor(e->nb[dd]);
    if (m < mcp2) mcp2 = m;
  }
  /* forest.cpp:77-82: hitting a symmetry dir resets the divisor to 1; later dirs can only
   * lower it further, earlier ones are overwritten -> effectively 1 if x3 (last) is
   * symmetric; if only an earlier dir is symmetric the min continues from 1 as well. */
  if (any_sym) mcp2 = 1;
  if (p->nx[d] == 1) {
    double u0 = index_to_sym_coord(0, 0, 1), u1 = index_to_sym_coord(0, 2, 1);
    *bmin = sym_logical_to_actual(u0, p->xmin[d], p->xmax[d]);
    *bmax = sym_logical_to_actual(u1, p->xmin[d], p->xmax[d]);
    return;
  }
  int ntree = e->nb[d] / mcp2;
  const int per_tree = mcp2 << level;
  int it = ib / per_tree, lx = ib % per_tree;
  double tmin = sym_logical_to_actual(index_to_sym_coord(it, 0, ntree), p->xmin[d], p->xmax[d]);
  double tmax = sym_logical_to_actual(index_to_sym_coord(it, 2, ntree), p->xmin[d], p->xmax[d]);
  *bmin = sym_logical_to_actual(index_to_sym_coord(lx, 0, per_tree), tmin, tmax);
  *bmax = sym_logical_to_actual(index_to_sym_coord(lx, 2, per_tree), tmin, tmax);
}

static inline size_t IDX(const apko_engine *e, int n, int k, int j, int i) {
  return (((size_t)n * e->n[2] + k) * e->n[1] + j) * e->n[0] + i;
}

/* uniform_coordinates.hpp:110-114 */
static inline double xc(const block_t *b, int d, int idx) {
  return b->xmin[d] + (idx + 0.5) * b->dx[d];
}

/* ---- static mesh refinement: the leaf set Parthenon's forest ends up with ------------------- */
typedef struct {
  int level, lx[3];
} leaf_t;

static int leaf_find(const leaf_t *L, int n, int level, const int *lx) {
  for (int q = 0; q < n; ++q)
    if (L[q].level == level && L[q].lx[0] == lx[0] && L[q].lx[1] == lx[1] && L[q].lx[2] == lx[2])
      return q;
  return -1;
}

/* leaf covering location lx of `level` (the leaf itself or an ancestor); -1: region is refined
 * further (or outside) */
static int leaf_cover(const leaf_t *L, int n, int level, const int *lx, int *found_level) {
  for (int l = level; l >= 0; --l) {
    int a[3] = {lx[0] >> (level - l), lx[1] >> (level - l), lx[2] >> (level - l)};
    int q = leaf_find(L, n, l, a);
    if (q >= 0) {
      if (found_level) *found_level = l;
      return q;
    }
  }
  return -1;
}

/* mesh.cpp:1076-1096 GetStaticRefLLIndexRange */
static void static_ref_range(const apko_engine *e, int d, int ref_level, double rmin, double rmax,
                             int *lo, int *hi) {
  const apko_params *p = &e->p;
  const int lxtot = e->nb[d] * (1 << ref_level);
  int lxmin, lxmax;
  for (lxmin = 0; lxmin < lxtot; lxmin++) {
    double r = index_to_sym_coord(lxmin + 1, 0, lxtot);
    if (sym_logical_to_actual(r, p->xmin[d], p->xmax[d]) > rmin) break;
  }
  for (lxmax = lxmin; lxmax < lxtot; lxmax++) {
    double r = index_to_sym_coord(lxmax + 1, 0, lxtot);
    if (sym_logical_to_actual(r, p->xmin[d], p->xmax[d]) >= rmax) break;
  }
  if (lxmin % 2 == 1) lxmin--;
  if (lxmax % 2 == 0) lxmax++;
  *lo = lxmin;
  *hi = lxmax;
}

/* Root grid refined until every block requested by a <parthenon/static_refinement> region exists
 * (mesh.cpp:1100-1165) and neighbouring leaves (faces, edges, corners; periodic wrap) differ by
 * at most one level -- the minimal properly nested tree, which is what forest.AddMeshBlock builds */
static int build_leaves(const apko_engine *e, leaf_t **out) {
  const apko_params *p = &e->p;
  int cap = 64, n = 0;
  leaf_t *L = (leaf_t *)malloc(cap * sizeof(leaf_t));
  for (int z = 0; z < e->nb[2]; ++z)
    for (int y = 0; y < e->nb[1]; ++y)
      for (int x = 0; x < e->nb[0]; ++x) {
        if (n == cap) L = (leaf_t *)realloc(L, (cap *= 2) * sizeof(leaf_t));
        L[n].level = 0;
        L[n].lx[0] = x; L[n].lx[1] = y; L[n].lx[2] = z;
        n++;
      }
  int rlo[4][3], rhi[4][3];
  for (int r = 0; r < p->nref; ++r)
    for (int d = 0; d < 3; ++d) {
      rlo[r][d] = 0;
      rhi[r][d] = 0;
      if (d < e->ndim)
        static_ref_range(e, d, p->ref_level[r], p->ref_xmin[r][d], p->ref_xmax[r][d], &rlo[r][d],
                         &rhi[r][d]);
    }
  const int nchild = 1 << e->ndim;
  for (;;) {
    char *mark = (char *)calloc(n, 1);
    int any = 0;
    for (int q = 0; q < n; ++q) {
      /* (1) a requested block lies below this leaf */
      for (int r = 0; r < p->nref && !mark[q]; ++r) {
        const int dl = p->ref_level[r] - L[q].level;
        if (dl <= 0) continue;
        int hit = 1;
        for (int d = 0; d < e->ndim; ++d) {
          const int lo = L[q].lx[d] << dl, hi = ((L[q].lx[d] + 1) << dl) - 1;
          if (hi < rlo[r][d] || lo > rhi[r][d]) hit = 0;
        }
        if (hit) mark[q] = 1;
      }
      /* (2) proper nesting: whatever covers a neighbour location may be at most one level coarser */
      if (L[q].level >= 2) {
        const int lev = L[q].level;
        for (int o3 = (e->ndim > 2 ? -1 : 0); o3 <= (e->ndim > 2 ? 1 : 0); ++o3)
          for (int o2 = (e->ndim > 1 ? -1 : 0); o2 <= (e->ndim > 1 ? 1 : 0); ++o2)
            for (int o1 = -1; o1 <= 1; ++o1) {
              int off[3] = {o1, o2, o3}, nl[3], valid = 1;
              for (int d = 0; d < 3; ++d) {
                const int tot = e->nb[d] << lev;
                nl[d] = L[q].lx[d] + off[d];
                if (nl[d] < 0 || nl[d] >= tot) {
                  const int face = 2 * d + (nl[d] < 0 ? 0 : 1);
                  if (p->bc[face] != APKO_BC_PERIODIC) valid = 0;
                  nl[d] = (nl[d] + tot) % tot;
                }
              }
              if (!valid) continue;
              int fl = 0;
              const int c = leaf_cover(L, n, lev, nl, &fl);
              if (c >= 0 && fl < lev - 1) mark[c] = 1;
            }
      }
    }
    int n_new = n;
    for (int q = 0; q < n; ++q) any |= mark[q];
    if (!any) {
      free(mark);
      break;
    }
    for (int q = 0; q < n; ++q) {
      if (!mark[q]) continue;
      const leaf_t parent = L[q];
      for (int c = 0; c < nchild; ++c) {
        leaf_t ch;
        ch.level = parent.level + 1;
        for (int d = 0; d < 3; ++d)
          ch.lx[d] = (d < e->ndim) ? 2 * parent.lx[d] + ((c >> d) & 1) : 0;
        if (c == 0) {
          L[q] = ch; /* first child replaces the parent */
        } else {
          if (n_new == cap) L = (leaf_t *)realloc(L, (cap *= 2) * sizeof(leaf_t));
          L[n_new++] = ch;
        }
      }
    }
    n = n_new;
    free(mark);
  }
  *out = L;
  return n;
}

static void build_neighbor_tables(apko_engine *e) {
  const apko_params *p = &e->p;
  leaf_t *L = (leaf_t *)malloc(e->nblocks * sizeof(leaf_t));
  for (int b = 0; b < e->nblocks; ++b) {
    L[b].level = e->blk[b].level;
    for (int d = 0; d < 3; ++d) L[b].lx[d] = e->blk[b].loc[d];
  }
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    B->has_coarser = 0;
    const int lev = B->level;
    for (int o3 = -1; o3 <= 1; ++o3)
      for (int o2 = -1; o2 <= 1; ++o2)
        for (int o1 = -1; o1 <= 1; ++o1) {
          const int idx = (o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1);
          B->ntype[idx] = 0;
          B->nchild[idx] = 0;
          const int off[3] = {o1, o2, o3};
          if (!o1 && !o2 && !o3) continue;
          int nl[3], valid = 1;
          for (int d = 0; d < 3; ++d) {
            if (off[d] != 0 && d >= e->ndim) valid = 0;
            const int tot = e->nb[d] << lev;
            nl[d] = B->loc[d] + off[d];
            if (d < e->ndim && (nl[d] < 0 || nl[d] >= tot)) {
              const int face = 2 * d + (nl[d] < 0 ? 0 : 1);
              if (p->bc[face] != APKO_BC_PERIODIC) valid = 0;
              nl[d] = (nl[d] + tot) % tot;
            }
          }
          if (!valid) continue;
          int fl = 0;
          const int c = leaf_cover(L, e->nblocks, lev, nl, &fl);
          if (c >= 0 && fl == lev) {
            B->ntype[idx] = 1;
            B->nidx[idx][0] = c;
            B->nchild[idx] = 1;
          } else if (c >= 0) { /* fl == lev - 1 by construction */
            B->ntype[idx] = 2;
            B->nidx[idx][0] = c;
            B->nchild[idx] = 1;
            B->has_coarser = 1;
          } else {
            /* finer: the children of nl that touch this block */
            B->ntype[idx] = 3;
            int cnt = 0;
            for (int c3 = 0; c3 < 2; ++c3)
              for (int c2 = 0; c2 < 2; ++c2)
                for (int c1 = 0; c1 < 2; ++c1) {
                  const int cb[3] = {c1, c2, c3};
                  int ok = 1, cl[3];
                  for (int d = 0; d < 3; ++d) {
                    if (d >= e->ndim) {
                      if (cb[d]) ok = 0;
                      cl[d] = 0;
                      continue;
                    }
                    if (off[d] > 0 && cb[d] != 0) ok = 0;
                    if (off[d] < 0 && cb[d] != 1) ok = 0;
                    cl[d] = 2 * nl[d] + cb[d];
                  }
                  if (!ok) continue;
                  const int f = leaf_find(L, e->nblocks, lev + 1, cl);
                  if (f < 0) {
                    fprintf(stderr, "apk_oracle: mesh is not properly nested\n");
                    abort();
                  }
                  B->nidx[idx][cnt++] = f;
                }
            B->nchild[idx] = cnt;
          }
        }
  }
  free(L);
}

static void block_alloc(const apko_engine *e, block_t *B) {
  const size_t nfield = e->ncell * e->nvar;
  B->cons = (double *)calloc(nfield, sizeof(double));
  B->prim = (double *)calloc(nfield, sizeof(double));
  B->u1 = (double *)calloc(nfield, sizeof(double));
  if (e->p.diffint == 2) { /* RKL2 registers "MY0", "Yjm2" (hydro_driver.cpp:211-212) */
    B->my0 = (double *)calloc(nfield, sizeof(double));
    B->yjm2 = (double *)calloc(nfield, sizeof(double));
  }
  for (int d = 0; d < 3; ++d) B->flux[d] = (double *)calloc(nfield, sizeof(double));
  if (e->multilevel)
    B->coarse = (double *)calloc((size_t)e->nc[0] * e->nc[1] * e->nc[2] * e->nvar, sizeof(double));
}

static void block_free(block_t *B) {
  free(B->cons);
  free(B->prim);
  free(B->u1);
  free(B->my0);
  free(B->yjm2);
  free(B->coarse);
  for (int d = 0; d < 3; ++d) free(B->flux[d]);
}

/* bounds and coordinates from (level, loc) */
static void block_geometry(const apko_engine *e, block_t *B) {
  const apko_params *p = &e->p;
  for (int d = 0; d < 3; ++d) {
    block_bounds_level(e, d, B->level, B->loc[d], &B->bxmin[d], &B->bxmax[d]);
    /* uniform_coordinates.hpp:32-36 */
    B->dx[d] = (B->bxmax[d] - B->bxmin[d]) / p->mb[d];
    int istart = (p->nx[d] == 1) ? 0 : p->ng;
    B->xmin[d] = B->bxmin[d] - istart * B->dx[d];
  }
}

apko_engine *apko_create(const apko_params *p) {
  apko_engine *e = (apko_engine *)calloc(1, sizeof(*e));
  e->p = *p;
  e->nhydro = (p->fluid == APKO_FLUID_GLMMHD) ? NGLMMHD : NHYDRO;
  e->nvar = e->nhydro + p->nscalars;
  e->ndim = (p->nx[2] > 1) ? 3 : ((p->nx[1] > 1) ? 2 : 1);
  e->nblocks = 1;
  for (int d = 0; d < 3; ++d) {
    if (p->nx[d] % p->mb[d] != 0) {
      free(e);
      return NULL;
    }
    e->nb[d] = p->nx[d] / p->mb[d];
    e->nblocks *= e->nb[d];
    int sym = (p->nx[d] == 1);
    e->n[d] = sym ? 1 : p->mb[d] + 2 * p->ng;
    e->is[d] = sym ? 0 : p->ng;
    e->ie[d] = sym ? 0 : p->ng + p->mb[d] - 1;
  }
  e->ncell = (size_t)e->n[0] * e->n[1] * e->n[2];
  leaf_t *leaves = NULL;
  e->multilevel = 0;
  if (p->nref > 0 || p->adaptive) {
    /* mesh_refinement.cpp:60-70: refinement needs an even number of ghost cells */
    if (p->ng % 2 != 0 || p->nref > 4 || (p->adaptive && e->ndim < 2)) {
      free(e);
      return NULL;
    }
    e->nblocks = build_leaves(e, &leaves);
    e->multilevel = 1;
    for (int d = 0; d < 3; ++d) e->nc[d] = (d < e->ndim) ? p->mb[d] / 2 + 2 * p->ng : 1;
  }
  e->blk = (block_t *)calloc(e->nblocks, sizeof(block_t));
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    block_alloc(e, B);
    if (leaves) {
      B->level = leaves[b].level;
      for (int d = 0; d < 3; ++d) B->loc[d] = leaves[b].lx[d];
      if (B->level > e->maxlevel) e->maxlevel = B->level;
    } else {
      B->loc[0] = b % e->nb[0];
      B->loc[1] = (b / e->nb[0]) % e->nb[1];
      B->loc[2] = b / (e->nb[0] * e->nb[1]);
    }
    block_geometry(e, B);
  }
  free(leaves);
  if (e->multilevel) build_neighbor_tables(e);
  e->dt = DBL_MAX;
  e->mindx = DBL_MAX;
  e->dt_hyp = DBL_MAX;
  e->new_dt = DBL_MAX;
  e->tlim = DBL_MAX;
  e->pgen_id = -1;
  return e;
}

void apko_block_loc(const apko_engine *e, int b, int *level, int *lx3) {
  *level = e->blk[b].level;
  for (int d = 0; d < 3; ++d) lx3[d] = e->blk[b].loc[d];
}

void apko_destroy(apko_engine *e) {
  if (!e) return;
  for (int b = 0; b < e->nblocks; ++b) block_free(&e->blk[b]);
  free(e->amr_last);
  free(e->blk);
  free(e);
}

int apko_num_blocks(const apko_engine *e) { return e->nblocks; }
void apko_block_dims(const apko_engine *e, int *n) { n[0] = e->n[0]; n[1] = e->n[1]; n[2] = e->n[2]; }
int apko_nvar(const apko_engine *e) { return e->nvar; }
double *apko_cons(apko_engine *e, int b) { return e->blk[b].cons; }
double *apko_prim(apko_engine *e, int b) { return e->blk[b].prim; }
double *apko_flux(apko_engine *e, int b, int dir) { return e->blk[b].flux[dir]; }
void apko_block_coords(const apko_engine *e, int b, double *xmin3, double *dx3) {
  for (int d = 0; d < 3; ++d) {
    xmin3[d] = e->blk[b].xmin[d];
    dx3[d] = e->blk[b].dx[d];
  }
}
double apko_time(const apko_engine *e) { return e->time; }
double apko_dt(const apko_engine *e) { return e->dt; }
int apko_ncycle(const apko_engine *e) { return e->ncycle; }
double apko_c_h(const apko_engine *e) { return e->c_h; }
void apko_set_c_h(apko_engine *e, double c_h) { e->c_h = c_h; }
double apko_lw_tlim(const apko_engine *e) { return e->lw.tlim_period; }
long long apko_fofc_total(const apko_engine *e) { return e->fofc_total; }

/* ===================================================================================== */
/* CalculateFluxes: src/hydro/hydro.cpp:1025-1208                                         */
/* ===================================================================================== */

/* reconstruct all variables of cell (k,j,i) along dir d -> ql (left state of face +1), qr
 * (right state of the cell's lower face); the Reconstruct<> wrappers in src/recon */
static void recon_cell(const apko_engine *e, const block_t *B, int recon, int d, int k, int j,
                       int i, double *ql, double *qr) {
  const ptrdiff_t st = (d == 0) ? 1 : (d == 1) ? e->n[0] : (ptrdiff_t)e->n[0] * e->n[1];
  for (int n = 0; n < e->nvar; ++n) {
    const double *q = &B->prim[IDX(e, n, k, j, i)];
    switch (recon) {
    case APKO_RECON_DC:
      ql[n] = qr[n] = q[0];
      break;
    case APKO_RECON_PLM:
      rc_plm(q[-st], q[0], q[st], &ql[n], &qr[n]);
      break;
    case APKO_RECON_PPM:
      rc_ppm(q[-2 * st], q[-st], q[0], q[st], q[2 * st], &ql[n], &qr[n]);
      break;
    case APKO_RECON_WENOZ:
      rc_wenoz(q[-2 * st], q[-st], q[0], q[st], q[2 * st], &ql[n], &qr[n]);
      break;
    case APKO_RECON_WENO3:
      rc_weno3(q[-st], q[0], q[st], B->dx[d] * B->dx[d], &ql[n], &qr[n]);
      break;
    case APKO_RECON_LIMO3:
      rc_limo3(q[-st], q[0], q[st], B->dx[d], (n == IDN || n == IPR), &ql[n], &qr[n]);
      break;
    }
  }
}

/* Riemann solve + passive scalar upwinding for one face, store to flux array */
static void face_flux(const apko_engine *e, block_t *B, int d, int k, int j, int i,
                      const double *wl, const double *wr) {
  double f[MAXVAR];
  riemann_face(e->p.fluid, e->p.riemann, IV1 + d, wl, wr, e->p.gamma, e->c_h, f);
  /* hydro.cpp:1089-1097: scalars upwinded by the sign of the mass flux */
  for (int n = e->nhydro; n < e->nvar; ++n) f[n] = (f[IDN] >= 0.0) ? f[IDN] * wl[n] : f[IDN] * wr[n];
  for (int n = 0; n < e->nvar; ++n) B->flux[d][IDX(e, n, k, j, i)] = f[n];
}

void apko_calc_fluxes(apko_engine *e, int recon) {
  if (e->p.riemann == APKO_RS_LLF) {
    /* CalculateFluxesTight, hydro.cpp:980-1022 (DC + LLF, cell-wise) */
    for (int b = 0; b < e->nblocks; ++b) {
      block_t *B = &e->blk[b];
      int ku = e->ie[2] + (e->ndim >= 3), ju = e->ie[1] + (e->ndim >= 2);
      for (int k = e->is[2]; k <= ku; ++k)
        for (int j = e->is[1]; j <= ju; ++j)
          for (int i = e->is[0]; i <= e->ie[0] + 1; ++i)
            for (int d = 0; d < e->ndim; ++d) {
              double wl[MAXVAR], wr[MAXVAR];
              int kk = k - (d == 2), jj = j - (d == 1), ii = i - (d == 0);
              for (int n = 0; n < e->nvar; ++n) {
                wl[n] = B->prim[IDX(e, n, kk, jj, ii)];
                wr[n] = B->prim[IDX(e, n, k, j, i)];
              }
              face_flux(e, B, d, k, j, i, wl, wr);
            }
    }
    if (e->p.diffint == 1) apko_calc_diff_fluxes(e); /* hydro.cpp:1202-1205 */
    return;
  }
  const int n1 = e->n[0];
  double *wl = (double *)malloc(sizeof(double) * MAXVAR * (n1 + 1) * 3);
  double *wr = wl + MAXVAR * (n1 + 1);
  double *wlb = wr + MAXVAR * (n1 + 1);
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    const int is = e->is[0], ie = e->ie[0], js = e->is[1], je = e->ie[1], ks = e->is[2],
              ke = e->ie[2];
    /* x1: loop limits hydro.cpp:1031-1039 (over-compute into transverse ghosts) */
    int jl = js, ju = je, kl = ks, ku = ke;
    if (e->ndim >= 2) { jl = js - 1; ju = je + 1; }
    if (e->ndim >= 3) { kl = ks - 1; ku = ke + 1; }
    for (int k = kl; k <= ku; ++k)
      for (int j = jl; j <= ju; ++j) {
        for (int i = is - 1; i <= ie + 1; ++i) /* ql(n,i+1), qr(n,i) */
          recon_cell(e, B, recon, 0, k, j, i, &wl[MAXVAR * (i + 1)], &wr[MAXVAR * i]);
        for (int i = is; i <= ie + 1; ++i)
          face_flux(e, B, 0, k, j, i, &wl[MAXVAR * i], &wr[MAXVAR * i]);
      }
    /* x2: hydro.cpp:1101-1153 */
    if (e->ndim >= 2) {
      int il = is - 1, iu = ie + 1;
      kl = ks; ku = ke;
      if (e->ndim >= 3) { kl = ks - 1; ku = ke + 1; }
      for (int k = kl; k <= ku; ++k) {
        double *pl = wl, *plb = wlb;
        for (int j = js - 1; j <= je + 1; ++j) {
          for (int i = il; i <= iu; ++i)
            recon_cell(e, B, recon, 1, k, j, i, &plb[MAXVAR * i], &wr[MAXVAR * i]);
          if (j > js - 1)
            for (int i = il; i <= iu; ++i)
              face_flux(e, B, 1, k, j, i, &pl[MAXVAR * i], &wr[MAXVAR * i]);
          double *t = pl; pl = plb; plb = t;
        }
      }
    }
    /* x3: hydro.cpp:1156-1199 */
    if (e->ndim >= 3) {
      int il = is - 1, iu = ie + 1;
      for (int j = js - 1; j <= je + 1; ++j) {
        double *pl = wl, *plb = wlb;
        for (int k = ks - 1; k <= ke + 1; ++k) {
          for (int i = il; i <= iu; ++i)
            recon_cell(e, B, recon, 2, k, j, i, &plb[MAXVAR * i], &wr[MAXVAR * i]);
          if (k > ks - 1)
            for (int i = il; i <= iu; ++i)
              face_flux(e, B, 2, k, j, i, &pl[MAXVAR * i], &wr[MAXVAR * i]);
          double *t = pl; pl = plb; plb = t;
        }
      }
    }
  }
  free(wl);
  if (e->p.diffint == 1) apko_calc_diff_fluxes(e); /* hydro.cpp:1202-1205 */
}

/* ===================================================================================== */
/* Flux divergence update: external/parthenon/src/interface/update.hpp:42-56,             */
/* update.cpp:115-141                                                                     */
/* ===================================================================================== */
static inline double flux_div(const apko_engine *e, const block_t *B, int l, int k, int j,
                              int i) {
  /* uniform_coordinates.hpp:37-40 */
  const double a1 = B->dx[1] * B->dx[2], a2 = B->dx[0] * B->dx[2], a3 = B->dx[0] * B->dx[1];
  const double vol = B->dx[0] * B->dx[1] * B->dx[2];
  double du = (a1 * B->flux[0][IDX(e, l, k, j, i + 1)] - a1 * B->flux[0][IDX(e, l, k, j, i)]);
  if (e->ndim >= 2)
    du += (a2 * B->flux[1][IDX(e, l, k, j + 1, i)] - a2 * B->flux[1][IDX(e, l, k, j, i)]);
  if (e->ndim == 3)
    du += (a3 * B->flux[2][IDX(e, l, k + 1, j, i)] - a3 * B->flux[2][IDX(e, l, k, j, i)]);
  return -du / vol;
}

void apko_flux_div_update(apko_engine *e, double g0, double g1, double bdt) {
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    for (int l = 0; l < e->nvar; ++l)
      for (int k = e->is[2]; k <= e->ie[2]; ++k)
        for (int j = e->is[1]; j <= e->ie[1]; ++j)
          for (int i = e->is[0]; i <= e->ie[0]; ++i) {
            size_t c = IDX(e, l, k, j, i);
            B->cons[c] = g0 * B->cons[c] + g1 * B->u1[c] + bdt * flux_div(e, B, l, k, j, i);
          }
  }
}

/* ===================================================================================== */
/* First-order flux correction: src/hydro/hydro.cpp:1223-1342                             */
/* ===================================================================================== */
static void llf_face(const apko_engine *e, block_t *B, double *const *fdst, int d, int k, int j,
                     int i) {
  double wl[MAXVAR], wr[MAXVAR], f[MAXVAR];
  int kk = k - (d == 2), jj = j - (d == 1), ii = i - (d == 0);
  for (int n = 0; n < e->nvar; ++n) {
    wl[n] = B->prim[IDX(e, n, kk, jj, ii)];
    wr[n] = B->prim[IDX(e, n, k, j, i)];
  }
  if (e->p.fluid == APKO_FLUID_EULER) rs_hydro_llf(IV1 + d, wl, wr, e->p.gamma, f);
  else rs_mhd_llf(IV1 + d, wl, wr, e->p.gamma, e->c_h, f);
  /* scalar fluxes: *_dc_llf.hpp "Passive scalar fluxes" */
  for (int n = e->nhydro; n < e->nvar; ++n) f[n] = (f[IDN] >= 0.0) ? f[IDN] * wl[n] : f[IDN] * wr[n];
  for (int n = 0; n < e->nvar; ++n) fdst[d][IDX(e, n, k, j, i)] = f[n];
}

long long apko_fofc(apko_engine *e, double g0, double g1, double bdt) {
  long long total = 0, num_corrected;
  int attempts = 0;
  const int nh = e->nhydro; /* NVAR = GetNVars<fluid>() : scalars are not part of the check */
  size_t nfield = e->ncell * e->nvar;
  double *snap[3] = {NULL, NULL, NULL};
  do {
    num_corrected = 0;
    for (int b = 0; b < e->nblocks; ++b) {
      block_t *B = &e->blk[b];
      double *fsrc[3] = {B->flux[0], B->flux[1], B->flux[2]};
      if (e->p.fofc_jacobi) {
        /* Jacobi variant: decisions from a snapshot of the fluxes of the previous attempt */
        for (int d = 0; d < e->ndim; ++d) {
          if (!snap[d]) snap[d] = (double *)malloc(nfield * sizeof(double));
          memcpy(snap[d], B->flux[d], nfield * sizeof(double));
          fsrc[d] = snap[d];
        }
      }
      block_t S = *B;
      for (int d = 0; d < 3; ++d) S.flux[d] = fsrc[d];
      for (int k = e->is[2]; k <= e->ie[2]; ++k)
        for (int j = e->is[1]; j <= e->ie[1]; ++j)
          for (int i = e->is[0]; i <= e->ie[0]; ++i) {
            double nc[NGLMMHD];
            for (int v = 0; v < nh; ++v) {
              size_t c = IDX(e, v, k, j, i);
              nc[v] = g0 * B->cons[c] + g1 * B->u1[c] + bdt * flux_div(e, &S, v, k, j, i);
            }
            double new_p =
                nc[IEN] - 0.5 * (SQR(nc[IM1]) + SQR(nc[IM2]) + SQR(nc[IM3])) / nc[IDN];
            if (e->p.fluid == APKO_FLUID_GLMMHD)
              new_p -= 0.5 * (SQR(nc[IB1]) + SQR(nc[IB2]) + SQR(nc[IB3]));
            if (nc[IDN] > 0.0 && new_p > 0.0) continue;
            if (attempts > 2 && nc[IDN] > 0.0 && new_p < 0.0) continue;
            llf_face(e, B, B->flux, 0, k, j, i);
            llf_face(e, B, B->flux, 0, k, j, i + 1);
            if (e->ndim >= 2) {
              llf_face(e, B, B->flux, 1, k, j, i);
              llf_face(e, B, B->flux, 1, k, j + 1, i);
            }
            if (e->ndim >= 3) {
              llf_face(e, B, B->flux, 2, k, j, i);
              llf_face(e, B, B->flux, 2, k + 1, j, i);
            }
            num_corrected += 1;
          }
    }
    total += num_corrected;
    attempts += 1;
  } while (num_corrected > 0 && attempts < 4);
  for (int d = 0; d < 3; ++d) free(snap[d]);
  return total;
}

/* ===================================================================================== */
/* Dedner source: src/hydro/glmmhd/dedner_source.cpp:17-75                                */
/* ===================================================================================== */
void apko_dedner(apko_engine *e, double bdt) {
  if (e->p.fluid != APKO_FLUID_GLMMHD) return;
  const double coeff = exp(-e->p.glmmhd_alpha * e->c_h * bdt / e->mindx);
  const int ko = (e->ndim < 3) ? 0 : 1;
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    for (int k = e->is[2]; k <= e->ie[2]; ++k)
      for (int j = e->is[1]; j <= e->ie[1]; ++j)
        for (int i = e->is[0]; i <= e->ie[0]; ++i) {
          if (e->p.dedner_extended) {
#define P(n, kk, jj, ii) B->prim[IDX(e, n, kk, jj, ii)]
            /* NOTE: 2-D assumed to have j-neighbours (ndim>=2) like the reference */
            const double divB =
                0.5 * ((P(IB1, k, j, i + 1) - P(IB1, k, j, i - 1)) / B->dx[0] +
                       (P(IB2, k, j + 1, i) - P(IB2, k, j - 1, i)) / B->dx[1] +
                       (P(IB3, k + ko, j, i) - P(IB3, k - ko, j, i)) / B->dx[2]);
            B->cons[IDX(e, IM1, k, j, i)] -= bdt * divB * P(IB1, k, j, i);
            B->cons[IDX(e, IM2, k, j, i)] -= bdt * divB * P(IB2, k, j, i);
            B->cons[IDX(e, IM3, k, j, i)] -= bdt * divB * P(IB3, k, j, i);
            B->cons[IDX(e, IEN, k, j, i)] -=
                0.5 * bdt *
                (P(IB1, k, j, i) * (P(IPS, k, j, i + 1) - P(IPS, k, j, i - 1)) / B->dx[0] +
                 P(IB2, k, j, i) * (P(IPS, k, j + 1, i) - P(IPS, k, j - 1, i)) / B->dx[1] +
                 P(IB3, k, j, i) * (P(IPS, k + ko, j, i) - P(IPS, k - ko, j, i)) / B->dx[2]);
#undef P
          }
          B->cons[IDX(e, IPS, k, j, i)] *= coeff;
        }
  }
}

/* ===================================================================================== */
/* Ghost exchange (same level) + physical BCs                                             */
/* bvals/comms/bnd_info.cpp:105-230 (index ranges), boundary_communication.cpp:48-348,    */
/* bvals/boundary_conditions.cpp:36-59, boundary_conditions_generic.hpp:174-262,          */
/* src/bvals/boundary_conditions_apk.hpp:38-85                                            */
/* ===================================================================================== */
/* GenericBC (boundary_conditions_generic.hpp:174-262) on an array of extents n[] whose interior is
 * [is, ie]: the fine block (cellbounds) or its coarse buffer (c_cellbounds, `coarse = true`) */
static void apply_bc_array(const apko_engine *e, double *arr, const int *n, const int *is,
                           const int *ie, int d, int outer, int type) {
  if (e->p.nx[d] == 1) return;
  int lo[3] = {0, 0, 0}, hi[3] = {n[0] - 1, n[1] - 1, n[2] - 1};
  const int ref = outer ? ie[d] : is[d];
  if (outer) lo[d] = ie[d] + 1;
  else hi[d] = is[d] - 1;
  const int offset = 2 * ref + (outer ? 1 : -1);
  const size_t ncell = (size_t)n[0] * n[1] * n[2];
  for (int v = 0; v < e->nvar; ++v) {
    /* AthenaPK ReflectBC flips v == DIR, i.e. the normal momentum (euler only) */
    const double sgn = (type == APKO_BC_REFLECT && v == d + 1) ? -1.0 : 1.0;
    double *a = arr + (size_t)v * ncell;
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          int c[3] = {i, j, k};
          c[d] = (type == APKO_BC_REFLECT) ? offset - c[d] : ref;
          a[((size_t)k * n[1] + j) * n[0] + i] = sgn * a[((size_t)c[2] * n[1] + c[1]) * n[0] + c[0]];
        }
  }
}

static void apply_bc(apko_engine *e, block_t *B, int d, int outer, int type) {
  apply_bc_array(e, B->cons, e->n, e->is, e->ie, d, outer, type);
}

/* physical BCs of a block of a refined mesh, in face order (boundary_conditions.cpp:36-59), on the
 * block itself or on its coarse buffer (ApplyBoundaryConditionsOnCoarseOrFine) */
static void apply_block_bcs(apko_engine *e, block_t *B, int coarse) {
  int cis[3], cie[3];
  for (int d = 0; d < 3; ++d) {
    cis[d] = (d < e->ndim) ? e->p.ng : 0;
    cie[d] = (d < e->ndim) ? e->p.ng + e->p.mb[d] / 2 - 1 : 0;
  }
  for (int face = 0; face < 2 * e->ndim; ++face) {
    const int d = face / 2, outer = face % 2;
    if (e->p.bc[face] == APKO_BC_PERIODIC) continue;
    const int last = (e->nb[d] << B->level) - 1;
    if (outer ? (B->loc[d] != last) : (B->loc[d] != 0)) continue;
    if (coarse) apply_bc_array(e, B->coarse, e->nc, cis, cie, d, outer, e->p.bc[face]);
    else apply_bc(e, B, d, outer, e->p.bc[face]);
  }
}

/* ---- multilevel ghost exchange (static refinement, periodic domains) ---------------------------
 * What parthenon::AddBoundaryExchangeTasks does on a multilevel mesh
 * (boundary_communication.cpp:405-445, index ranges bnd_info.cpp:105-230, :354-445), stated per
 * ghost region of a block (27 neighbour offsets):
 *   same-level neighbour : copy its interior cells;
 *   finer neighbours     : each abutting child restricts (RestrictAverage, pr_ops.hpp:105-165) its
 *                          interior into ITS coarse buffer and those cells become our ghost cells;
 *   coarser neighbour    : our coarse buffer receives its interior cells (ng deep); every other
 *                          part of our coarse buffer near the boundary is the restriction of OUR
 *                          cells -- interior and the ghost cells filled above (ProResInfo::GetSet
 *                          restricts the received same-level / finer ghosts) -- and the ghost cells
 *                          are then prolongated from the coarse buffer
 *                          (ProlongateCellMinModMultiD, custom_ops.hpp:49-186) over the ng/2 coarse
 *                          layers that have fine ghost cells. */
static void ml_range(const apko_engine *e, int d, int o, int coarse, int depth, int *lo, int *hi) {
  if (d >= e->ndim) {
    *lo = *hi = 0;
    return;
  }
  const int s = coarse ? e->p.ng : e->is[d];
  const int t = coarse ? e->p.ng + e->p.mb[d] / 2 - 1 : e->ie[d];
  if (o == 0) { *lo = s; *hi = t; }
  else if (o > 0) { *lo = t + 1; *hi = t + depth; }
  else { *lo = s - depth; *hi = s - 1; }
}

#define CIDX(e, n, k, j, i) \
  ((((size_t)(n) * (e)->nc[2] + (k)) * (e)->nc[1] + (j)) * (e)->nc[0] + (i))

static void exchange_multilevel(apko_engine *e) {
  const apko_params *p = &e->p;
  const int ng = p->ng, nv = e->nvar;
  /* (1) same level and (2) finer neighbours -> ghost cells */
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    for (int idx = 0; idx < 27; ++idx) {
      const int off[3] = {idx % 3 - 1, (idx / 3) % 3 - 1, idx / 9 - 1};
      if (B->ntype[idx] == 1) {
        const block_t *N = &e->blk[B->nidx[idx][0]];
        int lo[3], hi[3], shift[3];
        for (int d = 0; d < 3; ++d) {
          ml_range(e, d, off[d], 0, ng, &lo[d], &hi[d]);
          shift[d] = (d < e->ndim) ? -off[d] * p->mb[d] : 0;
        }
        for (int n = 0; n < nv; ++n)
          for (int k = lo[2]; k <= hi[2]; ++k)
            for (int j = lo[1]; j <= hi[1]; ++j)
              for (int i = lo[0]; i <= hi[0]; ++i)
                B->cons[IDX(e, n, k, j, i)] =
                    N->cons[IDX(e, n, k + shift[2], j + shift[1], i + shift[0])];
      } else if (B->ntype[idx] == 3) {
        for (int c = 0; c < B->nchild[idx]; ++c) {
          block_t *F = &e->blk[B->nidx[idx][c]];
          /* the child's coarse cells next to the shared face / edge / corner */
          int clo[3], chi[3];
          for (int d = 0; d < 3; ++d) {
            if (d >= e->ndim) { clo[d] = chi[d] = 0; continue; }
            const int cis = ng, cie = ng + p->mb[d] / 2 - 1;
            if (off[d] > 0) { clo[d] = cis; chi[d] = cis + ng - 1; }
            else if (off[d] < 0) { clo[d] = cie - ng + 1; chi[d] = cie; }
            else { clo[d] = cis; chi[d] = cie; }
          }
          apko_restrict_average(e->ndim, ng, 0, F->bxmin, F->dx, nv, e->n, F->cons, e->nc,
                                F->coarse, clo, chi);
          for (int n = 0; n < nv; ++n)
            for (int ck = clo[2]; ck <= chi[2]; ++ck)
              for (int cj = clo[1]; cj <= chi[1]; ++cj)
                for (int ci = clo[0]; ci <= chi[0]; ++ci) {
                  const int cc[3] = {ci, cj, ck};
                  int t[3];
                  for (int d = 0; d < 3; ++d) {
                    if (d >= e->ndim) { t[d] = 0; continue; }
                    if (off[d] > 0) t[d] = e->ie[d] + 1 + (cc[d] - clo[d]);
                    else if (off[d] < 0) t[d] = e->is[d] - ng + (cc[d] - clo[d]);
                    else t[d] = e->is[d] + (F->loc[d] & 1) * (p->mb[d] / 2) + (cc[d] - clo[d]);
                  }
                  B->cons[IDX(e, n, t[2], t[1], t[0])] = F->coarse[CIDX(e, n, ck, cj, ci)];
                }
        }
      }
    }
  }
  /* (3) coarse buffers of the blocks that have a coarser neighbour, (4) prolongation */
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    if (!B->has_coarser) continue;
    for (int idx = 0; idx < 27; ++idx) {
      const int off[3] = {idx % 3 - 1, (idx / 3) % 3 - 1, idx / 9 - 1};
      int lo[3], hi[3];
      if (idx == 13 || B->ntype[idx] == 1 || B->ntype[idx] == 3) {
        /* restriction of our own cells: interior, or ghost cells received from same-level /
         * finer neighbours (ng fine = ng/2 coarse layers) */
        for (int d = 0; d < 3; ++d) ml_range(e, d, off[d], 1, ng / 2, &lo[d], &hi[d]);
        apko_restrict_average(e->ndim, ng, 0, B->bxmin, B->dx, nv, e->n, B->cons, e->nc, B->coarse,
                              lo, hi);
      } else if (B->ntype[idx] == 2) {
        const block_t *Cn = &e->blk[B->nidx[idx][0]];
        for (int d = 0; d < 3; ++d) ml_range(e, d, off[d], 1, ng, &lo[d], &hi[d]);
        for (int n = 0; n < nv; ++n)
          for (int ck = lo[2]; ck <= hi[2]; ++ck)
            for (int cj = lo[1]; cj <= hi[1]; ++cj)
              for (int ci = lo[0]; ci <= hi[0]; ++ci) {
                const int cc[3] = {ci, cj, ck};
                int t[3];
                for (int d = 0; d < 3; ++d) {
                  if (d >= e->ndim) { t[d] = 0; continue; }
                  /* global cell index on the coarser level (periodic wrap), then local in Cn */
                  const long tot = (long)(e->nb[d] << (B->level - 1)) * p->mb[d];
                  long G = (long)B->loc[d] * (p->mb[d] / 2) + (cc[d] - ng);
                  G = ((G % tot) + tot) % tot;
                  t[d] = e->is[d] + (int)(G - (long)Cn->loc[d] * p->mb[d]);
                  if (t[d] < e->is[d] || t[d] > e->ie[d]) {
                    fprintf(stderr, "apk_oracle: coarse neighbour does not cover a buffer cell\n");
                    abort();
                  }
                }
                B->coarse[CIDX(e, n, ck, cj, ci)] = Cn->cons[IDX(e, n, t[2], t[1], t[0])];
              }
      }
    }
    /* physical boundaries of the coarse buffer (cbound, boundary_communication.cpp:438-440) */
    apply_block_bcs(e, B, 1);
    for (int idx = 0; idx < 27; ++idx) {
      if (B->ntype[idx] != 2) continue;
      const int off[3] = {idx % 3 - 1, (idx / 3) % 3 - 1, idx / 9 - 1};
      int lo[3], hi[3];
      for (int d = 0; d < 3; ++d) ml_range(e, d, off[d], 1, ng / 2, &lo[d], &hi[d]);
      apko_prolongate_minmod_multid(e->ndim, ng, B->bxmin, B->dx, nv, e->n, B->cons, e->nc,
                                    B->coarse, lo, hi);
    }
  }
  /* fbound */
  for (int b = 0; b < e->nblocks; ++b) apply_block_bcs(e, &e->blk[b], 0);
}

/* ---- flux correction (bvals_in_one.hpp:69-80, bnd_info.cpp:71-103, hydro_driver.cpp:527-531):
 * on a face shared with finer blocks the coarse block's flux is REPLACED by the area-weighted mean
 * of the fine faces (RestrictAverage on the face element of the sweep direction) */
static void flux_correction(apko_engine *e) {
  const apko_params *p = &e->p;
  const int ng = p->ng, nv = e->nvar;
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    for (int d = 0; d < e->ndim; ++d)
      for (int side = -1; side <= 1; side += 2) {
        int off[3] = {0, 0, 0};
        off[d] = side;
        const int idx = (off[2] + 1) * 9 + (off[1] + 1) * 3 + (off[0] + 1);
        if (B->ntype[idx] != 3) continue;
        for (int c = 0; c < B->nchild[idx]; ++c) {
          block_t *F = &e->blk[B->nidx[idx][c]];
          int clo[3], chi[3];
          for (int dd = 0; dd < 3; ++dd) {
            if (dd >= e->ndim) { clo[dd] = chi[dd] = 0; continue; }
            const int cis = ng, cie = ng + p->mb[dd] / 2 - 1;
            if (dd == d) clo[dd] = chi[dd] = (side > 0) ? cis : cie + 1;
            else { clo[dd] = cis; chi[dd] = cie; }
          }
          /* the child's coarse buffer is free between exchanges: scratch for the restricted face */
          apko_restrict_average(e->ndim, ng, d + 1, F->bxmin, F->dx, nv, e->n, F->flux[d], e->nc,
                                F->coarse, clo, chi);
          for (int n = 0; n < nv; ++n)
            for (int ck = clo[2]; ck <= chi[2]; ++ck)
              for (int cj = clo[1]; cj <= chi[1]; ++cj)
                for (int ci = clo[0]; ci <= chi[0]; ++ci) {
                  const int cc[3] = {ci, cj, ck};
                  int t[3];
                  for (int dd = 0; dd < 3; ++dd) {
                    if (dd >= e->ndim) { t[dd] = 0; continue; }
                    if (dd == d) t[dd] = (side > 0) ? e->ie[dd] + 1 : e->is[dd];
                    else t[dd] = e->is[dd] + (F->loc[dd] & 1) * (p->mb[dd] / 2) + (cc[dd] - ng);
                  }
                  B->flux[d][IDX(e, n, t[2], t[1], t[0])] = F->coarse[CIDX(e, n, ck, cj, ci)];
                }
        }
      }
  }
}

void apko_exchange(apko_engine *e) {
  if (e->multilevel) {
    exchange_multilevel(e);
    return;
  }
  const int ng = e->p.ng;
  /* same-level neighbour copies; reads interiors only, writes ghosts only */
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    int omax[3] = {e->p.nx[0] > 1, e->p.nx[1] > 1, e->p.nx[2] > 1};
    for (int o3 = -omax[2]; o3 <= omax[2]; ++o3)
      for (int o2 = -omax[1]; o2 <= omax[1]; ++o2)
        for (int o1 = -omax[0]; o1 <= omax[0]; ++o1) {
          if (!o1 && !o2 && !o3) continue;
          int off[3] = {o1, o2, o3}, nl[3], valid = 1;
          for (int d = 0; d < 3; ++d) {
            nl[d] = B->loc[d] + off[d];
            if (nl[d] < 0 || nl[d] >= e->nb[d]) {
              int face = 2 * d + (nl[d] < 0 ? 0 : 1);
              if (e->p.bc[face] != APKO_BC_PERIODIC) valid = 0;
              nl[d] = (nl[d] + e->nb[d]) % e->nb[d];
            }
          }
          if (!valid) continue;
          const block_t *N = &e->blk[nl[0] + e->nb[0] * (nl[1] + e->nb[1] * nl[2])];
          int dlo[3], dhi[3], shift[3];
          for (int d = 0; d < 3; ++d) {
            if (off[d] == 0) { dlo[d] = e->is[d]; dhi[d] = e->ie[d]; shift[d] = 0; }
            else if (off[d] > 0) { dlo[d] = e->ie[d] + 1; dhi[d] = e->ie[d] + ng; shift[d] = -e->p.mb[d]; }
            else { dlo[d] = e->is[d] - ng; dhi[d] = e->is[d] - 1; shift[d] = e->p.mb[d]; }
          }
          for (int n = 0; n < e->nvar; ++n)
            for (int k = dlo[2]; k <= dhi[2]; ++k)
              for (int j = dlo[1]; j <= dhi[1]; ++j)
                for (int i = dlo[0]; i <= dhi[0]; ++i)
                  B->cons[IDX(e, n, k, j, i)] =
                      N->cons[IDX(e, n, k + shift[2], j + shift[1], i + shift[0])];
        }
  }
  /* physical BCs in face order ix1, ox1, ix2, ox2, ix3, ox3 (boundary_conditions.cpp:47-56) */
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    for (int face = 0; face < 6; ++face) {
      int d = face / 2, outer = face % 2;
      if (d >= e->ndim) continue;
      if (e->p.bc[face] == APKO_BC_PERIODIC) continue;
      if (outer ? (B->loc[d] != e->nb[d] - 1) : (B->loc[d] != 0)) continue;
      apply_bc(e, B, d, outer, e->p.bc[face]);
    }
  }
}

/* ===================================================================================== */
/* cons -> prim over the ENTIRE index range: src/eos/adiabatic_glmmhd.cpp:33-56           */
/* ===================================================================================== */
void apko_cons_to_prim(apko_engine *e) {
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    for (size_t c = 0; c < e->ncell; ++c) {
      int err = c2p_cell(e->p.fluid, e->p.nscalars, B->cons + c, B->prim + c, (int)e->ncell,
                         e->p.gamma, e->p.dfloor, e->p.pfloor, e->p.efloor, e->p.vceil,
                         e->p.eceil);
      if (err) {
        fprintf(stderr, "apk_oracle: negative %s in cons2prim (block %d cell %zu)\n",
                err == 1 ? "density" : "pressure", b, c);
        abort(); /* PARTHENON_REQUIRE, adiabatic_hydro.hpp:77-79,111-113 */
      }
    }
  }
}

/* ===================================================================================== */
/* Diffusion: src/hydro/diffusion/{conduction,viscosity,resistivity}.cpp (fixed            */
/* coefficients), ResetFluxes / RKL2 super time stepping of src/hydro/hydro_driver.cpp     */
/* ===================================================================================== */
#define PR(n, k, j, i) B->prim[IDX(e, n, k, j, i)]
#define FX(d, n, k, j, i) B->flux[d][IDX(e, n, k, j, i)]

/* hydro_driver.cpp:36-91 */
void apko_reset_fluxes(apko_engine *e) {
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    for (int d = 0; d < e->ndim; ++d)
      for (int n = 0; n < e->nvar; ++n)
        for (int k = e->is[2]; k <= e->ie[2] + (d == 2); ++k)
          for (int j = e->is[1]; j <= e->ie[1] + (d == 1); ++j)
            for (int i = e->is[0]; i <= e->ie[0] + (d == 0); ++i) FX(d, n, k, j, i) = 0.0;
  }
}

static inline double temp_cell(const apko_engine *e, const block_t *B, int k, int j, int i) {
  return PR(IPR, k, j, i) / PR(IDN, k, j, i);
}

/* conduction.cpp:193-260 (ThermalFluxIsoFixed) */
static void thermal_flux_iso_fixed(apko_engine *e, block_t *B) {
  const double chi = e->p.thermal_diff_coeff;
  for (int d = 0; d < e->ndim; ++d) {
    const int o[3] = {d == 0, d == 1, d == 2};
    for (int k = e->is[2]; k <= e->ie[2] + o[2]; ++k)
      for (int j = e->is[1]; j <= e->ie[1] + o[1]; ++j)
        for (int i = e->is[0]; i <= e->ie[0] + o[0]; ++i) {
          const double T_i = temp_cell(e, B, k, j, i);
          const double T_im1 = temp_cell(e, B, k - o[2], j - o[1], i - o[0]);
          const double dTdx = (T_i - T_im1) / B->dx[d];
          const double denf = 0.5 * (PR(IDN, k, j, i) + PR(IDN, k - o[2], j - o[1], i - o[0]));
          FX(d, IEN, k, j, i) -= chi * denf * dTdx;
        }
  }
}

/* diffusion.hpp:24-66 */
static double lim_vanleer_unused(double A, double B) { return (A * B > 0) ? 2.0 * A * B / (A + B) : 0.0; }
static double lim_mc(double A, double B) {
  return lim_minmod(2.0 * lim_minmod(A, B), (A + B) / 2.0);
}
static double lim4(double A, double B, double C, double D) {
  return lim_mc(lim_mc(A, B), lim_mc(C, D));
}

/* monotonized transverse temperature difference across the d-face of cell (k,j,i) along t
 * (conduction.cpp:291-316 and the x2 / x3 analogues) */
static double dT_transverse(const apko_engine *e, const block_t *B, int d, int t, int k, int j,
                            int i) {
  const int od[3] = {d == 0, d == 1, d == 2}, ot[3] = {t == 0, t == 1, t == 2};
  const int km = k - od[2], jm = j - od[1], im = i - od[0];
  return lim4(temp_cell(e, B, k + ot[2], j + ot[1], i + ot[0]) - temp_cell(e, B, k, j, i),
              temp_cell(e, B, k, j, i) - temp_cell(e, B, k - ot[2], j - ot[1], i - ot[0]),
              temp_cell(e, B, km + ot[2], jm + ot[1], im + ot[0]) - temp_cell(e, B, km, jm, im),
              temp_cell(e, B, km, jm, im) - temp_cell(e, B, km - ot[2], jm - ot[1], im - ot[0])) /
         B->dx[t];
}

/* conduction.cpp:265-550 (ThermalFluxGeneral) for conduction = anisotropic, fixed coefficient
 * (the isotropic type only reaches this function with the Spitzer coefficient); ndim >= 2 */
static void thermal_flux_general(apko_engine *e, block_t *B) {
  const double chi = e->p.thermal_diff_coeff;
  const double sat_prefac = 5.0 * e->p.conduction_sat_phi; /* hydro.cpp:604 */
  const int ndim = e->ndim;
  for (int d = 0; d < ndim; ++d) {
    const int o[3] = {d == 0, d == 1, d == 2};
    for (int k = e->is[2]; k <= e->ie[2] + o[2]; ++k)
      for (int j = e->is[1]; j <= e->ie[1] + o[1]; ++j)
        for (int i = e->is[0]; i <= e->ie[0] + o[0]; ++i) {
          const int km = k - o[2], jm = j - o[1], im = i - o[0];
          double dT[3] = {0.0, 0.0, 0.0};
          for (int t = 0; t < 3; ++t) {
            if (t == d) {
              dT[t] = (temp_cell(e, B, k, j, i) - temp_cell(e, B, km, jm, im)) / B->dx[d];
            } else if (t < 2 || ndim >= 3) {
              dT[t] = dT_transverse(e, B, d, t, k, j, i);
            }
          }
          const double denf = 0.5 * (PR(IDN, k, j, i) + PR(IDN, km, jm, im));
          const double tdf = 0.5 * (chi + chi);
          const double Bx = 0.5 * (PR(IB1, km, jm, im) + PR(IB1, k, j, i));
          const double By = 0.5 * (PR(IB2, km, jm, im) + PR(IB2, k, j, i));
          const double Bz = (d == 2 || ndim >= 3) ? 0.5 * (PR(IB3, km, jm, im) + PR(IB3, k, j, i)) : 0.0;
          double Bmag = sqrt(SQR(Bx) + SQR(By) + SQR(Bz));
          Bmag = smax(Bmag, TINY_NUMBER);
          const double bd = (d == 0 ? Bx : (d == 1 ? By : Bz)) / Bmag;
          const double bDotGradT = (Bx * dT[0] + By * dT[1] + Bz * dT[2]) / Bmag;
          const double flux_classic = -tdf * denf * bDotGradT * bd;
          const double flux_classic_mag = fabs(tdf * denf * bDotGradT);
          double flux_sat;
          if (flux_classic > 0.0) {
            flux_sat = sat_prefac * sqrt(PR(IPR, km, jm, im) / denf) * PR(IPR, km, jm, im);
          } else if (flux_classic < 0.0) {
            flux_sat = sat_prefac * sqrt(PR(IPR, k, j, i) / denf) * PR(IPR, k, j, i);
          } else {
            const double presf = 0.5 * (PR(IPR, k, j, i) + PR(IPR, km, jm, im));
            flux_sat = sat_prefac * sqrt(presf / denf) * presf;
          }
          FX(d, IEN, k, j, i) += (flux_sat / (flux_sat + flux_classic_mag)) * flux_classic;
        }
  }
}

/* viscosity.cpp:98-291 (MomentumDiffFluxIsoFixed) */
static void momentum_flux_iso_fixed(apko_engine *e, block_t *B) {
  const double nu = e->p.mom_diff_coeff;
  const int ndim = e->ndim;
  const double dx1 = B->dx[0], dx2 = B->dx[1], dx3 = B->dx[2];
  for (int k = e->is[2]; k <= e->ie[2]; ++k)
    for (int j = e->is[1]; j <= e->ie[1]; ++j)
      for (int i = e->is[0]; i <= e->ie[0] + 1; ++i) {
        double fvx = 4.0 * (PR(IV1, k, j, i) - PR(IV1, k, j, i - 1)) / (3.0 * dx1);
        double fvy = (PR(IV2, k, j, i) - PR(IV2, k, j, i - 1)) / dx1;
        double fvz = (PR(IV3, k, j, i) - PR(IV3, k, j, i - 1)) / dx1;
        if (ndim > 1) {
          fvx -= ((PR(IV2, k, j + 1, i) + PR(IV2, k, j + 1, i - 1)) -
                  (PR(IV2, k, j - 1, i) + PR(IV2, k, j - 1, i - 1))) / (6.0 * dx2);
          fvy += ((PR(IV1, k, j + 1, i) + PR(IV1, k, j + 1, i - 1)) -
                  (PR(IV1, k, j - 1, i) + PR(IV1, k, j - 1, i - 1))) / (4.0 * dx2);
        }
        if (ndim > 2) {
          fvx -= ((PR(IV3, k + 1, j, i) + PR(IV3, k + 1, j, i - 1)) -
                  (PR(IV3, k - 1, j, i) + PR(IV3, k - 1, j, i - 1))) / (6.0 * dx3);
          fvz += ((PR(IV1, k + 1, j, i) + PR(IV1, k + 1, j, i - 1)) -
                  (PR(IV1, k - 1, j, i) + PR(IV1, k - 1, j, i - 1))) / (4.0 * dx3);
        }
        const double nud = 0.5 * nu * (PR(IDN, k, j, i) + PR(IDN, k, j, i - 1));
        FX(0, IM1, k, j, i) -= nud * fvx;
        FX(0, IM2, k, j, i) -= nud * fvy;
        FX(0, IM3, k, j, i) -= nud * fvz;
        FX(0, IEN, k, j, i) -= 0.5 * nud * ((PR(IV1, k, j, i - 1) + PR(IV1, k, j, i)) * fvx +
                                            (PR(IV2, k, j, i - 1) + PR(IV2, k, j, i)) * fvy +
                                            (PR(IV3, k, j, i - 1) + PR(IV3, k, j, i)) * fvz);
      }
  if (ndim < 2) return;
  for (int k = e->is[2]; k <= e->ie[2]; ++k)
    for (int j = e->is[1]; j <= e->ie[1] + 1; ++j)
      for (int i = e->is[0]; i <= e->ie[0]; ++i) {
        double fvx = (PR(IV1, k, j, i) - PR(IV1, k, j - 1, i)) / dx2 +
                     ((PR(IV2, k, j, i + 1) + PR(IV2, k, j - 1, i + 1)) -
                      (PR(IV2, k, j, i - 1) + PR(IV2, k, j - 1, i - 1))) / (4.0 * dx1);
        double fvy = (PR(IV2, k, j, i) - PR(IV2, k, j - 1, i)) * 4.0 / (3.0 * dx2) -
                     ((PR(IV1, k, j, i + 1) + PR(IV1, k, j - 1, i + 1)) -
                      (PR(IV1, k, j, i - 1) + PR(IV1, k, j - 1, i - 1))) / (6.0 * dx1);
        double fvz = (PR(IV3, k, j, i) - PR(IV3, k, j - 1, i)) / dx2;
        if (ndim > 2) {
          fvy -= ((PR(IV3, k + 1, j, i) + PR(IV3, k + 1, j - 1, i)) -
                  (PR(IV3, k - 1, j, i) + PR(IV3, k - 1, j - 1, i))) / (6.0 * dx3);
          fvz += ((PR(IV2, k + 1, j, i) + PR(IV2, k + 1, j - 1, i)) -
                  (PR(IV2, k - 1, j, i) + PR(IV2, k - 1, j - 1, i))) / (4.0 * dx3);
        }
        const double nud = 0.5 * nu * (PR(IDN, k, j, i) + PR(IDN, k, j - 1, i));
        FX(1, IM1, k, j, i) -= nud * fvx;
        FX(1, IM2, k, j, i) -= nud * fvy;
        FX(1, IM3, k, j, i) -= nud * fvz;
        FX(1, IEN, k, j, i) -= 0.5 * nud * ((PR(IV1, k, j - 1, i) + PR(IV1, k, j, i)) * fvx +
                                            (PR(IV2, k, j - 1, i) + PR(IV2, k, j, i)) * fvy +
                                            (PR(IV3, k, j - 1, i) + PR(IV3, k, j, i)) * fvz);
      }
  if (ndim < 3) return;
  for (int k = e->is[2]; k <= e->ie[2] + 1; ++k)
    for (int j = e->is[1]; j <= e->ie[1]; ++j)
      for (int i = e->is[0]; i <= e->ie[0]; ++i) {
        const double fvx = (PR(IV1, k, j, i) - PR(IV1, k - 1, j, i)) / dx3 +
                           ((PR(IV3, k, j, i + 1) + PR(IV3, k - 1, j, i + 1)) -
                            (PR(IV3, k, j, i - 1) + PR(IV3, k - 1, j, i - 1))) / (4.0 * dx1);
        const double fvy = (PR(IV2, k, j, i) - PR(IV2, k - 1, j, i)) / dx3 +
                           ((PR(IV3, k, j + 1, i) + PR(IV3, k - 1, j + 1, i)) -
                            (PR(IV3, k, j - 1, i) + PR(IV3, k - 1, j - 1, i))) / (4.0 * dx2);
        const double fvz = (PR(IV3, k, j, i) - PR(IV3, k - 1, j, i)) * 4.0 / (3.0 * dx3) -
                           ((PR(IV1, k, j, i + 1) + PR(IV1, k - 1, j, i + 1)) -
                            (PR(IV1, k, j, i - 1) + PR(IV1, k - 1, j, i - 1))) / (6.0 * dx1) -
                           ((PR(IV2, k, j + 1, i) + PR(IV2, k - 1, j + 1, i)) -
                            (PR(IV2, k, j - 1, i) + PR(IV2, k - 1, j - 1, i))) / (6.0 * dx2);
        const double nud = 0.5 * nu * (PR(IDN, k, j, i) + PR(IDN, k - 1, j, i));
        FX(2, IM1, k, j, i) -= nud * fvx;
        FX(2, IM2, k, j, i) -= nud * fvy;
        FX(2, IM3, k, j, i) -= nud * fvz;
        FX(2, IEN, k, j, i) -= 0.5 * nud * ((PR(IV1, k - 1, j, i) + PR(IV1, k, j, i)) * fvx +
                                            (PR(IV2, k - 1, j, i) + PR(IV2, k, j, i)) * fvy +
                                            (PR(IV3, k - 1, j, i) + PR(IV3, k, j, i)) * fvz);
      }
}

/* resistivity.cpp:89-232 (OhmicDiffFluxIsoFixed).  coords.Xf<d, face != d> is the cell-centre
 * coordinate Xc<d> (uniform_coordinates.hpp:143-157): the spacing 2 dx is evaluated from positions */
static void ohmic_flux_iso_fixed(apko_engine *e, block_t *B) {
  const double eta = e->p.ohm_diff_coeff;
  const int ndim = e->ndim;
  const double dx1 = B->dx[0], dx2 = B->dx[1], dx3 = B->dx[2];
  for (int k = e->is[2]; k <= e->ie[2]; ++k)
    for (int j = e->is[1]; j <= e->ie[1]; ++j)
      for (int i = e->is[0]; i <= e->ie[0] + 1; ++i) {
        const double d3B1 =
            ndim > 2 ? (0.5 * (PR(IB1, k + 1, j, i - 1) + PR(IB1, k + 1, j, i)) -
                        0.5 * (PR(IB1, k - 1, j, i - 1) + PR(IB1, k - 1, j, i))) /
                           (xc(B, 2, k + 1) - xc(B, 2, k - 1))
                     : 0.0;
        const double d1B3 = (PR(IB3, k, j, i) - PR(IB3, k, j, i - 1)) / dx1;
        const double j2 = d3B1 - d1B3;
        const double d1B2 = (PR(IB2, k, j, i) - PR(IB2, k, j, i - 1)) / dx1;
        const double d2B1 =
            ndim > 1 ? (0.5 * (PR(IB1, k, j + 1, i - 1) + PR(IB1, k, j + 1, i)) -
                        0.5 * (PR(IB1, k, j - 1, i - 1) + PR(IB1, k, j - 1, i))) /
                           (xc(B, 1, j + 1) - xc(B, 1, j - 1))
                     : 0.0;
        const double j3 = d1B2 - d2B1;
        FX(0, IB2, k, j, i) += -eta * j3;
        FX(0, IB3, k, j, i) += eta * j2;
        FX(0, IEN, k, j, i) += 0.5 * eta * ((PR(IB3, k, j, i - 1) + PR(IB3, k, j, i)) * j2 -
                                            (PR(IB2, k, j, i - 1) + PR(IB2, k, j, i)) * j3);
      }
  if (ndim < 2) return;
  for (int k = e->is[2]; k <= e->ie[2]; ++k)
    for (int j = e->is[1]; j <= e->ie[1] + 1; ++j)
      for (int i = e->is[0]; i <= e->ie[0]; ++i) {
        const double d1B2 = (0.5 * (PR(IB2, k, j - 1, i + 1) + PR(IB2, k, j, i + 1)) -
                             0.5 * (PR(IB2, k, j - 1, i - 1) + PR(IB2, k, j, i - 1))) /
                            (xc(B, 0, i + 1) - xc(B, 0, i - 1));
        const double d2B1 = (PR(IB1, k, j, i) - PR(IB1, k, j - 1, i)) / dx2;
        const double j3 = d1B2 - d2B1;
        const double d2B3 = (PR(IB3, k, j, i) - PR(IB3, k, j - 1, i)) / dx2;
        const double d3B2 =
            ndim > 2 ? (0.5 * (PR(IB2, k + 1, j - 1, i) + PR(IB2, k + 1, j, i)) -
                        0.5 * (PR(IB2, k - 1, j - 1, i) + PR(IB2, k - 1, j, i))) /
                           (xc(B, 2, k + 1) - xc(B, 2, k - 1))
                     : 0.0;
        const double j1 = d2B3 - d3B2;
        FX(1, IB1, k, j, i) += eta * j3;
        FX(1, IB3, k, j, i) += -eta * j1;
        FX(1, IEN, k, j, i) += 0.5 * eta * ((PR(IB1, k, j - 1, i) + PR(IB1, k, j, i)) * j3 -
                                            (PR(IB3, k, j - 1, i) + PR(IB3, k, j, i)) * j1);
      }
  if (ndim < 3) return;
  for (int k = e->is[2]; k <= e->ie[2] + 1; ++k)
    for (int j = e->is[1]; j <= e->ie[1]; ++j)
      for (int i = e->is[0]; i <= e->ie[0]; ++i) {
        const double d2B3 = (0.5 * (PR(IB3, k - 1, j + 1, i) + PR(IB3, k, j + 1, i)) -
                             0.5 * (PR(IB3, k - 1, j - 1, i) + PR(IB3, k, j - 1, i))) /
                            (xc(B, 1, j + 1) - xc(B, 1, j - 1));
        const double d3B2 = (PR(IB2, k, j, i) - PR(IB2, k - 1, j, i)) / dx3;
        const double j1 = d2B3 - d3B2;
        const double d3B1 = (PR(IB1, k, j, i) - PR(IB1, k - 1, j, i)) / dx3;
        const double d1B3 = (0.5 * (PR(IB3, k - 1, j, i + 1) + PR(IB3, k, j, i + 1)) -
                             0.5 * (PR(IB3, k - 1, j, i - 1) + PR(IB3, k, j, i - 1))) /
                            (xc(B, 0, i + 1) - xc(B, 0, i - 1));
        const double j2 = d3B1 - d1B3;
        FX(2, IB1, k, j, i) += -eta * j2;
        FX(2, IB2, k, j, i) += eta * j1;
        FX(2, IEN, k, j, i) += 0.5 * eta * ((PR(IB2, k - 1, j, i) + PR(IB2, k, j, i)) * j1 -
                                            (PR(IB1, k - 1, j, i) + PR(IB1, k, j, i)) * j2);
      }
}

/* diffusion.cpp:19-53 (CalcDiffFluxes) */
void apko_calc_diff_fluxes(apko_engine *e) {
  (void)lim_vanleer_unused;
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    if (e->p.conduction == 1) thermal_flux_iso_fixed(e, B);
    if (e->p.conduction == 2) thermal_flux_general(e, B);
    if (e->p.viscosity == 1) momentum_flux_iso_fixed(e, B);
    if (e->p.resistivity == 1) ohmic_flux_iso_fixed(e, B);
  }
}

/* conduction.cpp:44-190, viscosity.cpp:34-92, resistivity.cpp:33-84 and hydro.cpp:935-946 */
double apko_estimate_dt_diff(apko_engine *e) {
  const int ndim = e->ndim;
  const double fac = (ndim == 1) ? 0.5 : (ndim == 2 ? 0.25 : 1.0 / 6.0);
  const double cfl_diff = (e->p.cfl_diff > 0.0) ? e->p.cfl_diff : e->p.cfl;
  double dt_diff = DBL_MAX;
  const double fixed[3] = {e->p.conduction == 1 ? e->p.thermal_diff_coeff : -1.0,
                           e->p.viscosity == 1 ? e->p.mom_diff_coeff : -1.0,
                           e->p.resistivity == 1 ? e->p.ohm_diff_coeff : -1.0};
  if (e->p.conduction == 2) { /* general estimate, conduction.cpp:91-186 */
    const double sat_prefac = 5.0 * e->p.conduction_sat_phi;
    const double chi = e->p.thermal_diff_coeff;
    double m = DBL_MAX;
    for (int b = 0; b < e->nblocks; ++b) {
      const block_t *B = &e->blk[b];
      for (int k = e->is[2]; k <= e->ie[2]; ++k)
        for (int j = e->is[1]; j <= e->ie[1]; ++j)
          for (int i = e->is[0]; i <= e->ie[0]; ++i) {
            const double rho = PR(IDN, k, j, i), p = PR(IPR, k, j, i);
            const double dTdx = 0.5 * (temp_cell(e, B, k, j, i + 1) - temp_cell(e, B, k, j, i - 1)) / B->dx[0];
            const double dTdy = ndim >= 2 ? 0.5 * (temp_cell(e, B, k, j + 1, i) - temp_cell(e, B, k, j - 1, i)) / B->dx[1] : 0.0;
            const double dTdz = ndim >= 3 ? 0.5 * (temp_cell(e, B, k + 1, j, i) - temp_cell(e, B, k - 1, j, i)) / B->dx[2] : 0.0;
            const double gradTmag = sqrt(SQR(dTdx) + SQR(dTdy) + SQR(dTdz));
            if (gradTmag == 0.0) continue;
            const double Bx = PR(IB1, k, j, i), By = PR(IB2, k, j, i), Bz = PR(IB3, k, j, i);
            const double Bmag = sqrt(SQR(Bx) + SQR(By) + SQR(Bz));
            if (Bmag == 0.0) continue;
            const double flux_sat = sat_prefac * sqrt(p / rho) * p;
            const double flux_classic = chi * rho * gradTmag;
            if (flux_classic / flux_sat > 100.) continue;
            const double costheta = fabs(Bx * dTdx + By * dTdy + Bz * dTdz) / (Bmag * gradTmag);
            m = fmin(m, SQR(B->dx[0]) / (chi * fabs(Bx) / Bmag * costheta + TINY_NUMBER));
            if (ndim >= 2) m = fmin(m, SQR(B->dx[1]) / (chi * fabs(By) / Bmag * costheta + TINY_NUMBER));
            if (ndim >= 3) m = fmin(m, SQR(B->dx[2]) / (chi * fabs(Bz) / Bmag * costheta + TINY_NUMBER));
          }
    }
    dt_diff = smin(dt_diff, cfl_diff * fac * m);
  }
  for (int c = 0; c < 3; ++c) {
    if (fixed[c] < 0.0) continue;
    double m = DBL_MAX;
    for (int b = 0; b < e->nblocks; ++b) {
      const block_t *B = &e->blk[b];
      m = fmin(m, SQR(B->dx[0]) / (fixed[c] + TINY_NUMBER));
      if (ndim >= 2) m = fmin(m, SQR(B->dx[1]) / (fixed[c] + TINY_NUMBER));
      if (ndim >= 3) m = fmin(m, SQR(B->dx[2]) / (fixed[c] + TINY_NUMBER));
    }
    dt_diff = smin(dt_diff, cfl_diff * fac * m);
  }
  return dt_diff;
}

int apko_last_s_rkl(const apko_engine *e) { return e->last_s_rkl; }

/* One RKL2 super time step of length tau (AddSTSTasks, hydro_driver.cpp:170-344; RKL2StepFirst
 * :93-126, RKL2StepOther :128-166; Meyer, Balsara & Aslam 2014) */
void apko_sts(apko_engine *e, double tau) {
  const double mindt_diff = e->dt_diff;
  int s_rkl = (int)(0.5 * (sqrt(9.0 + 16.0 * tau / mindt_diff) - 1.0)) + 1;
  if (s_rkl % 2 == 0) s_rkl += 1;
  e->last_s_rkl = s_rkl;
  const size_t nfield = e->ncell * e->nvar;
  /* Y0 <- base (register u1) */
  for (int b = 0; b < e->nblocks; ++b) memcpy(e->blk[b].u1, e->blk[b].cons, nfield * sizeof(double));
  apko_reset_fluxes(e);
  apko_calc_diff_fluxes(e);
  if (e->multilevel) flux_correction(e);
  const double mu_tilde_1 = 4. / 3. / ((double)s_rkl * (double)s_rkl + (double)s_rkl - 2.);
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    for (int l = 0; l < e->nvar; ++l)
      for (int k = e->is[2]; k <= e->ie[2]; ++k)
        for (int j = e->is[1]; j <= e->ie[1]; ++j)
          for (int i = e->is[0]; i <= e->ie[0]; ++i) {
            const size_t c = IDX(e, l, k, j, i);
            B->my0[c] = flux_div(e, B, l, k, j, i); /* Update::FluxDivergence, update.cpp:63-85 */
            B->cons[c] = B->u1[c] + mu_tilde_1 * tau * B->my0[c]; /* Y_1 */
            B->yjm2[c] = B->u1[c];                                /* Y_0 */
          }
  }
  apko_exchange(e);
  apko_cons_to_prim(e);
  /* Meyer+2012 eq. (16) */
  double b_j = 1. / 3., b_jm1 = 1. / 3., b_jm2 = 1. / 3.;
  const double w1 = 4. / ((double)s_rkl * (double)s_rkl + (double)s_rkl - 2.);
  for (int jj = 2; jj <= s_rkl; jj++) {
    const double j = (double)jj;
    b_j = (j * j + j - 2.0) / (2 * j * (j + 1.0));
    const double mu_j = (2.0 * j - 1.0) / j * b_j / b_jm1;
    const double nu_j = -(j - 1.0) / j * b_j / b_jm2;
    const double mu_tilde_j = mu_j * w1;
    const double gamma_tilde_j = -(1.0 - b_jm1) * mu_tilde_j;
    apko_reset_fluxes(e);
    apko_calc_diff_fluxes(e);
    if (e->multilevel) flux_correction(e);
    for (int b = 0; b < e->nblocks; ++b) {
      block_t *B = &e->blk[b];
      for (int l = 0; l < e->nvar; ++l)
        for (int k = e->is[2]; k <= e->ie[2]; ++k)
          for (int jc = e->is[1]; jc <= e->ie[1]; ++jc)
            for (int i = e->is[0]; i <= e->ie[0]; ++i) {
              const size_t c = IDX(e, l, k, jc, i);
              const double MYjm1 = flux_div(e, B, l, k, jc, i);
              const double Yj = mu_j * B->cons[c] + nu_j * B->yjm2[c] +
                                (1.0 - mu_j - nu_j) * B->u1[c] + mu_tilde_j * tau * MYjm1 +
                                gamma_tilde_j * tau * B->my0[c];
              B->yjm2[c] = B->cons[c];
              B->cons[c] = Yj;
            }
    }
    apko_exchange(e);
    apko_cons_to_prim(e);
    b_jm2 = b_jm1;
    b_jm1 = b_j;
  }
}
#undef PR
#undef FX

/* ===================================================================================== */
/* Timestep: src/hydro/hydro.cpp:828-976                                                  */
/* ===================================================================================== */
double apko_estimate_dt(apko_engine *e) {
  double min_dt = DBL_MAX;
  for (int b = 0; b < e->nblocks; ++b) {
    const block_t *B = &e->blk[b];
    for (int k = e->is[2]; k <= e->ie[2]; ++k)
      for (int j = e->is[1]; j <= e->ie[1]; ++j)
        for (int i = e->is[0]; i <= e->ie[0]; ++i) {
#define P(n) B->prim[IDX(e, n, k, j, i)]
          double lx, ly = 0, lz = 0;
          if (e->p.fluid == APKO_FLUID_EULER) {
            lx = eos_cs(e->p.gamma, P(IDN), P(IPR));
            ly = lx;
            lz = lx;
          } else {
            lx = eos_cf(e->p.gamma, P(IDN), P(IPR), P(IB1), P(IB2), P(IB3));
            if (e->ndim > 1) ly = eos_cf(e->p.gamma, P(IDN), P(IPR), P(IB2), P(IB3), P(IB1));
            if (e->ndim > 2) lz = eos_cf(e->p.gamma, P(IDN), P(IPR), P(IB3), P(IB1), P(IB2));
          }
          min_dt = fmin(min_dt, B->dx[0] / (fabs(P(IV1)) + lx));
          if (e->ndim > 1) min_dt = fmin(min_dt, B->dx[1] / (fabs(P(IV2)) + ly));
          if (e->ndim > 2) min_dt = fmin(min_dt, B->dx[2] / (fabs(P(IV3)) + lz));
#undef P
        }
  }
  /* hydro.cpp:913-923: calc_dt_hyp = false for riemann = none (hydro.cpp:354-361) */
  double dt_hyp = DBL_MAX, dt_min = DBL_MAX;
  if (e->p.riemann != APKO_RS_NONE) {
    /* hydro.cpp:903-908: dt_hyp Param keeps the raw hyperbolic estimate (MHD only) */
    if (e->p.fluid == APKO_FLUID_GLMMHD) {
      if (e->p.cfl * min_dt < e->dt_hyp) e->dt_hyp = e->p.cfl * min_dt;
    }
    dt_hyp = e->p.cfl * min_dt;
    dt_min = smin(dt_min, dt_hyp);
  }
  /* hydro.cpp:935-963 */
  if (e->p.diffint != 0) {
    const double dt_diff = apko_estimate_dt_diff(e);
    if (e->p.diffint == 1) {
      dt_min = smin(dt_min, dt_diff);
    } else {
      const double r = e->p.rkl2_max_dt_ratio;
      if (r > 0.0 && dt_hyp / dt_diff > r) dt_min = smin(dt_min, r * dt_diff);
    }
    e->dt_diff = smin(dt_diff, e->dt_diff);
  }
  return dt_min;
}

/* hydro.cpp:65-95 */
static double global_min_dx(const apko_engine *e) {
  double m = DBL_MAX;
  for (int b = 0; b < e->nblocks; ++b) {
    m = fmin(m, e->blk[b].dx[0]);
    if (e->n[1] > 1) m = fmin(m, e->blk[b].dx[1]);
    if (e->n[2] > 1) m = fmin(m, e->blk[b].dx[2]);
  }
  return m;
}

/* driver.cpp:210-242,267-269 with default dt_* parameters */
static void set_global_dt(apko_engine *e) {
  if (e->dt < 0.1 * DBL_MAX) e->dt *= 2.0;
  e->dt = smin(e->dt, e->new_dt);
  e->new_dt = DBL_MAX;
  if (e->time < e->tlim && (e->tlim - e->time) < e->dt) e->dt = e->tlim - e->time;
}

void apko_save_u1(apko_engine *e) { /* hydro_driver.cpp:474-495 */
  for (int b = 0; b < e->nblocks; ++b)
    memcpy(e->blk[b].u1, e->blk[b].cons, e->ncell * e->nvar * sizeof(double));
}

/* ===================================================================================== */
/* Adaptive mesh refinement: Refinement::Tag + MeshRefinement::SetRefinement               */
/* (amr_criteria/refinement_package.cpp:205-233, mesh/mesh_refinement.cpp:72-110),         */
/* Mesh::UpdateMeshBlockTree / RedistributeAndRefineMeshBlocks                             */
/* (mesh/mesh-amr_loadbalance.cpp:485-995), Tree::Refine / Derefine (forest/tree.cpp)      */
/* ===================================================================================== */
int apko_amr_tag(apko_engine *e) {
  const apko_params *p = &e->p;
  const int max_level = p->numlevel - 1; /* root level is 0 here */
  int nflag = 0;
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    /* CheckAllRefinement: max(derefine, package criteria) */
    int aret;
    if (p->refine_type == 1)
      aret = apko_velocity_gradient_tag(e->ndim, p->ng, e->n, B->prim, p->refine_threshold, NULL);
    else if (p->refine_type == 2)
      aret = apko_max_density_tag(e->ndim, p->ng, e->n, B->prim, p->refine_threshold,
                                  p->refine_threshold2, NULL);
    else
      aret = apko_pressure_gradient_tag(e->ndim, p->ng, e->n, B->prim, IPR, p->refine_threshold,
                                        NULL);
    if (aret == 0) B->refine_flag = 0;
    if (aret >= 0) B->deref_count = 0;
    if (aret > 0) {
      B->refine_flag = (B->level == max_level) ? 0 : 1;
    } else if (aret < 0) {
      if (B->level == 0) {
        B->refine_flag = 0;
        B->deref_count = 0;
      } else {
        B->deref_count++;
        int ec = 0;
        for (int idx = 0; idx < 27; ++idx) ec += (B->ntype[idx] == 3) ? B->nchild[idx] : 0;
        if (ec > 0) B->refine_flag = 0;
        else B->refine_flag = (B->deref_count >= p->derefine_count) ? -1 : 0;
      }
    }
    nflag += (B->refine_flag != 0);
  }
  return nflag;
}

int apko_amr_flag(const apko_engine *e, int b) { return e->blk[b].refine_flag; }
const int *apko_amr_last_input(const apko_engine *e, int *n) {
  *n = e->amr_last_n;
  return e->amr_last;
}

void apko_amr_counts(const apko_engine *e, int *nbnew, int *nbdel, int *per_level8) {
  if (nbnew) *nbnew = e->nbnew;
  if (nbdel) *nbdel = e->nbdel;
  if (per_level8) {
    for (int l = 0; l < 8; ++l) per_level8[l] = 0;
    for (int b = 0; b < e->nblocks; ++b)
      if (e->blk[b].level < 8) per_level8[e->blk[b].level]++;
  }
}

typedef struct {
  leaf_t *L;
  int n, cap;
} leafset_t;

/* neighbour location at the same level (periodic wrap); 0 when it lies outside the domain */
static int amr_neighbor_loc(const apko_engine *e, int level, const int *lx, const int *off,
                            int *nl) {
  for (int d = 0; d < 3; ++d) {
    nl[d] = lx[d] + off[d];
    if (d >= e->ndim) {
      if (off[d]) return 0;
      continue;
    }
    const int tot = e->nb[d] << level;
    if (nl[d] < 0 || nl[d] >= tot) {
      const int face = 2 * d + (nl[d] < 0 ? 0 : 1);
      if (e->p.bc[face] != APKO_BC_PERIODIC) return 0;
      nl[d] = (nl[d] + tot) % tot;
    }
  }
  return 1;
}

/* Tree::Refine with enforce_proper_nesting (tree.cpp:95-136) */
static int amr_refine(const apko_engine *e, leafset_t *S, int level, const int *lx) {
  const int q = leaf_find(S->L, S->n, level, lx);
  if (q < 0) return 0; /* can't refine a block that doesn't exist */
  const int nchild = 1 << e->ndim;
  for (int c = 0; c < nchild; ++c) {
    leaf_t ch;
    ch.level = level + 1;
    for (int d = 0; d < 3; ++d) ch.lx[d] = (d < e->ndim) ? 2 * lx[d] + ((c >> d) & 1) : 0;
    if (c == 0) {
      S->L[q] = ch;
    } else {
      if (S->n == S->cap) S->L = (leaf_t *)realloc(S->L, (S->cap *= 2) * sizeof(leaf_t));
      S->L[S->n++] = ch;
    }
  }
  int nadded = nchild - 1;
  if (level > 0) {
    /* the 2^ndim blocks on level - 1 that could abut the new blocks on level + 1 */
    int plx[3], ox[3];
    for (int d = 0; d < 3; ++d) {
      plx[d] = lx[d] >> 1;
      ox[d] = lx[d] & 1;
    }
    for (int k = -(e->ndim > 2); k < 1; ++k)
      for (int j = -(e->ndim > 1); j < 1; ++j)
        for (int i = -1; i < 1; ++i) {
          const int off[3] = {i + ox[0], (e->ndim > 1) ? j + ox[1] : 0,
                              (e->ndim > 2) ? k + ox[2] : 0};
          int nl[3];
          if (!amr_neighbor_loc(e, level - 1, plx, off, nl)) continue;
          nadded += amr_refine(e, S, level - 1, nl);
        }
  }
  return nadded;
}

/* Tree::Derefine with enforce_proper_nesting (tree.cpp:217-272): lx at `level` becomes a leaf */
static int amr_derefine(const apko_engine *e, leafset_t *S, int level, const int *lx) {
  const int nchild = 1 << e->ndim;
  int cq[8] = {0};
  for (int c = 0; c < nchild; ++c) {
    int cl[3];
    for (int d = 0; d < 3; ++d) cl[d] = (d < e->ndim) ? 2 * lx[d] + ((c >> d) & 1) : 0;
    cq[c] = leaf_find(S->L, S->n, level + 1, cl);
    if (cq[c] < 0) return 0;
    /* none of the daughter's same-level neighbours may be an internal node */
    for (int k = -(e->ndim > 2); k <= (e->ndim > 2); ++k)
      for (int j = -(e->ndim > 1); j <= (e->ndim > 1); ++j)
        for (int i = -1; i <= 1; ++i) {
          const int off[3] = {i, j, k};
          int nl[3];
          if (!amr_neighbor_loc(e, level + 1, cl, off, nl)) continue;
          if (leaf_cover(S->L, S->n, level + 1, nl, NULL) < 0) return 0;
        }
  }
  /* the parent takes the first daughter's slot, the others are removed (largest index first) */
  S->L[cq[0]].level = level;
  for (int d = 0; d < 3; ++d) S->L[cq[0]].lx[d] = lx[d];
  for (int a = 1; a < nchild; ++a)
    for (int b2 = a + 1; b2 < nchild; ++b2)
      if (cq[b2] > cq[a]) {
        int t = cq[a];
        cq[a] = cq[b2];
        cq[b2] = t;
      }
  for (int c = 1; c < nchild; ++c) {
    /* cq[0] may be the last element: keep track of it when it is moved */
    const int last = S->n - 1;
    if (cq[c] != last) {
      S->L[cq[c]] = S->L[last];
      if (cq[0] == last) cq[0] = cq[c];
    }
    S->n--;
  }
  return nchild - 1;
}

/* position of a location in Parthenon's block list: trees in Morton order of the tree location
 * (forest.cpp:108-144), leaves of a tree in Morton order (logical_location.hpp:150-154) */
static uint64_t amr_interleave(const int *x, int bits) {
  uint64_t m = 0;
  for (int b = 0; b < bits; ++b)
    for (int d = 0; d < 3; ++d) m |= (uint64_t)((x[d] >> b) & 1) << (3 * b + d);
  return m;
}
static void amr_order_key(const apko_engine *e, int level, const int *lx, uint64_t key[2]) {
  int mcp2 = 1 << 30, any_sym = 0;
  for (int d = 0; d < 3; ++d) {
    if (e->p.nx[d] == 1) {
      any_sym = 1;
      continue;
    }
    const int m = max_pow2_divisor(e->nb[d]);
    if (m < mcp2) mcp2 = m;
  }
  if (any_sym) mcp2 = 1;
  int t[3], in[3];
  const int fin = e->p.numlevel - 1; /* finest possible level */
  for (int d = 0; d < 3; ++d) {
    const int per_tree = mcp2 << level;
    t[d] = (d < e->ndim) ? lx[d] / per_tree : 0;
    in[d] = (d < e->ndim) ? (lx[d] % per_tree) << (fin - level) : 0;
  }
  key[0] = amr_interleave(t, 20);
  key[1] = amr_interleave(in, 20);
}
static int amr_key_less(const uint64_t a[2], int la, const uint64_t b[2], int lb) {
  if (a[0] != b[0]) return a[0] < b[0];
  if (a[1] != b[1]) return a[1] < b[1];
  return la < lb;
}

/* Mesh::LoadBalancingAndAdaptiveMeshRefinement (one rank) followed by what the driver does when
 * the mesh changed (EvolutionDriver::InitializeBlockTimeSteps, driver.cpp:128-131) */
int apko_amr_remesh(apko_engine *e) {
  const apko_params *p = &e->p;
  const int ng = p->ng, nv = e->nvar;
  const int nleaf = 1 << e->ndim;
  free(e->amr_last);
  e->amr_last_n = e->nblocks;
  e->amr_last = (int *)malloc((size_t)e->nblocks * 5 * sizeof(int));
  for (int b = 0; b < e->nblocks; ++b) {
    e->amr_last[5 * b] = e->blk[b].level;
    for (int d = 0; d < 3; ++d) e->amr_last[5 * b + 1 + d] = e->blk[b].loc[d];
    e->amr_last[5 * b + 4] = e->blk[b].refine_flag;
  }
  /* ---- UpdateMeshBlockTree (mesh-amr_loadbalance.cpp:485-604) */
  int tnref = 0, tnderef = 0;
  for (int b = 0; b < e->nblocks; ++b) {
    tnref += (e->blk[b].refine_flag == 1);
    tnderef += (e->blk[b].refine_flag == -1);
  }
  if (tnref == 0 && tnderef < nleaf) return 0;
  const int nold = e->nblocks;
  leaf_t *old = (leaf_t *)malloc(nold * sizeof(leaf_t));
  leafset_t S;
  S.cap = 2 * nold + 64;
  S.n = nold;
  S.L = (leaf_t *)malloc(S.cap * sizeof(leaf_t));
  for (int b = 0; b < nold; ++b) {
    old[b].level = e->blk[b].level;
    for (int d = 0; d < 3; ++d) old[b].lx[d] = e->blk[b].loc[d];
    S.L[b] = old[b];
  }
  /* newly derefined blocks: parents whose 2^ndim daughters are all flagged, in block-list order
   * of the first daughter */
  leaf_t *cl = NULL;
  uint64_t(*ckey)[2] = NULL;
  int ctnd = 0;
  if (tnderef >= nleaf) {
    cl = (leaf_t *)malloc((tnderef / nleaf + 1) * sizeof(leaf_t));
    ckey = (uint64_t(*)[2])malloc((tnderef / nleaf + 1) * sizeof(uint64_t[2]));
    for (int b = 0; b < nold; ++b) {
      const block_t *B = &e->blk[b];
      if (B->refine_flag != -1) continue;
      int lower_left = 1;
      for (int d = 0; d < e->ndim; ++d) lower_left &= !(B->loc[d] & 1);
      if (!lower_left) continue;
      int all = 1;
      for (int c = 1; c < nleaf && all; ++c) {
        int sl[3];
        for (int d = 0; d < 3; ++d) sl[d] = (d < e->ndim) ? B->loc[d] + ((c >> d) & 1) : 0;
        const int q = leaf_find(old, nold, B->level, sl);
        all = (q >= 0 && e->blk[q].refine_flag == -1);
      }
      if (!all) continue;
      cl[ctnd].level = B->level - 1;
      for (int d = 0; d < 3; ++d) cl[ctnd].lx[d] = B->loc[d] >> 1;
      amr_order_key(e, B->level, B->loc, ckey[ctnd]);
      ctnd++;
    }
    /* block-list order (insertion sort on the first daughter's key) */
    for (int a = 1; a < ctnd; ++a)
      for (int b2 = a; b2 > 0 && amr_key_less(ckey[b2], 0, ckey[b2 - 1], 0); --b2) {
        leaf_t t = cl[b2];
        cl[b2] = cl[b2 - 1];
        cl[b2 - 1] = t;
        uint64_t k0 = ckey[b2][0], k1 = ckey[b2][1];
        ckey[b2][0] = ckey[b2 - 1][0];
        ckey[b2][1] = ckey[b2 - 1][1];
        ckey[b2 - 1][0] = k0;
        ckey[b2 - 1][1] = k1;
      }
    /* "sort the lists by level": std::sort(&clderef[0], &clderef[ctnd - 1]) leaves the LAST
     * element where it is (mesh-amr_loadbalance.cpp:587-592); finer parents first */
    if (ctnd > 1)
      for (int a = 1; a < ctnd - 1; ++a)
        for (int b2 = a; b2 > 0 && cl[b2].level > cl[b2 - 1].level; --b2) {
          leaf_t t = cl[b2];
          cl[b2] = cl[b2 - 1];
          cl[b2 - 1] = t;
        }
  }
  int nnew = 0, ndel = 0;
  /* Step 1: refinement, in block-list order (the closure does not depend on the order) */
  for (int b = 0; b < nold; ++b)
    if (e->blk[b].refine_flag == 1) nnew += amr_refine(e, &S, old[b].level, old[b].lx);
  /* Step 2: derefinement */
  for (int c = 0; c < ctnd; ++c) ndel += amr_derefine(e, &S, cl[c].level, cl[c].lx);
  free(cl);
  free(ckey);
  e->nbnew += nnew;
  e->nbdel += ndel;
  if (nnew == 0 && ndel == 0) {
    free(old);
    free(S.L);
    return 0;
  }
  /* ---- RedistributeAndRefineMeshBlocks (mesh-amr_loadbalance.cpp:626-995), one rank: the new
   * block list in Parthenon's order */
  {
    uint64_t(*key)[2] = (uint64_t(*)[2])malloc(S.n * sizeof(uint64_t[2]));
    int *ord = (int *)malloc(S.n * sizeof(int));
    for (int q = 0; q < S.n; ++q) {
      amr_order_key(e, S.L[q].level, S.L[q].lx, key[q]);
      ord[q] = q;
    }
    /* shell sort of the permutation */
    for (int gap = S.n / 2; gap > 0; gap /= 2)
      for (int a = gap; a < S.n; ++a) {
        const int t = ord[a];
        int b2 = a;
        for (; b2 >= gap && amr_key_less(key[t], S.L[t].level, key[ord[b2 - gap]],
                                         S.L[ord[b2 - gap]].level);
             b2 -= gap)
          ord[b2] = ord[b2 - gap];
        ord[b2] = t;
      }
    leaf_t *sorted = (leaf_t *)malloc(S.n * sizeof(leaf_t));
    for (int q = 0; q < S.n; ++q) sorted[q] = S.L[ord[q]];
    free(S.L);
    S.L = sorted;
    free(key);
    free(ord);
  }
  block_t *oblk = e->blk;
  block_t *nblk = (block_t *)calloc(S.n, sizeof(block_t));
  char *moved = (char *)calloc(nold, 1);
  int cis[3], cie[3];
  for (int d = 0; d < 3; ++d) {
    cis[d] = (d < e->ndim) ? ng : 0;
    cie[d] = (d < e->ndim) ? ng + p->mb[d] / 2 - 1 : 0;
  }
  for (int n = 0; n < S.n; ++n) {
    block_t *B = &nblk[n];
    const leaf_t *lf = &S.L[n];
    int q = leaf_find(old, nold, lf->level, lf->lx);
    if (q >= 0) { /* same level: the block moves as it is (incl. its derefinement counter) */
      *B = oblk[q];
      moved[q] = 1;
      continue;
    }
    B->level = lf->level;
    for (int d = 0; d < 3; ++d) B->loc[d] = lf->lx[d];
    block_alloc(e, B);
    block_geometry(e, B);
    int plx[3];
    for (int d = 0; d < 3; ++d) plx[d] = lf->lx[d] >> 1;
    q = (lf->level > 0) ? leaf_find(old, nold, lf->level - 1, plx) : -1;
    if (q >= 0) {
      /* coarse to fine (TryRecvCoarseToFine :82-146): the ENTIRE coarse buffer is the octant of
       * the parent's array (ghost cells included), then the interior is prolongated */
      const block_t *P = &oblk[q];
      int s[3];
      for (int d = 0; d < 3; ++d) s[d] = (d < e->ndim && (lf->lx[d] & 1)) ? p->mb[d] / 2 : 0;
      for (int v = 0; v < nv; ++v)
        for (int k = 0; k < e->nc[2]; ++k)
          for (int j = 0; j < e->nc[1]; ++j)
            for (int i = 0; i < e->nc[0]; ++i)
              B->coarse[CIDX(e, v, k, j, i)] = P->cons[IDX(e, v, k + s[2], j + s[1], i + s[0])];
      apko_prolongate_minmod_multid(e->ndim, ng, B->bxmin, B->dx, nv, e->n, B->cons, e->nc,
                                    B->coarse, cis, cie);
    } else {
      /* fine to coarse (:710-748 restriction of the interiors into the daughters' coarse
       * buffers, TryRecvFineToCoarse :163-228 copies them into the octants) */
      for (int c = 0; c < nleaf; ++c) {
        int clx[3];
        for (int d = 0; d < 3; ++d) clx[d] = (d < e->ndim) ? 2 * lf->lx[d] + ((c >> d) & 1) : 0;
        const int f = leaf_find(old, nold, lf->level + 1, clx);
        if (f < 0) {
          fprintf(stderr, "apk_oracle: remesh lost a block\n");
          abort();
        }
        block_t *F = &oblk[f];
        apko_restrict_average(e->ndim, ng, 0, F->bxmin, F->dx, nv, e->n, F->cons, e->nc, F->coarse,
                              cis, cie);
        int s[3];
        for (int d = 0; d < 3; ++d) s[d] = (d < e->ndim && ((c >> d) & 1)) ? p->mb[d] / 2 : 0;
        for (int v = 0; v < nv; ++v)
          for (int k = cis[2]; k <= cie[2]; ++k)
            for (int j = cis[1]; j <= cie[1]; ++j)
              for (int i = cis[0]; i <= cie[0]; ++i)
                B->cons[IDX(e, v, k + s[2], j + s[1], i + s[0])] = F->coarse[CIDX(e, v, k, j, i)];
      }
    }
  }
  for (int q = 0; q < nold; ++q)
    if (!moved[q]) block_free(&oblk[q]);
  free(moved);
  free(oblk);
  free(old);
  free(S.L);
  e->blk = nblk;
  e->nblocks = S.n;
  build_neighbor_tables(e);
  /* :973-990 ghosts with real data, derived quantities */
  apko_exchange(e);
  apko_cons_to_prim(e);
  return 1;
}

void apko_initialize(apko_engine *e, double tlim) {
  e->tlim = tlim;
  e->time = 0.0;
  e->ncycle = 0;
  apko_exchange(e);
  apko_cons_to_prim(e);
  if (e->p.adaptive) {
    /* Mesh::Initialize (mesh.cpp:746-840): tag, refine, re-run the problem generator on the new
     * mesh until the number of blocks stops changing */
    if (e->pgen_id < 0) {
      fprintf(stderr, "apk_oracle: adaptive meshes are initialised through apko_pgen\n");
      abort();
    }
    for (;;) {
      apko_amr_tag(e);
      const int nb_before = e->nblocks;
      apko_amr_remesh(e);
      if (e->nblocks == nb_before) break;
      apko_pgen(e, e->pgen_id, e->pgen_npars ? e->pgen_pars : NULL, e->pgen_npars);
      apko_exchange(e);
      apko_cons_to_prim(e);
    }
  }
  /* EvolutionDriver::InitializeBlockTimeSteps + SetGlobalTimeStep (driver.cpp:194-242) */
  e->dt = DBL_MAX;
  e->dt_hyp = DBL_MAX;
  e->dt_diff = DBL_MAX;
  e->new_dt = apko_estimate_dt(e);
  set_global_dt(e);
}

/* low_storage_integrator.cpp:53-132 */
static int integrator_coeffs(int integ, double *g0, double *g1, double *beta) {
  switch (integ) {
  case APKO_INT_RK1:
    g0[0] = 0.0; g1[0] = 1.0; beta[0] = 1.0;
    return 1;
  case APKO_INT_RK2:
    g0[0] = 0.0; g1[0] = 1.0; beta[0] = 1.0;
    g0[1] = 0.5; g1[1] = 0.5; beta[1] = 0.5;
    return 2;
  case APKO_INT_VL2:
    g0[0] = 0.0; g1[0] = 1.0; beta[0] = 0.5;
    g0[1] = 0.0; g1[1] = 1.0; beta[1] = 1.0;
    return 2;
  case APKO_INT_RK3:
    g0[0] = 0.0; g1[0] = 1.0; beta[0] = 1.0;
    g0[1] = 0.25; g1[1] = 0.75; beta[1] = 0.25;
    g0[2] = 2.0 / 3.0; g1[2] = 1.0 / 3.0; beta[2] = 2.0 / 3.0;
    return 3;
  }
  return 0;
}

/* One cycle: driver.cpp:101-133 + hydro_driver.cpp:474-613 */
void apko_cycle(apko_engine *e) {
  /* PreStepMeshUserWorkInLoop, hydro.cpp:102-143 */
  if (e->p.fluid == APKO_FLUID_GLMMHD) {
    e->mindx = global_min_dx(e);
    e->c_h = e->p.cfl * e->mindx / e->dt_hyp;
  }
  double g0[3], g1[3], beta[3];
  int nstages = integrator_coeffs(e->p.integrator, g0, g1, beta);
  /* Strang-split super time stepping, first half (hydro_driver.cpp:455-458) */
  if (e->p.diffint == 2) apko_sts(e, 0.5 * e->dt);
  for (int s = 0; s < nstages; ++s) {
    if (s == 0) apko_save_u1(e);
    /* VL2 predictor uses donor cell, hydro.cpp:459-464 */
    int recon = (s == 0 && e->p.integrator == APKO_INT_VL2) ? APKO_RECON_DC : e->p.recon;
    const double bdt = beta[s] * e->dt;
    apko_calc_fluxes(e, recon);
    if (e->p.fofc) e->fofc_total += apko_fofc(e, g0[s], g1[s], bdt);
    if (e->multilevel) flux_correction(e); /* hydro_driver.cpp:527-531 */
    apko_flux_div_update(e, g0[s], g1[s], bdt);
    apko_dedner(e, bdt);
    apko_exchange(e);
    apko_cons_to_prim(e);
    if (s == nstages - 1) {
      /* second half of the super time step (hydro_driver.cpp:578-583) */
      if (e->p.diffint == 2) apko_sts(e, 0.5 * e->dt);
      /* hydro_driver.cpp:589-613: reset reduction Params, then EstimateTimestep */
      e->mindx = DBL_MAX;
      e->dt_hyp = DBL_MAX;
      e->dt_diff = DBL_MAX;
      e->new_dt = apko_estimate_dt(e);
      /* hydro_driver.cpp:662-670 */
      if (e->p.adaptive) apko_amr_tag(e);
    }
  }
  e->ncycle++;
  e->time += e->dt;
  /* driver.cpp:128-131 */
  if (e->p.adaptive && apko_amr_remesh(e)) e->new_dt = apko_estimate_dt(e);
  set_global_dt(e);
}

int apko_run(apko_engine *e, double tlim, int nlim) {
  int n0 = e->ncycle;
  if (tlim != e->tlim) {
    e->tlim = tlim;
    if (e->time < e->tlim && (e->tlim - e->time) < e->dt) e->dt = e->tlim - e->time;
  }
  /* SimTime::KeepGoing, basic_types.hpp:244 */
  while ((e->time < e->tlim) && (nlim < 0 || e->ncycle < nlim)) apko_cycle(e);
  return e->ncycle - n0;
}

/* ===================================================================================== */
/* Problem generators (input side only; src/pgen)                                         */
/* ===================================================================================== */

/* src/pgen/linear_wave.cpp:424-505 (hydro eigensystem, right eigenvectors only) */
static void lw_eigensystem(double gm1, double d, double v1, double v2, double v3, double h,
                           double ev[5], double rem[5][5]) {
  (void)d;
  double vsq = v1 * v1 + v2 * v2 + v3 * v3;
  double asq = gm1 * smax((h - 0.5 * vsq), TINY_NUMBER);
  double a = sqrt(asq);
  ev[0] = v1 - a; ev[1] = v1; ev[2] = v1; ev[3] = v1; ev[4] = v1 + a;
  rem[0][0] = 1.0; rem[1][0] = v1 - a; rem[2][0] = v2; rem[3][0] = v3; rem[4][0] = h - v1 * a;
  rem[0][1] = 0.0; rem[1][1] = 0.0; rem[2][1] = 1.0; rem[3][1] = 0.0; rem[4][1] = v2;
  rem[0][2] = 0.0; rem[1][2] = 0.0; rem[2][2] = 0.0; rem[3][2] = 1.0; rem[4][2] = v3;
  rem[0][3] = 1.0; rem[1][3] = v1; rem[2][3] = v2; rem[3][3] = v3; rem[4][3] = 0.5 * vsq;
  rem[0][4] = 1.0; rem[1][4] = v1 + a; rem[2][4] = v2; rem[3][4] = v3; rem[4][4] = h + v1 * a;
}

/* src/pgen/linear_wave_mhd.cpp:484-640: eigenvalues and RIGHT eigenvectors (columns) of adiabatic
 * MHD in the conserved variables (d, m1, m2, m3, E, b2, b3); x, y are the Roe-averaging factors
 * (0, 1 here) */
static void lw_mhd_eigensystem(double gm1, double d, double v1, double v2, double v3, double h,
                               double b1, double b2, double b3, double x, double y, double ev[7],
                               double rem[7][7]) {
  double vsq = v1 * v1 + v2 * v2 + v3 * v3;
  double btsq = b2 * b2 + b3 * b3;
  double bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq;
  double vaxsq = b1 * b1 / d;
  double hp = h - (vaxsq + btsq / d);
  double twid_asq = smax((gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x), TINY_NUMBER);
  double ct2 = bt_starsq / d;
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = sqrt(tdif * tdif + 4.0 * twid_asq * ct2);
  double cfsq = 0.5 * (tsum + cf2_cs2);
  double cf = sqrt(cfsq);
  double cssq = twid_asq * vaxsq / cfsq;
  double cs = sqrt(cssq);
  double bt = sqrt(btsq);
  double bt_star = sqrt(bt_starsq);
  double bet2, bet3;
  if (bt == 0.0) {
    bet2 = 1.0;
    bet3 = 0.0;
  } else {
    bet2 = b2 / bt;
    bet3 = b3 / bt;
  }
  double bet2_star = bet2 / sqrt(gm1 - (gm1 - 1.0) * y);
  double bet3_star = bet3 / sqrt(gm1 - (gm1 - 1.0) * y);
  double bet_starsq = bet2_star * bet2_star + bet3_star * bet3_star;
  double vbet = v2 * bet2_star + v3 * bet3_star;
  double alpha_f, alpha_s;
  if ((cfsq - cssq) == 0.0) {
    alpha_f = 1.0;
    alpha_s = 0.0;
  } else if ((twid_asq - cssq) <= 0.0) {
    alpha_f = 0.0;
    alpha_s = 1.0;
  } else if ((cfsq - twid_asq) <= 0.0) {
    alpha_f = 1.0;
    alpha_s = 0.0;
  } else {
    alpha_f = sqrt((twid_asq - cssq) / (cfsq - cssq));
    alpha_s = sqrt((cfsq - twid_asq) / (cfsq - cssq));
  }
  double sqrtd = sqrt(d);
  double isqrtd = 1.0 / sqrtd;
  double sg = SIGN(b1);
  double twid_a = sqrt(twid_asq);
  double qf = cf * alpha_f * sg;
  double qs = cs * alpha_s * sg;
  double af_prime = twid_a * alpha_f * isqrtd;
  double as_prime = twid_a * alpha_s * isqrtd;
  double afpbb = af_prime * bt_star * bet_starsq;
  double aspbb = as_prime * bt_star * bet_starsq;
  double vax = sqrt(vaxsq);
  ev[0] = v1 - cf; ev[1] = v1 - vax; ev[2] = v1 - cs; ev[3] = v1;
  ev[4] = v1 + cs; ev[5] = v1 + vax; ev[6] = v1 + cf;
  rem[0][0] = alpha_f; rem[0][1] = 0.0; rem[0][2] = alpha_s; rem[0][3] = 1.0;
  rem[0][4] = alpha_s; rem[0][5] = 0.0; rem[0][6] = alpha_f;
  rem[1][0] = alpha_f * ev[0]; rem[1][1] = 0.0; rem[1][2] = alpha_s * ev[2]; rem[1][3] = v1;
  rem[1][4] = alpha_s * ev[4]; rem[1][5] = 0.0; rem[1][6] = alpha_f * ev[6];
  double qa = alpha_f * v2, qb = alpha_s * v2, qc = qs * bet2_star, qd = qf * bet2_star;
  rem[2][0] = qa + qc; rem[2][1] = -bet3; rem[2][2] = qb - qd; rem[2][3] = v2;
  rem[2][4] = qb + qd; rem[2][5] = bet3; rem[2][6] = qa - qc;
  qa = alpha_f * v3; qb = alpha_s * v3; qc = qs * bet3_star; qd = qf * bet3_star;
  rem[3][0] = qa + qc; rem[3][1] = bet2; rem[3][2] = qb - qd; rem[3][3] = v3;
  rem[3][4] = qb + qd; rem[3][5] = -bet2; rem[3][6] = qa - qc;
  rem[4][0] = alpha_f * (hp - v1 * cf) + qs * vbet + aspbb;
  rem[4][1] = -(v2 * bet3 - v3 * bet2);
  rem[4][2] = alpha_s * (hp - v1 * cs) - qf * vbet - afpbb;
  rem[4][3] = 0.5 * vsq + (gm1 - 1.0) * x / gm1;
  rem[4][4] = alpha_s * (hp + v1 * cs) + qf * vbet - afpbb;
  rem[4][5] = -rem[4][1];
  rem[4][6] = alpha_f * (hp + v1 * cf) - qs * vbet + aspbb;
  rem[5][0] = as_prime * bet2_star; rem[5][1] = -bet3 * sg * isqrtd;
  rem[5][2] = -af_prime * bet2_star; rem[5][3] = 0.0;
  rem[5][4] = rem[5][2]; rem[5][5] = rem[5][1]; rem[5][6] = rem[5][0];
  rem[6][0] = as_prime * bet3_star; rem[6][1] = bet2 * sg * isqrtd;
  rem[6][2] = -af_prime * bet3_star; rem[6][3] = 0.0;
  rem[6][4] = rem[6][2]; rem[6][5] = rem[6][1]; rem[6][6] = rem[6][0];
}

/* src/pgen/linear_wave.cpp:69-175 (InitUserMeshData). pars = {wave_flag, amp, vflow,
 * ang_2, ang_3, ang_2_vert, ang_3_vert}; angles -999.9 => automatic */
static void lw_setup(apko_engine *e, const double *pars, int npars, int mhd_wave) {
  const apko_params *p = &e->p;
  e->lw.mhd = mhd_wave;
  e->lw.wave_flag = (npars > 0) ? (int)pars[0] : 0;
  e->lw.amp = (npars > 1) ? pars[1] : 1.0e-6;
  e->lw.vflow = (npars > 2) ? pars[2] : 0.0;
  double ang_2 = (npars > 3) ? pars[3] : -999.9;
  double ang_3 = (npars > 4) ? pars[4] : -999.9;
  int ang_2_vert = (npars > 5) ? (int)pars[5] : 0;
  int ang_3_vert = (npars > 6) ? (int)pars[6] : 0;
  e->lw.gam = p->gamma;
  e->lw.gm1 = p->gamma - 1.0;
  double x1size = p->xmax[0] - p->xmin[0];
  double x2size = p->xmax[1] - p->xmin[1];
  double x3size = p->xmax[2] - p->xmin[2];
  if (ang_3 == -999.9) ang_3 = atan(x1size / x2size);
  double sin_a3 = sin(ang_3), cos_a3 = cos(ang_3);
  if (ang_3_vert) { sin_a3 = 1.0; cos_a3 = 0.0; ang_3 = 0.5 * M_PI; }
  if (ang_2 == -999.9) ang_2 = atan(0.5 * (x1size * cos_a3 + x2size * sin_a3) / x3size);
  double sin_a2 = sin(ang_2), cos_a2 = cos(ang_2);
  if (ang_2_vert) { sin_a2 = 1.0; cos_a2 = 0.0; ang_2 = 0.5 * M_PI; }
  double x1 = x1size * cos_a2 * cos_a3;
  double x2 = x2size * cos_a2 * sin_a3;
  double x3 = x3size * sin_a2;
  const int f2 = (p->nx[1] > 1), f3 = (p->nx[2] > 1);
  double lambda = x1;
  if (f2 && ang_3 != 0.0) lambda = smin(lambda, x2);
  if (f3 && ang_2 != 0.0) lambda = smin(lambda, x3);
  if (ang_3_vert) lambda = x2;
  if (ang_2_vert) lambda = x3;
  e->lw.sin_a2 = sin_a2; e->lw.cos_a2 = cos_a2; e->lw.sin_a3 = sin_a3; e->lw.cos_a3 = cos_a3;
  e->lw.lambda = lambda;
  e->lw.k_par = 2.0 * (M_PI) / lambda;
  e->lw.d0 = 1.0;
  e->lw.u0 = e->lw.vflow;
  double v0 = 0.0, w0 = 0.0;
  e->lw.p0 = 1.0 / e->lw.gam;
  double h0 = ((e->lw.p0 / e->lw.gm1 +
                0.5 * e->lw.d0 * (e->lw.u0 * e->lw.u0 + v0 * v0 + w0 * w0)) + e->lw.p0) / e->lw.d0;
  if (mhd_wave) { /* linear_wave_mhd.cpp:143-168 */
    e->lw.bx0 = 1.0;
    e->lw.by0 = sqrt(2.0);
    e->lw.bz0 = 0.5;
    h0 += (e->lw.bx0 * e->lw.bx0 + e->lw.by0 * e->lw.by0 + e->lw.bz0 * e->lw.bz0) / e->lw.d0;
    lw_mhd_eigensystem(e->lw.gm1, e->lw.d0, e->lw.u0, v0, w0, h0, e->lw.bx0, e->lw.by0, e->lw.bz0,
                       0.0, 1.0, e->lw.ev7, e->lw.rem7);
    e->lw.tlim_period = lambda / fabs(e->lw.ev7[e->lw.wave_flag]) * 1.0;
    /* wave amplitudes of the transverse field (linear_wave_mhd.cpp:392-393) */
    e->lw.dby = e->lw.amp * e->lw.rem7[5][e->lw.wave_flag];
    e->lw.dbz = e->lw.amp * e->lw.rem7[6][e->lw.wave_flag];
    e->lw.set = 1;
    return;
  }
  lw_eigensystem(e->lw.gm1, e->lw.d0, e->lw.u0, v0, w0, h0, e->lw.ev, e->lw.rem);
  /* test=true: tlim counted in wave periods (linear_wave.cpp:166-174), tlim=1 */
  e->lw.tlim_period = lambda / fabs(e->lw.ev[e->lw.wave_flag]) * 1.0;
  e->lw.set = 1;
}

/* analytic/initial state at a cell, linear_wave.cpp:356-372 */
static void lw_state(const apko_engine *e, const block_t *B, int k, int j, int i, double u[5]) {
  const int wf = e->lw.wave_flag;
  const double amp = e->lw.amp;
  double x = e->lw.cos_a2 * (xc(B, 0, i) * e->lw.cos_a3 + xc(B, 1, j) * e->lw.sin_a3) +
             xc(B, 2, k) * e->lw.sin_a2;
  double sn = sin(e->lw.k_par * x);
  u[IDN] = e->lw.d0 + amp * sn * e->lw.rem[0][wf];
  double mx = e->lw.d0 * e->lw.vflow + amp * sn * e->lw.rem[1][wf];
  double my = amp * sn * e->lw.rem[2][wf];
  double mz = amp * sn * e->lw.rem[3][wf];
  u[IM1] = mx * e->lw.cos_a2 * e->lw.cos_a3 - my * e->lw.sin_a3 - mz * e->lw.sin_a2 * e->lw.cos_a3;
  u[IM2] = mx * e->lw.cos_a2 * e->lw.sin_a3 + my * e->lw.cos_a3 - mz * e->lw.sin_a2 * e->lw.sin_a3;
  u[IM3] = mx * e->lw.sin_a2 + mz * e->lw.cos_a2;
  u[IEN] = e->lw.p0 / e->lw.gm1 + 0.5 * e->lw.d0 * e->lw.u0 * e->lw.u0 + amp * sn * e->lw.rem[4][wf];
}

/* analytic state of the magnetised wave at a cell centre (d, m1..m3, E, b1..b3),
 * linear_wave_mhd.cpp:200-232 (the reference measures the error against these) */
static void lw_mhd_state(const apko_engine *e, const block_t *B, int k, int j, int i, double u[8]) {
  const int wf = e->lw.wave_flag;
  const double amp = e->lw.amp;
  const double cos_a2 = e->lw.cos_a2, sin_a2 = e->lw.sin_a2, cos_a3 = e->lw.cos_a3,
               sin_a3 = e->lw.sin_a3;
  double x = cos_a2 * (xc(B, 0, i) * cos_a3 + xc(B, 1, j) * sin_a3) + xc(B, 2, k) * sin_a2;
  double sn = sin(e->lw.k_par * x);
  u[IDN] = e->lw.d0 + amp * sn * e->lw.rem7[0][wf];
  double mx = e->lw.d0 * e->lw.vflow + amp * sn * e->lw.rem7[1][wf];
  double my = amp * sn * e->lw.rem7[2][wf];
  double mz = amp * sn * e->lw.rem7[3][wf];
  u[IM1] = mx * cos_a2 * cos_a3 - my * sin_a3 - mz * sin_a2 * cos_a3;
  u[IM2] = mx * cos_a2 * sin_a3 + my * cos_a3 - mz * sin_a2 * sin_a3;
  u[IM3] = mx * sin_a2 + mz * cos_a2;
  double e0 = e->lw.p0 / e->lw.gm1 + 0.5 * e->lw.d0 * e->lw.u0 * e->lw.u0 +
              amp * sn * e->lw.rem7[4][wf];
  e0 += 0.5 * (e->lw.bx0 * e->lw.bx0 + e->lw.by0 * e->lw.by0 + e->lw.bz0 * e->lw.bz0);
  u[IEN] = e0;
  double bx = e->lw.bx0;
  double by = e->lw.by0 + amp * sn * e->lw.rem7[5][wf];
  double bz = e->lw.bz0 + amp * sn * e->lw.rem7[6][wf];
  u[5] = bx * cos_a2 * cos_a3 - by * sin_a3 - bz * sin_a2 * cos_a3;
  u[6] = bx * cos_a2 * sin_a3 + by * cos_a3 - bz * sin_a2 * sin_a3;
  u[7] = bx * sin_a2 + bz * cos_a2;
}

/* vector potential of the magnetised wave, linear_wave_mhd.cpp:441-480; c = 1, 2, 3 */
static double lw_mhd_A(const apko_engine *e, int c, double x1, double x2, double x3) {
  const double cos_a2 = e->lw.cos_a2, sin_a2 = e->lw.sin_a2, cos_a3 = e->lw.cos_a3,
               sin_a3 = e->lw.sin_a3, k_par = e->lw.k_par;
  double x = x1 * cos_a2 * cos_a3 + x2 * cos_a2 * sin_a3 + x3 * sin_a2;
  double y = -x1 * sin_a3 + x2 * cos_a3;
  double Ay = e->lw.bz0 * x - (e->lw.dbz / k_par) * cos(k_par * (x));
  double Az = -e->lw.by0 * x + (e->lw.dby / k_par) * cos(k_par * (x)) + e->lw.bx0 * y;
  if (c == 1) return -Ay * sin_a3 - Az * sin_a2 * cos_a3;
  if (c == 2) return Ay * cos_a3 - Az * sin_a2 * sin_a3;
  return Az * cos_a2;
}

/* deterministic hash -> [0,1) for APKO_PGEN_RANDOM (our own synthetic generator) */
static double hash01(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  x = x ^ (x >> 31);
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

int apko_pgen(apko_engine *e, int id, const double *pars, int npars) {
  const apko_params *p = &e->p;
  const double gm1 = p->gamma - 1.0;
  const int mhd = (p->fluid == APKO_FLUID_GLMMHD);
  if (pars != e->pgen_pars) {
    e->pgen_id = id;
    e->pgen_npars = npars > 16 ? 16 : (npars < 0 ? 0 : npars);
    for (int q = 0; q < e->pgen_npars; ++q) e->pgen_pars[q] = pars[q];
  }
  if (id == APKO_PGEN_LINEAR_WAVE) lw_setup(e, pars, npars, 0);
  if (id == APKO_PGEN_LINEAR_WAVE_MHD) {
    if (!mhd) return -1;
    lw_setup(e, pars, npars, 1);
  }
  for (int b = 0; b < e->nblocks; ++b) {
    block_t *B = &e->blk[b];
    memset(B->cons, 0, e->ncell * e->nvar * sizeof(double));
    for (int k = e->is[2]; k <= e->ie[2]; ++k)
      for (int j = e->is[1]; j <= e->ie[1]; ++j)
        for (int i = e->is[0]; i <= e->ie[0]; ++i) {
#define U(n) B->cons[IDX(e, n, k, j, i)]
          switch (id) {
          case APKO_PGEN_LINEAR_WAVE: {
            double u[5];
            lw_state(e, B, k, j, i, u);
            for (int n = 0; n < 5; ++n) U(n) = u[n];
          } break;
          case APKO_PGEN_LINEAR_WAVE_MHD: { /* linear_wave_mhd.cpp:373-430 (3-D meshes) */
            double u[8];
            lw_mhd_state(e, B, k, j, i, u);
            for (int n = 0; n < 5; ++n) U(n) = u[n];
            /* cell-centred B = curl of the cell-centred vector potential, centred differences */
            const double X1 = xc(B, 0, i), X2 = xc(B, 1, j), X3 = xc(B, 2, k);
            const double X1m = xc(B, 0, i - 1), X1p = xc(B, 0, i + 1);
            const double X2m = xc(B, 1, j - 1), X2p = xc(B, 1, j + 1);
            const double X3m = xc(B, 2, k - 1), X3p = xc(B, 2, k + 1);
            U(5) = (lw_mhd_A(e, 3, X1, X2p, X3) - lw_mhd_A(e, 3, X1, X2m, X3)) / B->dx[1] / 2.0 -
                   (lw_mhd_A(e, 2, X1, X2, X3p) - lw_mhd_A(e, 2, X1, X2, X3m)) / B->dx[2] / 2.0;
            U(6) = (lw_mhd_A(e, 1, X1, X2, X3p) - lw_mhd_A(e, 1, X1, X2, X3m)) / B->dx[2] / 2.0 -
                   (lw_mhd_A(e, 3, X1p, X2, X3) - lw_mhd_A(e, 3, X1m, X2, X3)) / B->dx[0] / 2.0;
            U(7) = (lw_mhd_A(e, 2, X1p, X2, X3) - lw_mhd_A(e, 2, X1m, X2, X3)) / B->dx[0] / 2.0 -
                   (lw_mhd_A(e, 1, X1, X2p, X3) - lw_mhd_A(e, 1, X1, X2m, X3)) / B->dx[1] / 2.0;
          } break;
          case APKO_PGEN_SOD: { /* src/pgen/sod.cpp:23-50; pars = rho_l,pres_l,u_l,rho_r,pres_r,u_r,x_discont */
            double rho_l = npars > 0 ? pars[0] : 1.0, pres_l = npars > 1 ? pars[1] : 1.0;
            double u_l = npars > 2 ? pars[2] : 0.0, rho_r = npars > 3 ? pars[3] : 0.125;
            double pres_r = npars > 4 ? pars[4] : 0.1, u_r = npars > 5 ? pars[5] : 0.0;
            double x_discont = npars > 6 ? pars[6] : 0.5;
            if (xc(B, 0, i) < x_discont) {
              U(IDN) = rho_l;
              U(IM1) = rho_l * u_l;
              U(IEN) = 0.5 * rho_l * u_l * u_l + pres_l / (p->gamma - 1.0);
            } else {
              U(IDN) = rho_r;
              U(IM1) = rho_r * u_r;
              U(IEN) = 0.5 * rho_r * u_r * u_r + pres_r / (p->gamma - 1.0);
            }
          } break;
          case APKO_PGEN_ORSZAG_TANG: { /* src/pgen/orszag_tang.cpp:34-61 */
            if (!mhd) return -1;
            double B0 = 1.0 / sqrt(4.0 * M_PI);
            double d0 = 25.0 / (36.0 * M_PI);
            double v0 = 1.0;
            double p0 = 5.0 / (12.0 * M_PI);
            U(IDN) = d0;
            U(IM1) = d0 * v0 * sin(2.0 * M_PI * xc(B, 1, j));
            U(IM2) = -d0 * v0 * sin(2.0 * M_PI * xc(B, 0, i));
            U(IM3) = 0.0;
            U(IB1) = B0 * sin(2.0 * M_PI * xc(B, 1, j));
            U(IB2) = B0 * sin(4.0 * M_PI * xc(B, 0, i));
            U(IB3) = 0.0;
            U(IEN) = p0 / gm1 + 0.5 * (SQR(U(IB1)) + SQR(U(IB2)) + SQR(U(IB3)) +
                                       (SQR(U(IM1)) + SQR(U(IM2)) + SQR(U(IM3))) / U(IDN));
          } break;
          case APKO_PGEN_UNIFORM: { /* src/pgen/turbulence.cpp:241-356 with zero velocity;
                                       pars = rho0, p0, b0 */
            double rho0 = npars > 0 ? pars[0] : 1.0, p0 = npars > 1 ? pars[1] : 1.0;
            double b0 = npars > 2 ? pars[2] : 0.01;
            U(IDN) = rho0;
            U(IEN) = p0 / gm1 + 0.5 * rho0 * (SQR(0.0) + SQR(0.0) + SQR(0.0));
            if (mhd) {
              U(IB1) = b0;
              U(IEN) += 0.5 * b0 * b0;
            }
          } break;
          case APKO_PGEN_OT3D: { /* closed-form 3-D Orszag-Tang-like */
            double B0 = 1.0 / sqrt(4.0 * M_PI);
            double d0 = 25.0 / (36.0 * M_PI);
            double p0 = 5.0 / (12.0 * M_PI);
            double x = xc(B, 0, i), y = xc(B, 1, j), z = xc(B, 2, k);
            double v1 = sin(2.0 * M_PI * y), v2 = -sin(2.0 * M_PI * x), v3 = 0.2 * sin(2.0 * M_PI * z);
            U(IDN) = d0;
            U(IM1) = d0 * v1;
            U(IM2) = d0 * v2;
            U(IM3) = d0 * v3;
            double eb = 0.0;
            if (mhd) {
              U(IB1) = B0 * sin(2.0 * M_PI * y);
              U(IB2) = B0 * sin(4.0 * M_PI * x);
              U(IB3) = B0 * 0.2 * sin(2.0 * M_PI * x);
              eb = 0.5 * (SQR(U(IB1)) + SQR(U(IB2)) + SQR(U(IB3)));
            }
            U(IEN) = p0 / gm1 + 0.5 * d0 * (SQR(v1) + SQR(v2) + SQR(v3)) + eb;
          } break;
          case APKO_PGEN_DIFFUSION: { /* src/pgen/diffusion.cpp:28-158; pars = iprob, Bx, By, t0, amp,
                                         diffusion coefficient of the process under test */
            const int iprob = npars > 0 ? (int)pars[0] : 20;
            const double Bx = npars > 1 ? pars[1] : 0.0, By = npars > 2 ? pars[2] : 0.0;
            const double t0 = npars > 3 ? pars[3] : 0.5, amp = npars > 4 ? pars[4] : 1e-6;
            const double dc = npars > 5 ? pars[5] : 0.0;
            const double X = xc(B, 0, i), Y = xc(B, 1, j), Z = xc(B, 2, k);
            double eint = -1.0;
            if (!mhd && iprob != 30) return -1; /* diffusion.cpp:35-37 */
            U(IDN) = 1.0;
            if (iprob == 0) {
              U(IB1) = Bx; U(IB2) = By;
              eint = X <= 0.0 ? 10.0 : 12.0;
            } else if (iprob == 1) {
              U(IB2) = Bx; U(IB3) = By;
              eint = Y <= 0.0 ? 10.0 : 12.0;
            } else if (iprob == 2) {
              U(IB3) = Bx; U(IB1) = By;
              eint = Z <= 0.0 ? 10.0 : 12.0;
            } else if (iprob == 10) {
              U(IB1) = Bx; U(IB2) = By;
              const double eff = Bx == 0.0 ? dc : dc * Bx * Bx;
              eint = 1 + amp / sqrt(4. * M_PI * eff * t0) * exp(-(pow(X, 2.)) / (4. * eff * t0));
            } else if (iprob >= 20 && iprob <= 22) {
              const double x = iprob == 20 ? X : (iprob == 21 ? Y : Z);
              const double y = iprob == 20 ? Y : (iprob == 21 ? Z : X);
              const double r = sqrt(SQR(x) + SQR(y)), phi = atan2(y, x);
              const int b1 = iprob == 20 ? IB1 : (iprob == 21 ? IB2 : IB3);
              const int b2 = iprob == 20 ? IB2 : (iprob == 21 ? IB3 : IB1);
              U(b1) = y / r;
              U(b2) = -x / r;
              eint = (fabs(r - 0.6) < 0.1 && fabs(phi) < M_PI / 12.0) ? 12.0 : 10.0;
            } else if (iprob == 30) {
              U(IM2) = U(IDN) * amp / pow(sqrt(4. * M_PI * dc * t0), 1.0) *
                       exp(-(pow(X, 2.)) / (4. * dc * t0));
              eint = 1.0 / (p->gamma * (p->gamma - 1.0));
            } else if (iprob == 40) {
              U(IB2) = amp / pow(sqrt(4. * M_PI * dc * t0), 1.0) * exp(-(pow(X, 2.)) / (4. * dc * t0));
              eint = 1.0 / (p->gamma * (p->gamma - 1.0));
            }
            if (!(eint > 0.0)) return -1;
            U(IEN) = U(IDN) * eint + 0.5 * ((SQR(U(IM1)) + SQR(U(IM2)) + SQR(U(IM3))) / U(IDN));
            if (mhd) U(IEN) += 0.5 * (SQR(U(IB1)) + SQR(U(IB2)) + SQR(U(IB3)));
          } break;
          case APKO_PGEN_LW_IMPLODE: { /* src/pgen/lw_implode.cpp:34-75 (Liska-Wendroff implosion);
                                          pars = d_in, p_in, d_out, p_out */
            const double d_in = npars > 0 ? pars[0] : 0.125, p_in = npars > 1 ? pars[1] : 0.14;
            const double d_out = npars > 2 ? pars[2] : 1.0, p_out = npars > 3 ? pars[3] : 1.0;
            /* y0 between cell centres so that the initial state is symmetric (:47-60) */
            double y0 = 0.5 * (p->xmax[1] + p->xmin[1]);
            for (int jj = e->is[1]; jj <= e->ie[1]; jj++)
              if (xc(B, 1, jj) > y0) {
                y0 = (B->xmin[1] + jj * B->dx[1]) + 0.5 * B->dx[1];
                break;
              }
            if (xc(B, 1, j) > (y0 - xc(B, 0, i))) {
              U(IDN) = d_out;
              U(IEN) = p_out / gm1;
            } else {
              U(IDN) = d_in;
              U(IEN) = p_in / gm1;
            }
          } break;
          case APKO_PGEN_BLAST: { /* src/pgen/blast.cpp (spherical over-pressure; B=0);
                                     pars = radius, pamb, prat, damb */
            double rad = npars > 0 ? pars[0] : 0.1, pa = npars > 1 ? pars[1] : 0.1;
            double prat = npars > 2 ? pars[2] : 100.0, da = npars > 3 ? pars[3] : 1.0;
            double x = xc(B, 0, i), y = (e->ndim > 1) ? xc(B, 1, j) : 0.0,
                   z = (e->ndim > 2) ? xc(B, 2, k) : 0.0;
            /* centre of the blast: problem/blast/x1_0, x2_0, x3_0 (blast.cpp:131-133) */
            double x0 = npars > 4 ? pars[4] : 0.0, y0 = npars > 5 ? pars[5] : 0.0,
                   z0 = npars > 6 ? pars[6] : 0.0;
            double r = sqrt(SQR(x - x0) + SQR(y - y0) + SQR(z - z0));
            U(IDN) = da;
            U(IEN) = ((r < rad) ? prat * pa : pa) / gm1;
          } break;
          case APKO_PGEN_RANDOM: { /* synthetic smooth-ish random state; pars = seed, amplitude */
            uint64_t seed = npars > 0 ? (uint64_t)pars[0] : 1;
            double amp = npars > 1 ? pars[1] : 0.2;
            uint64_t gi = (uint64_t)(B->loc[0] * p->mb[0] + (i - e->is[0]));
            uint64_t gj = (uint64_t)(B->loc[1] * p->mb[1] + (j - e->is[1]));
            uint64_t gk = (uint64_t)(B->loc[2] * p->mb[2] + (k - e->is[2]));
            uint64_t cell = (gk * (uint64_t)p->nx[1] + gj) * (uint64_t)p->nx[0] + gi;
            double r[8];
            for (int q = 0; q < 8; ++q) r[q] = 2.0 * hash01(seed * 1315423911ULL + cell * 8 + q) - 1.0;
            double d = 1.0 + amp * r[0], pr = 1.0 + amp * r[1];
            double v1 = amp * r[2], v2 = amp * r[3], v3 = amp * r[4];
            U(IDN) = d;
            U(IM1) = d * v1;
            U(IM2) = d * v2;
            U(IM3) = d * v3;
            double eb = 0.0;
            if (mhd) {
              U(IB1) = 0.5 + amp * r[5];
              U(IB2) = 0.3 + amp * r[6];
              U(IB3) = -0.2 + amp * r[7];
              eb = 0.5 * (SQR(U(IB1)) + SQR(U(IB2)) + SQR(U(IB3)));
            }
            U(IEN) = pr / gm1 + 0.5 * d * (SQR(v1) + SQR(v2) + SQR(v3)) + eb;
          } break;
          default:
            return -1;
          }
          for (int n = e->nhydro; n < e->nvar; ++n)
            U(n) = U(IDN) * (0.25 + 0.5 * ((i + j + k + n) % 2));
#undef U
        }
  }
  return 0;
}

/* linear_wave::UserWorkAfterLoop, src/pgen/linear_wave.cpp:183-295 */
void apko_linear_wave_errors(apko_engine *e, double *out) {
  if (e->lw.mhd) { /* linear_wave_mhd.cpp:183-330: 8 variables (psi excluded); out[0..8] */
    double l1m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < e->nblocks; ++b) {
      const block_t *B = &e->blk[b];
      const double vol = B->dx[0] * B->dx[1] * B->dx[2];
      for (int k = e->is[2]; k <= e->ie[2]; ++k)
        for (int j = e->is[1]; j <= e->ie[1]; ++j)
          for (int i = e->is[0]; i <= e->ie[0]; ++i) {
            double u[8];
            lw_mhd_state(e, B, k, j, i, u);
            for (int n = 0; n < 8; ++n) l1m[n] += fabs(u[n] - B->cons[IDX(e, n, k, j, i)]) * vol;
          }
    }
    const double volm = (e->p.xmax[0] - e->p.xmin[0]) * (e->p.xmax[1] - e->p.xmin[1]) *
                        (e->p.xmax[2] - e->p.xmin[2]);
    double rmsm = 0.0;
    for (int n = 0; n < 8; ++n) {
      l1m[n] = l1m[n] / volm;
      rmsm += SQR(l1m[n]);
      out[1 + n] = l1m[n];
    }
    out[0] = sqrt(rmsm);
    return;
  }
  double l1[5] = {0, 0, 0, 0, 0};
  for (int b = 0; b < e->nblocks; ++b) {
    const block_t *B = &e->blk[b];
    const double vol = B->dx[0] * B->dx[1] * B->dx[2];
    for (int k = e->is[2]; k <= e->ie[2]; ++k)
      for (int j = e->is[1]; j <= e->ie[1]; ++j)
        for (int i = e->is[0]; i <= e->ie[0]; ++i) {
          double u[5];
          lw_state(e, B, k, j, i, u);
          for (int n = 0; n < 5; ++n) l1[n] += fabs(u[n] - B->cons[IDX(e, n, k, j, i)]) * vol;
        }
  }
  const double vol = (e->p.xmax[0] - e->p.xmin[0]) * (e->p.xmax[1] - e->p.xmin[1]) *
                     (e->p.xmax[2] - e->p.xmin[2]);
  double rms = 0.0;
  for (int n = 0; n < 5; ++n) {
    l1[n] = l1[n] / vol;
    rms += SQR(l1[n]);
    out[1 + n] = l1[n];
  }
  out[0] = sqrt(rms);
}

/* HydroHst, src/hydro/hydro.cpp:145-208 */
void apko_history(apko_engine *e, double *out) {
  for (int q = 0; q < 8; ++q) out[q] = 0.0;
  const int mhd = (e->p.fluid == APKO_FLUID_GLMMHD);
  for (int b = 0; b < e->nblocks; ++b) {
    const block_t *B = &e->blk[b];
    const double vol = B->dx[0] * B->dx[1] * B->dx[2];
    for (int k = e->is[2]; k <= e->ie[2]; ++k)
      for (int j = e->is[1]; j <= e->ie[1]; ++j)
        for (int i = e->is[0]; i <= e->ie[0]; ++i) {
#define C(n, kk, jj, ii) B->cons[IDX(e, n, kk, jj, ii)]
          out[0] += C(IDN, k, j, i) * vol;
          out[1] += C(IM1, k, j, i) * vol;
          out[2] += C(IM2, k, j, i) * vol;
          out[3] += C(IM3, k, j, i) * vol;
          out[4] += 0.5 / C(IDN, k, j, i) *
                    (SQR(C(IM1, k, j, i)) + SQR(C(IM2, k, j, i)) + SQR(C(IM3, k, j, i))) * vol;
          out[5] += C(IEN, k, j, i) * vol;
          if (mhd && e->ndim >= 2) {
            out[6] += 0.5 * (SQR(C(IB1, k, j, i)) + SQR(C(IB2, k, j, i)) + SQR(C(IB3, k, j, i))) * vol;
            double divb = (C(IB1, k, j, i + 1) - C(IB1, k, j, i - 1)) / B->dx[0] +
                          (C(IB2, k, j + 1, i) - C(IB2, k, j - 1, i)) / B->dx[1];
            if (e->ndim == 3) divb += (C(IB3, k + 1, j, i) - C(IB3, k - 1, j, i)) / B->dx[2];
            double abs_b = sqrt(SQR(C(IB1, k, j, i)) + SQR(C(IB2, k, j, i)) + SQR(C(IB3, k, j, i)));
            out[7] += (abs_b != 0)
                          ? 0.5 * (sqrt(SQR(B->dx[0]) + SQR(B->dx[1]) + SQR(B->dx[2]))) *
                                fabs(divb) / abs_b * vol
                          : 0;
          }
#undef C
        }
  }
}


/* ============================================================================================
 * AMR operators
 * ============================================================================================ */
/* uniform_coordinates.hpp:31-58: fine coords of a block and the coarsen=2 copy used by the
 * coarse buffers */
typedef struct {
  double dx[3], xmin[3], cdx[3], cxmin[3];
  int istart[3];
} amr_coords;

static amr_coords amr_make_coords(int ndim, int ng, const double *xmin_block, const double *dx) {
  amr_coords c;
  for (int d = 0; d < 3; ++d) {
    c.istart[d] = (d < ndim) ? ng : 0;
    c.dx[d] = dx[d];
    c.xmin[d] = xmin_block[d] - c.istart[d] * c.dx[d];
    c.cxmin[d] = c.xmin[d];
    c.cdx[d] = c.dx[d];
  }
  const int coarsen = 2;
  for (int d = 0; d < 3; ++d) c.cxmin[d] += c.istart[d] * c.cdx[d] * (1 - coarsen);
  c.cdx[0] *= coarsen;
  c.cdx[1] *= (c.istart[1] > 0 ? coarsen : 1);
  c.cdx[2] *= (c.istart[2] > 0 ? coarsen : 1);
  return c;
}

#define AMR_F(n, k, j, i) fine[(((size_t)(n)*nfine[2] + (k)) * nfine[1] + (j)) * nfine[0] + (i)]
#define AMR_C(n, k, j, i)                                                                      \
  coarse[(((size_t)(n)*ncoarse[2] + (k)) * ncoarse[1] + (j)) * ncoarse[0] + (i)]

void apko_restrict_average(int ndim, int ng, int el, const double *xmin_block, const double *dx,
                           int nv, const int *nfine, const double *fine, const int *ncoarse,
                           double *coarse, const int *lo, const int *hi) {
  const amr_coords c = amr_make_coords(ndim, ng, xmin_block, dx);
  /* pr_ops.hpp:118-123: which directions are averaged for this element */
  const int inc1 = (ndim > 0) && (el == 0 || el == 2 || el == 3);
  const int inc2 = (ndim > 1) && (el == 0 || el == 3 || el == 1);
  const int inc3 = (ndim > 2) && (el == 0 || el == 1 || el == 2);
  /* uniform_coordinates.hpp:284-307 Volume<el>: cell volume or face area */
  double v;
  if (el == 0) v = c.dx[0] * c.dx[1] * c.dx[2];
  else if (el == 1) v = c.dx[1] * c.dx[2];
  else if (el == 2) v = c.dx[0] * c.dx[2];
  else v = c.dx[0] * c.dx[1];
  for (int n = 0; n < nv; ++n)
    for (int ck = lo[2]; ck <= hi[2]; ++ck)
      for (int cj = lo[1]; cj <= hi[1]; ++cj)
        for (int ci = lo[0]; ci <= hi[0]; ++ci) {
          const int i = (ndim > 0) ? (ci - c.istart[0]) * 2 + c.istart[0] : c.istart[0];
          const int j = (ndim > 1) ? (cj - c.istart[1]) * 2 + c.istart[1] : c.istart[1];
          const int k = (ndim > 2) ? (ck - c.istart[2]) * 2 + c.istart[2] : c.istart[2];
          double vol[2][2][2], terms[2][2][2];
          for (int ok = 0; ok < 2; ++ok)
            for (int oj = 0; oj < 2; ++oj)
              for (int oi = 0; oi < 2; ++oi) vol[ok][oj][oi] = terms[ok][oj][oi] = 0;
          for (int ok = 0; ok < 1 + inc3; ++ok)
            for (int oj = 0; oj < 1 + inc2; ++oj)
              for (int oi = 0; oi < 1 + inc1; ++oi) {
                vol[ok][oj][oi] = v;
                terms[ok][oj][oi] = vol[ok][oj][oi] * AMR_F(n, k + ok, j + oj, i + oi);
              }
          /* pr_ops.hpp:154-161: off-centred quantities first (FP symmetry) */
          const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                              ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
          AMR_C(n, ck, cj, ci) =
              (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
               ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
              tvol;
        }
}

/* pr_ops.hpp:94-100 */
static double amr_grad_minmod(double fc, double fm, double fp, double dxm, double dxp) {
  const double gxm = (fc - fm) / dxm;
  const double gxp = (fp - fc) / dxp;
  return 0.5 * (SIGN(gxm) + SIGN(gxp)) * fmin(fabs(gxm), fabs(gxp));
}

/* pr_ops.hpp:75-91 for cell-centred data: X<DIM,CC>(i) = xmin + (i + 0.5) dx */
static void amr_spacings(const amr_coords *c, int d, int i, int fi, double *dxm, double *dxp,
                         double *dxfm, double *dxfp) {
  const double xm = c->cxmin[d] + ((i - 1) + 0.5) * c->cdx[d];
  const double xc = c->cxmin[d] + (i + 0.5) * c->cdx[d];
  const double xp = c->cxmin[d] + ((i + 1) + 0.5) * c->cdx[d];
  *dxm = xc - xm;
  *dxp = xp - xc;
  const double fxm = c->xmin[d] + (fi + 0.5) * c->dx[d];
  const double fxp = c->xmin[d] + ((fi + 1) + 0.5) * c->dx[d];
  *dxfm = xc - fxm;
  *dxfp = fxp - xc;
}

void apko_prolongate_minmod_multid(int ndim, int ng, const double *xmin_block, const double *dx,
                                   int nv, const int *nfine, double *fine, const int *ncoarse,
                                   const double *coarse, const int *lo, const int *hi) {
  const amr_coords c = amr_make_coords(ndim, ng, xmin_block, dx);
  for (int n = 0; n < nv; ++n)
    for (int k = lo[2]; k <= hi[2]; ++k)
      for (int j = lo[1]; j <= hi[1]; ++j)
        for (int i = lo[0]; i <= hi[0]; ++i) {
          const int fi = (ndim > 0) ? (i - c.istart[0]) * 2 + c.istart[0] : c.istart[0];
          const int fj = (ndim > 1) ? (j - c.istart[1]) * 2 + c.istart[1] : c.istart[1];
          const int fk = (ndim > 2) ? (k - c.istart[2]) * 2 + c.istart[2] : c.istart[2];
          const double fc = AMR_C(n, k, j, i);
          double dx1fm = 0, dx1fp = 0, gx1c = 0, dx2fm = 0, dx2fp = 0, gx2c = 0, dx3fm = 0,
                 dx3fp = 0, gx3c = 0;
          double dm, dp;
          amr_spacings(&c, 0, i, fi, &dm, &dp, &dx1fm, &dx1fp);
          gx1c = amr_grad_minmod(fc, AMR_C(n, k, j, i - 1), AMR_C(n, k, j, i + 1), dm, dp);
          if (ndim > 1) {
            amr_spacings(&c, 1, j, fj, &dm, &dp, &dx2fm, &dx2fp);
            gx2c = amr_grad_minmod(fc, AMR_C(n, k, j - 1, i), AMR_C(n, k, j + 1, i), dm, dp);
          }
          if (ndim > 2) {
            amr_spacings(&c, 2, k, fk, &dm, &dp, &dx3fm, &dx3fp);
            gx3c = amr_grad_minmod(fc, AMR_C(n, k - 1, j, i), AMR_C(n, k + 1, j, i), dm, dp);
          }
          /* custom_ops.hpp:121-152 */
          double dqmax = fabs(gx1c) * fmax(dx1fm, dx1fp);
          int jlim = 0, klim = 0;
          if (ndim > 1) {
            dqmax += fabs(gx2c) * fmax(dx2fm, dx2fp);
            jlim = 1;
          }
          if (ndim > 2) {
            dqmax += fabs(gx3c) * fmax(dx3fm, dx3fp);
            klim = 1;
          }
          double qmin = fc, qmax = fc;
          for (int koff = -klim; koff <= klim; koff++)
            for (int joff = -jlim; joff <= jlim; joff++)
              for (int ioff = -1; ioff <= 1; ioff++) {
                qmin = fmin(qmin, AMR_C(n, k + koff, j + joff, i + ioff));
                qmax = fmax(qmax, AMR_C(n, k + koff, j + joff, i + ioff));
              }
          double alpha = 1.0;
          if (dqmax * alpha > (qmax - fc)) alpha = (qmax - fc) / dqmax;
          if (dqmax * alpha > (fc - qmin)) alpha = (fc - qmin) / dqmax;
          gx1c *= alpha;
          gx2c *= alpha;
          gx3c *= alpha;
          /* custom_ops.hpp:159-181 */
          AMR_F(n, fk, fj, fi) = fc - (gx1c * dx1fm + gx2c * dx2fm + gx3c * dx3fm);
          AMR_F(n, fk, fj, fi + 1) = fc + (gx1c * dx1fp - gx2c * dx2fm - gx3c * dx3fm);
          if (ndim > 1) {
            AMR_F(n, fk, fj + 1, fi) = fc - (gx1c * dx1fm - gx2c * dx2fp + gx3c * dx3fm);
            AMR_F(n, fk, fj + 1, fi + 1) = fc + (gx1c * dx1fp + gx2c * dx2fp - gx3c * dx3fm);
          }
          if (ndim > 2) {
            AMR_F(n, fk + 1, fj, fi) = fc - (gx1c * dx1fm + gx2c * dx2fm - gx3c * dx3fp);
            AMR_F(n, fk + 1, fj, fi + 1) = fc + (gx1c * dx1fp - gx2c * dx2fm + gx3c * dx3fp);
            AMR_F(n, fk + 1, fj + 1, fi) = fc - (gx1c * dx1fm - gx2c * dx2fp - gx3c * dx3fp);
            AMR_F(n, fk + 1, fj + 1, fi + 1) = fc + (gx1c * dx1fp + gx2c * dx2fp + gx3c * dx3fp);
          }
        }
}
#undef AMR_F
#undef AMR_C

int apko_pressure_gradient_tag(int ndim, int ng, const int *n, const double *prim, int ipr,
                               double threshold, double *maxeps_out) {
  /* gradient.cpp:18-61: loops over interior +- 1 */
  if (ndim < 2) {
    if (maxeps_out) *maxeps_out = 0.0;
    return 0;
  }
  const size_t n1 = n[0], n2 = n[1], ncell = (size_t)n[0] * n[1] * n[2];
  const double *w = prim + (size_t)ipr * ncell;
#define W(k, j, i) w[((size_t)(k)*n2 + (j)) * n1 + (i)]
  const int is = ng, ie = n[0] - ng - 1, js = ng, je = n[1] - ng - 1;
  const int ks = (ndim == 3) ? ng : 0, ke = (ndim == 3) ? n[2] - ng - 1 : 0;
  double maxeps = 0.0;
  if (ndim == 3) {
    for (int k = ks - 1; k <= ke + 1; ++k)
      for (int j = js - 1; j <= je + 1; ++j)
        for (int i = is - 1; i <= ie + 1; ++i) {
          const double eps = sqrt(SQR(0.5 * (W(k, j, i + 1) - W(k, j, i - 1))) +
                                  SQR(0.5 * (W(k, j + 1, i) - W(k, j - 1, i))) +
                                  SQR(0.5 * (W(k + 1, j, i) - W(k - 1, j, i)))) /
                             W(k, j, i);
          maxeps = fmax(maxeps, eps);
        }
  } else {
    const int k = ks;
    for (int j = js - 1; j <= je + 1; ++j)
      for (int i = is - 1; i <= ie + 1; ++i) {
        const double eps = sqrt(SQR(0.5 * (W(k, j, i + 1) - W(k, j, i - 1))) +
                                SQR(0.5 * (W(k, j + 1, i) - W(k, j - 1, i)))) /
                           W(k, j, i);
        maxeps = fmax(maxeps, eps);
      }
  }
#undef W
  if (maxeps_out) *maxeps_out = maxeps;
  if (maxeps > threshold) return 1;
  if (maxeps < 0.25 * threshold) return -1;
  return 0;
}

int apko_velocity_gradient_tag(int ndim, int ng, const int *n, const double *prim, double threshold,
                               double *vgmax_out) {
  /* gradient.cpp:64-95: k over the interior, j and i over interior +- 1 */
  const size_t n1 = n[0], n2 = n[1], ncell = (size_t)n[0] * n[1] * n[2];
  const double *v1 = prim + (size_t)IV1 * ncell, *v2 = prim + (size_t)IV2 * ncell;
#define AT(a, k, j, i) a[((size_t)(k)*n2 + (j)) * n1 + (i)]
  const int is = ng, ie = n[0] - ng - 1, js = ng, je = n[1] - ng - 1;
  const int ks = (ndim == 3) ? ng : 0, ke = (ndim == 3) ? n[2] - ng - 1 : 0;
  double vgmax = 0.0;
  for (int k = ks; k <= ke; ++k)
    for (int j = js - 1; j <= je + 1; ++j)
      for (int i = is - 1; i <= ie + 1; ++i) {
        const double vgy = fabs(AT(v2, k, j, i + 1) - AT(v2, k, j, i - 1)) * 0.5;
        const double vgx = fabs(AT(v1, k, j + 1, i) - AT(v1, k, j - 1, i)) * 0.5;
        const double vg = sqrt(vgx * vgx + vgy * vgy);
        if (vg > vgmax) vgmax = vg;
      }
#undef AT
  if (vgmax_out) *vgmax_out = vgmax;
  if (vgmax > threshold) return 1;
  if (vgmax < 0.5 * threshold) return -1;
  return 0;
}

int apko_max_density_tag(int ndim, int ng, const int *n, const double *prim, double refine_above,
                         double deref_below, double *maxrho_out) {
  /* other.cpp:18-44: note the x1 range ib.s .. ib.e + 1 */
  const size_t n1 = n[0], n2 = n[1];
  const int is = ng, ie = n[0] - ng - 1;
  const int js = (ndim >= 2) ? ng : 0, je = (ndim >= 2) ? n[1] - ng - 1 : 0;
  const int ks = (ndim == 3) ? ng : 0, ke = (ndim == 3) ? n[2] - ng - 1 : 0;
  double maxrho = 0.0;
  for (int k = ks; k <= ke; ++k)
    for (int j = js; j <= je; ++j)
      for (int i = is; i <= ie + 1; ++i) maxrho = fmax(maxrho, prim[((size_t)k * n2 + j) * n1 + i]);
  if (maxrho_out) *maxrho_out = maxrho;
  if (maxrho > refine_above) return 1;
  if (maxrho < deref_below) return -1;
  return 0;
}
