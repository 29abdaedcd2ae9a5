// pgen.cpp -- input generators for the BASELINE configs (host side, out of the hot path).
// Re-statement of the initial conditions of src/pgen/{linear_wave,sod,orszag_tang,turbulence}.cpp
// (evaluated at Parthenon's cell centres, uniform_coordinates.hpp:110-114) plus two synthetic
// closed-form states of our own (ot3d: SURVEY.md 8(d) item 4b; random: hashed noise for tests).
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "hydro_sim.hpp"

namespace apk_host {
namespace {
enum { IDN = 0, IM1 = 1, IM2 = 2, IM3 = 3, IEN = 4, IB1 = 5, IB2 = 6, IB3 = 7, IPS = 8 };
inline double SQR(double x) { return x * x; }
inline double xc(const BlockInfo &B, int d, int idx) { return B.xmin[d] + (idx + 0.5) * B.dx[d]; }

double hash01(std::uint64_t x) {  // splitmix64 -> [0,1)
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  x = x ^ (x >> 31);
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
}  // namespace

namespace {
// linear_wave_mhd.cpp:484-640: eigenvalues and RIGHT eigenvectors (columns) of adiabatic MHD in
// (d, m1, m2, m3, E, b2, b3); x, y: Roe-averaging factors (0 and 1 for a single state)
void MhdEigensystem(double gm1, double d, double v1, double v2, double v3, double h, double b1,
                    double b2, double b3, double x, double y, double ev[7], double rem[7][7]) {
  const double vsq = v1 * v1 + v2 * v2 + v3 * v3;
  const double btsq = b2 * b2 + b3 * b3;
  const double bt_starsq = (gm1 - (gm1 - 1.0) * y) * btsq;
  const double vaxsq = b1 * b1 / d;
  const double hp = h - (vaxsq + btsq / d);
  const double twid_asq = std::max((gm1 * (hp - 0.5 * vsq) - (gm1 - 1.0) * x), 1.0e-20);
  const double ct2 = bt_starsq / d;
  const double tsum = vaxsq + ct2 + twid_asq;
  const double tdif = vaxsq + ct2 - twid_asq;
  const double cf2_cs2 = std::sqrt(tdif * tdif + 4.0 * twid_asq * ct2);
  const double cfsq = 0.5 * (tsum + cf2_cs2);
  const double cf = std::sqrt(cfsq);
  const double cssq = twid_asq * vaxsq / cfsq;
  const double cs = std::sqrt(cssq);
  const double bt = std::sqrt(btsq);
  const double bt_star = std::sqrt(bt_starsq);
  double bet2 = 1.0, bet3 = 0.0;
  if (bt != 0.0) {
    bet2 = b2 / bt;
    bet3 = b3 / bt;
  }
  const double bet2_star = bet2 / std::sqrt(gm1 - (gm1 - 1.0) * y);
  const double bet3_star = bet3 / std::sqrt(gm1 - (gm1 - 1.0) * y);
  const double bet_starsq = bet2_star * bet2_star + bet3_star * bet3_star;
  const double vbet = v2 * bet2_star + v3 * bet3_star;
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
    alpha_f = std::sqrt((twid_asq - cssq) / (cfsq - cssq));
    alpha_s = std::sqrt((cfsq - twid_asq) / (cfsq - cssq));
  }
  const double sqrtd = std::sqrt(d);
  const double isqrtd = 1.0 / sqrtd;
  const double sg = (b1 < 0.0) ? -1.0 : 1.0;
  const double twid_a = std::sqrt(twid_asq);
  const double qf = cf * alpha_f * sg;
  const double qs = cs * alpha_s * sg;
  const double af_prime = twid_a * alpha_f * isqrtd;
  const double as_prime = twid_a * alpha_s * isqrtd;
  const double afpbb = af_prime * bt_star * bet_starsq;
  const double aspbb = as_prime * bt_star * bet_starsq;
  const double vax = std::sqrt(vaxsq);
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
}  // namespace

// linear_wave.cpp:69-175 (InitUserMeshData) + :424-505 (hydro eigensystem, right vectors);
// mhd_wave: linear_wave_mhd.cpp:68-177
void HydroSim::LinearWaveSetup(const std::vector<double> &pars, bool mhd_wave) {
  lw_.mhd = mhd_wave;
  auto par = [&](size_t i, double def) { return pars.size() > i ? pars[i] : def; };
  lw_.wave_flag = (int)par(0, 0);
  lw_.amp = par(1, 1.0e-6);
  lw_.vflow = par(2, 0.0);
  double ang_2 = par(3, -999.9), ang_3 = par(4, -999.9);
  const bool ang_2_vert = par(5, 0) != 0, ang_3_vert = par(6, 0) != 0;
  const double gam = cfg_.gamma;
  lw_.gm1 = gam - 1.0;
  const double x1size = cfg_.xmax[0] - cfg_.xmin[0], x2size = cfg_.xmax[1] - cfg_.xmin[1],
               x3size = cfg_.xmax[2] - cfg_.xmin[2];
  if (ang_3 == -999.9) ang_3 = std::atan(x1size / x2size);
  double sin_a3 = std::sin(ang_3), cos_a3 = std::cos(ang_3);
  if (ang_3_vert) { sin_a3 = 1.0; cos_a3 = 0.0; ang_3 = 0.5 * M_PI; }
  if (ang_2 == -999.9) ang_2 = std::atan(0.5 * (x1size * cos_a3 + x2size * sin_a3) / x3size);
  double sin_a2 = std::sin(ang_2), cos_a2 = std::cos(ang_2);
  if (ang_2_vert) { sin_a2 = 1.0; cos_a2 = 0.0; ang_2 = 0.5 * M_PI; }
  const double x1 = x1size * cos_a2 * cos_a3, x2 = x2size * cos_a2 * sin_a3, x3 = x3size * sin_a2;
  double lambda = x1;
  if (cfg_.nx[1] > 1 && ang_3 != 0.0) lambda = std::min(lambda, x2);
  if (cfg_.nx[2] > 1 && ang_2 != 0.0) lambda = std::min(lambda, x3);
  if (ang_3_vert) lambda = x2;
  if (ang_2_vert) lambda = x3;
  lw_.sin_a2 = sin_a2; lw_.cos_a2 = cos_a2; lw_.sin_a3 = sin_a3; lw_.cos_a3 = cos_a3;
  lw_.lambda = lambda;
  lw_.k_par = 2.0 * (M_PI) / lambda;
  lw_.d0 = 1.0;
  lw_.u0 = lw_.vflow;
  const double v0 = 0.0, w0 = 0.0;
  lw_.p0 = 1.0 / gam;
  double h0 =
      ((lw_.p0 / lw_.gm1 + 0.5 * lw_.d0 * (lw_.u0 * lw_.u0 + v0 * v0 + w0 * w0)) + lw_.p0) / lw_.d0;
  if (mhd_wave) {
    lw_.bx0 = 1.0;
    lw_.by0 = std::sqrt(2.0);
    lw_.bz0 = 0.5;
    h0 += (lw_.bx0 * lw_.bx0 + lw_.by0 * lw_.by0 + lw_.bz0 * lw_.bz0) / lw_.d0;
    MhdEigensystem(lw_.gm1, lw_.d0, lw_.u0, v0, w0, h0, lw_.bx0, lw_.by0, lw_.bz0, 0.0, 1.0,
                   lw_.ev7, lw_.rem7);
    lw_.tlim_period = lambda / std::abs(lw_.ev7[lw_.wave_flag]) * 1.0;
    lw_.dby = lw_.amp * lw_.rem7[5][lw_.wave_flag];  // linear_wave_mhd.cpp:392-393
    lw_.dbz = lw_.amp * lw_.rem7[6][lw_.wave_flag];
    lw_.set = true;
    return;
  }
  // eigensystem
  const double v1 = lw_.u0, v2 = v0, v3 = w0;
  const double vsq = v1 * v1 + v2 * v2 + v3 * v3;
  const double asq = lw_.gm1 * std::max((h0 - 0.5 * vsq), 1.0e-20);
  const double a = std::sqrt(asq);
  double(*rem)[5] = lw_.rem;
  lw_.ev[0] = v1 - a; lw_.ev[1] = v1; lw_.ev[2] = v1; lw_.ev[3] = v1; lw_.ev[4] = v1 + a;
  rem[0][0] = 1.0; rem[1][0] = v1 - a; rem[2][0] = v2; rem[3][0] = v3; rem[4][0] = h0 - v1 * a;
  rem[0][1] = 0.0; rem[1][1] = 0.0; rem[2][1] = 1.0; rem[3][1] = 0.0; rem[4][1] = v2;
  rem[0][2] = 0.0; rem[1][2] = 0.0; rem[2][2] = 0.0; rem[3][2] = 1.0; rem[4][2] = v3;
  rem[0][3] = 1.0; rem[1][3] = v1; rem[2][3] = v2; rem[3][3] = v3; rem[4][3] = 0.5 * vsq;
  rem[0][4] = 1.0; rem[1][4] = v1 + a; rem[2][4] = v2; rem[3][4] = v3; rem[4][4] = h0 + v1 * a;
  // problem/linear_wave/test = true with tlim = 1: one wave period (linear_wave.cpp:166-174)
  lw_.tlim_period = lambda / std::abs(lw_.ev[lw_.wave_flag]) * 1.0;
  lw_.set = true;
}

void HydroSim::LinearWaveState(const BlockInfo &B, int k, int j, int i, double u[5]) const {
  const int wf = lw_.wave_flag;
  const double amp = lw_.amp;
  const double x = lw_.cos_a2 * (xc(B, 0, i) * lw_.cos_a3 + xc(B, 1, j) * lw_.sin_a3) +
                   xc(B, 2, k) * lw_.sin_a2;
  const double sn = std::sin(lw_.k_par * x);
  u[IDN] = lw_.d0 + amp * sn * lw_.rem[0][wf];
  const double mx = lw_.d0 * lw_.vflow + amp * sn * lw_.rem[1][wf];
  const double my = amp * sn * lw_.rem[2][wf];
  const double mz = amp * sn * lw_.rem[3][wf];
  u[IM1] = mx * lw_.cos_a2 * lw_.cos_a3 - my * lw_.sin_a3 - mz * lw_.sin_a2 * lw_.cos_a3;
  u[IM2] = mx * lw_.cos_a2 * lw_.sin_a3 + my * lw_.cos_a3 - mz * lw_.sin_a2 * lw_.sin_a3;
  u[IM3] = mx * lw_.sin_a2 + mz * lw_.cos_a2;
  u[IEN] = lw_.p0 / lw_.gm1 + 0.5 * lw_.d0 * lw_.u0 * lw_.u0 + amp * sn * lw_.rem[4][wf];
}

// analytic cell-centre state of the magnetised wave (d, m1..m3, E, b1..b3),
// linear_wave_mhd.cpp:200-232
void HydroSim::LinearWaveMhdState(const BlockInfo &B, int k, int j, int i, double u[8]) const {
  const int wf = lw_.wave_flag;
  const double amp = lw_.amp, cos_a2 = lw_.cos_a2, sin_a2 = lw_.sin_a2, cos_a3 = lw_.cos_a3,
               sin_a3 = lw_.sin_a3;
  const double x = cos_a2 * (xc(B, 0, i) * cos_a3 + xc(B, 1, j) * sin_a3) + xc(B, 2, k) * sin_a2;
  const double sn = std::sin(lw_.k_par * x);
  u[IDN] = lw_.d0 + amp * sn * lw_.rem7[0][wf];
  const double mx = lw_.d0 * lw_.vflow + amp * sn * lw_.rem7[1][wf];
  const double my = amp * sn * lw_.rem7[2][wf];
  const double mz = amp * sn * lw_.rem7[3][wf];
  u[IM1] = mx * cos_a2 * cos_a3 - my * sin_a3 - mz * sin_a2 * cos_a3;
  u[IM2] = mx * cos_a2 * sin_a3 + my * cos_a3 - mz * sin_a2 * sin_a3;
  u[IM3] = mx * sin_a2 + mz * cos_a2;
  double e0 = lw_.p0 / lw_.gm1 + 0.5 * lw_.d0 * lw_.u0 * lw_.u0 + amp * sn * lw_.rem7[4][wf];
  e0 += 0.5 * (lw_.bx0 * lw_.bx0 + lw_.by0 * lw_.by0 + lw_.bz0 * lw_.bz0);
  u[IEN] = e0;
  const double bx = lw_.bx0;
  const double by = lw_.by0 + amp * sn * lw_.rem7[5][wf];
  const double bz = lw_.bz0 + amp * sn * lw_.rem7[6][wf];
  u[IB1] = bx * cos_a2 * cos_a3 - by * sin_a3 - bz * sin_a2 * cos_a3;
  u[IB2] = bx * cos_a2 * sin_a3 + by * cos_a3 - bz * sin_a2 * sin_a3;
  u[IB3] = bx * sin_a2 + bz * cos_a2;
}

// vector potential component c = 1, 2, 3 (linear_wave_mhd.cpp:441-480)
double HydroSim::LinearWaveMhdA(int c, double x1, double x2, double x3) const {
  const double cos_a2 = lw_.cos_a2, sin_a2 = lw_.sin_a2, cos_a3 = lw_.cos_a3, sin_a3 = lw_.sin_a3,
               k_par = lw_.k_par;
  const double x = x1 * cos_a2 * cos_a3 + x2 * cos_a2 * sin_a3 + x3 * sin_a2;
  const double y = -x1 * sin_a3 + x2 * cos_a3;
  const double Ay = lw_.bz0 * x - (lw_.dbz / k_par) * std::cos(k_par * (x));
  const double Az = -lw_.by0 * x + (lw_.dby / k_par) * std::cos(k_par * (x)) + lw_.bx0 * y;
  if (c == 1) return -Ay * sin_a3 - Az * sin_a2 * cos_a3;
  if (c == 2) return Ay * cos_a3 - Az * sin_a2 * sin_a3;
  return Az * cos_a2;
}

int HydroSim::ProblemGenerator(const std::string &id, const std::vector<double> &pars) {
  auto par = [&](size_t i, double def) { return pars.size() > i ? pars[i] : def; };
  const bool mhd = cfg_.fluid == APK_FLUID_GLMMHD;
  const double gm1 = cfg_.gamma - 1.0;
  const int is = cfg_.ng, js = (ndim_ >= 2) ? cfg_.ng : 0, ks = (ndim_ >= 3) ? cfg_.ng : 0;
  const size_t ncell = (size_t)n_[0] * n_[1] * n_[2];
  if (id == "linear_wave") LinearWaveSetup(pars, false);
  if (id == "linear_wave_mhd") {
    if (!mhd || ndim_ != 3) return -2;
    LinearWaveSetup(pars, true);
  }
  const bool known = id == "linear_wave" || id == "linear_wave_mhd" || id == "sod" ||
                     id == "orszag_tang" ||
                     id == "uniform" || id == "ot3d" || id == "random" || id == "blast" ||
                     id == "lw_implode" || id == "diffusion";
  if (id == "diffusion" && !mhd && (int)par(0, 20) != 30) return -2;  // diffusion.cpp:35-37
  if (!known) return -1;
  if (id == "orszag_tang" && !mhd) return -2;
  if (&pars != &pgen_pars_) {  // adaptive meshes re-run the generator while Mesh::Initialize refines
    pgen_id_ = id;
    pgen_pars_ = pars;
  }
  std::vector<double> host(BlockFieldElems());
  for (size_t b = 0; b < blocks_.size(); ++b) {
    const BlockInfo &B = blocks_[b];
    std::fill(host.begin(), host.end(), 0.0);
    for (int k = ks; k < ks + cfg_.mb[2]; ++k)
      for (int j = js; j < js + cfg_.mb[1]; ++j)
        for (int i = is; i < is + cfg_.mb[0]; ++i) {
          const size_t c = ((size_t)k * n_[1] + j) * n_[0] + i;
          auto U = [&](int n) -> double & { return host[(size_t)n * ncell + c]; };
          if (id == "linear_wave") {
            double u[5];
            LinearWaveState(B, k, j, i, u);
            for (int n = 0; n < 5; ++n) U(n) = u[n];
          } else if (id == "linear_wave_mhd") {  // linear_wave_mhd.cpp:373-430
            double u[8];
            LinearWaveMhdState(B, k, j, i, u);
            for (int n = 0; n < 5; ++n) U(n) = u[n];
            // cell-centred B = curl of the cell-centred vector potential, centred differences
            const double X1 = xc(B, 0, i), X2 = xc(B, 1, j), X3 = xc(B, 2, k);
            const double X1m = xc(B, 0, i - 1), X1p = xc(B, 0, i + 1);
            const double X2m = xc(B, 1, j - 1), X2p = xc(B, 1, j + 1);
            const double X3m = xc(B, 2, k - 1), X3p = xc(B, 2, k + 1);
            U(IB1) = (LinearWaveMhdA(3, X1, X2p, X3) - LinearWaveMhdA(3, X1, X2m, X3)) / B.dx[1] / 2.0 -
                     (LinearWaveMhdA(2, X1, X2, X3p) - LinearWaveMhdA(2, X1, X2, X3m)) / B.dx[2] / 2.0;
            U(IB2) = (LinearWaveMhdA(1, X1, X2, X3p) - LinearWaveMhdA(1, X1, X2, X3m)) / B.dx[2] / 2.0 -
                     (LinearWaveMhdA(3, X1p, X2, X3) - LinearWaveMhdA(3, X1m, X2, X3)) / B.dx[0] / 2.0;
            U(IB3) = (LinearWaveMhdA(2, X1p, X2, X3) - LinearWaveMhdA(2, X1m, X2, X3)) / B.dx[0] / 2.0 -
                     (LinearWaveMhdA(1, X1, X2p, X3) - LinearWaveMhdA(1, X1, X2m, X3)) / B.dx[1] / 2.0;
          } else if (id == "sod") {  // sod.cpp:23-50
            const double rho_l = par(0, 1.0), pres_l = par(1, 1.0), u_l = par(2, 0.0),
                         rho_r = par(3, 0.125), pres_r = par(4, 0.1), u_r = par(5, 0.0),
                         x_discont = par(6, 0.5);
            if (xc(B, 0, i) < x_discont) {
              U(IDN) = rho_l;
              U(IM1) = rho_l * u_l;
              U(IEN) = 0.5 * rho_l * u_l * u_l + pres_l / (cfg_.gamma - 1.0);
            } else {
              U(IDN) = rho_r;
              U(IM1) = rho_r * u_r;
              U(IEN) = 0.5 * rho_r * u_r * u_r + pres_r / (cfg_.gamma - 1.0);
            }
          } else if (id == "orszag_tang") {  // orszag_tang.cpp:34-61
            const double B0 = 1.0 / std::sqrt(4.0 * M_PI), d0 = 25.0 / (36.0 * M_PI), v0 = 1.0,
                         p0 = 5.0 / (12.0 * M_PI);
            U(IDN) = d0;
            U(IM1) = d0 * v0 * std::sin(2.0 * M_PI * xc(B, 1, j));
            U(IM2) = -d0 * v0 * std::sin(2.0 * M_PI * xc(B, 0, i));
            U(IM3) = 0.0;
            U(IB1) = B0 * std::sin(2.0 * M_PI * xc(B, 1, j));
            U(IB2) = B0 * std::sin(4.0 * M_PI * xc(B, 0, i));
            U(IB3) = 0.0;
            U(IEN) = p0 / gm1 + 0.5 * (SQR(U(IB1)) + SQR(U(IB2)) + SQR(U(IB3)) +
                                       (SQR(U(IM1)) + SQR(U(IM2)) + SQR(U(IM3))) / U(IDN));
          } else if (id == "uniform") {  // turbulence.cpp:241-356 with zero velocity
            const double rho0 = par(0, 1.0), p0 = par(1, 1.0), b0 = par(2, 0.01);
            U(IDN) = rho0;
            U(IEN) = p0 / gm1 + 0.5 * rho0 * (SQR(0.0) + SQR(0.0) + SQR(0.0));
            if (mhd) {
              U(IB1) = b0;
              U(IEN) += 0.5 * b0 * b0;
            }
          } else if (id == "ot3d") {
            const double B0 = 1.0 / std::sqrt(4.0 * M_PI), d0 = 25.0 / (36.0 * M_PI),
                         p0 = 5.0 / (12.0 * M_PI);
            const double x = xc(B, 0, i), y = xc(B, 1, j), z = xc(B, 2, k);
            const double v1 = std::sin(2.0 * M_PI * y), v2 = -std::sin(2.0 * M_PI * x),
                         v3 = 0.2 * std::sin(2.0 * M_PI * z);
            U(IDN) = d0;
            U(IM1) = d0 * v1;
            U(IM2) = d0 * v2;
            U(IM3) = d0 * v3;
            double eb = 0.0;
            if (mhd) {
              U(IB1) = B0 * std::sin(2.0 * M_PI * y);
              U(IB2) = B0 * std::sin(4.0 * M_PI * x);
              U(IB3) = B0 * 0.2 * std::sin(2.0 * M_PI * x);
              eb = 0.5 * (SQR(U(IB1)) + SQR(U(IB2)) + SQR(U(IB3)));
            }
            U(IEN) = p0 / gm1 + 0.5 * d0 * (SQR(v1) + SQR(v2) + SQR(v3)) + eb;
          } else if (id == "diffusion") {  // src/pgen/diffusion.cpp:28-158
            // pars = iprob, Bx, By, t0, amp, diffusion coefficient of the process under test
            const int iprob = (int)par(0, 20);
            const double Bx = par(1, 0.0), By = par(2, 0.0), t0 = par(3, 0.5), amp = par(4, 1e-6),
                         dc = par(5, 0.0);
            const double X = xc(B, 0, i), Y = xc(B, 1, j), Z = xc(B, 2, k);
            double eint = -1.0;
            U(IDN) = 1.0;
            if (iprob == 0) {         // step functions along x1 / x2 / x3
              U(IB1) = Bx; U(IB2) = By;
              eint = X <= 0.0 ? 10.0 : 12.0;
            } else if (iprob == 1) {
              U(IB2) = Bx; U(IB3) = By;
              eint = Y <= 0.0 ? 10.0 : 12.0;
            } else if (iprob == 2) {
              U(IB3) = Bx; U(IB1) = By;
              eint = Z <= 0.0 ? 10.0 : 12.0;
            } else if (iprob == 10) {  // heat diffusion of a 1-D Gaussian
              U(IB1) = Bx; U(IB2) = By;
              const double eff = Bx == 0.0 ? dc : dc * Bx * Bx;
              eint = 1 + amp / std::sqrt(4. * M_PI * eff * t0) *
                             std::exp(-(std::pow(X, 2.)) / (4. * eff * t0));
            } else if (iprob >= 20 && iprob <= 22) {  // ring diffusion in a coordinate plane
              const double x = iprob == 20 ? X : (iprob == 21 ? Y : Z);
              const double y = iprob == 20 ? Y : (iprob == 21 ? Z : X);
              const double r = std::sqrt(SQR(x) + SQR(y)), phi = std::atan2(y, x);
              const int b1 = iprob == 20 ? IB1 : (iprob == 21 ? IB2 : IB3);
              const int b2 = iprob == 20 ? IB2 : (iprob == 21 ? IB3 : IB1);
              U(b1) = y / r;
              U(b2) = -x / r;
              eint = (std::abs(r - 0.6) < 0.1 && std::abs(phi) < M_PI / 12.0) ? 12.0 : 10.0;
            } else if (iprob == 30) {  // viscous diffusion of a 1-D Gaussian
              U(IM2) = U(IDN) * amp / std::pow(std::sqrt(4. * M_PI * dc * t0), 1.0) *
                       std::exp(-(std::pow(X, 2.)) / (4. * dc * t0));
              eint = 1.0 / (cfg_.gamma * (cfg_.gamma - 1.0));
            } else if (iprob == 40) {  // Ohmic diffusion of a 1-D Gaussian
              U(IB2) = amp / std::pow(std::sqrt(4. * M_PI * dc * t0), 1.0) *
                       std::exp(-(std::pow(X, 2.)) / (4. * dc * t0));
              eint = 1.0 / (cfg_.gamma * (cfg_.gamma - 1.0));
            }
            if (!(eint > 0.0)) return -2;  // "Missing init of eint"
            U(IEN) = U(IDN) * eint + 0.5 * ((SQR(U(IM1)) + SQR(U(IM2)) + SQR(U(IM3))) / U(IDN));
            if (mhd) U(IEN) += 0.5 * (SQR(U(IB1)) + SQR(U(IB2)) + SQR(U(IB3)));
          } else if (id == "lw_implode") {  // lw_implode.cpp:34-75 (Liska-Wendroff implosion)
            const double d_in = par(0, 0.125), p_in = par(1, 0.14), d_out = par(2, 1.0),
                         p_out = par(3, 1.0);
            // y0 between cell centres so that the initial state is symmetric (:47-60)
            double y0 = 0.5 * (cfg_.xmax[1] + cfg_.xmin[1]);
            for (int jj = js; jj < js + cfg_.mb[1]; ++jj)
              if (xc(B, 1, jj) > y0) {
                y0 = (B.xmin[1] + jj * B.dx[1]) + 0.5 * B.dx[1];
                break;
              }
            if (xc(B, 1, j) > (y0 - xc(B, 0, i))) {
              U(IDN) = d_out;
              U(IEN) = p_out / gm1;
            } else {
              U(IDN) = d_in;
              U(IEN) = p_in / gm1;
            }
          } else if (id == "blast") {
            const double rad = par(0, 0.1), pa = par(1, 0.1), prat = par(2, 100.0), da = par(3, 1.0);
            const double x = xc(B, 0, i), y = (ndim_ > 1) ? xc(B, 1, j) : 0.0,
                         z = (ndim_ > 2) ? xc(B, 2, k) : 0.0;
            // centre of the blast: problem/blast/x1_0, x2_0, x3_0 (blast.cpp:131-133)
            const double x0 = par(4, 0.0), y0 = par(5, 0.0), z0 = par(6, 0.0);
            const double r = std::sqrt(SQR(x - x0) + SQR(y - y0) + SQR(z - z0));
            U(IDN) = da;
            U(IEN) = ((r < rad) ? prat * pa : pa) / gm1;
          } else {  // random
            const std::uint64_t seed = (std::uint64_t)par(0, 1);
            const double amp = par(1, 0.2);
            const std::uint64_t gi = (std::uint64_t)(B.loc[0] * cfg_.mb[0] + (i - is));
            const std::uint64_t gj = (std::uint64_t)(B.loc[1] * cfg_.mb[1] + (j - js));
            const std::uint64_t gk = (std::uint64_t)(B.loc[2] * cfg_.mb[2] + (k - ks));
            const std::uint64_t cell =
                (gk * (std::uint64_t)cfg_.nx[1] + gj) * (std::uint64_t)cfg_.nx[0] + gi;
            double r[8];
            for (int q = 0; q < 8; ++q)
              r[q] = 2.0 * hash01(seed * 1315423911ULL + cell * 8 + q) - 1.0;
            const double d = 1.0 + amp * r[0], pr = 1.0 + amp * r[1];
            const double v1 = amp * r[2], v2 = amp * r[3], v3 = amp * r[4];
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
          }
          for (int n = nhydro_; n < nvar_; ++n) U(n) = U(IDN) * (0.25 + 0.5 * ((i + j + k + n) % 2));
        }
    CopyFromHost(0, (int)b, host.data());
  }
  return 0;
}

// linear_wave::UserWorkAfterLoop (linear_wave.cpp:183-295); sums over LOCAL blocks, un-normalised
// per-variable L1 sums in out[1..5] are divided by the domain volume, out[0] = RMS (valid for
// single-rank runs; multi-rank callers sum out[1..5] first).
void HydroSim::LinearWaveErrors(double *out) {
  const int nerr = lw_.mhd ? 8 : 5;  // linear_wave_mhd.cpp:186: psi excluded
  for (int n = 0; n < 9; ++n) out[n] = 0.0;
  double l1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<double> host(BlockFieldElems());
  const int is = cfg_.ng, js = (ndim_ >= 2) ? cfg_.ng : 0, ks = (ndim_ >= 3) ? cfg_.ng : 0;
  const size_t ncell = (size_t)n_[0] * n_[1] * n_[2];
  for (size_t b = 0; b < blocks_.size(); ++b) {
    const BlockInfo &B = blocks_[b];
    CopyToHost(0, (int)b, host.data());
    const double vol = B.dx[0] * B.dx[1] * B.dx[2];
    for (int k = ks; k < ks + cfg_.mb[2]; ++k)
      for (int j = js; j < js + cfg_.mb[1]; ++j)
        for (int i = is; i < is + cfg_.mb[0]; ++i) {
          double u[8];
          if (lw_.mhd) LinearWaveMhdState(B, k, j, i, u);
          else LinearWaveState(B, k, j, i, u);
          const size_t c = ((size_t)k * n_[1] + j) * n_[0] + i;
          for (int n = 0; n < nerr; ++n) l1[n] += std::abs(u[n] - host[(size_t)n * ncell + c]) * vol;
        }
  }
  const double vol = (cfg_.xmax[0] - cfg_.xmin[0]) * (cfg_.xmax[1] - cfg_.xmin[1]) *
                     (cfg_.xmax[2] - cfg_.xmin[2]);
  double rms = 0.0;
  for (int n = 0; n < nerr; ++n) {
    l1[n] = l1[n] / vol;
    rms += SQR(l1[n]);
    out[1 + n] = l1[n];
  }
  out[0] = std::sqrt(rms);
}

}  // namespace apk_host
