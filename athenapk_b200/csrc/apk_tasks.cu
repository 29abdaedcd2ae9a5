// apk_tasks.cu -- un-fused task kernels: one CUDA kernel per AthenaPK task so the library is
// a drop-in for each FluxFun_t / Update / source / FillDerived / EstimateTimestep hook and
// materialises cons.flux when FOFC, AMR flux correction or diffusion need it.
// Compiled twice: -DAPK_NS=apk_fast (FMA) and -DAPK_NS=apk_strict -fmad=false.
//
// Layout: fields [nvar][n3][n2][n1] per block (i fastest) -> a warp covers 32 consecutive i of
// one (var,k,j) row: every global load/store below is a fully coalesced 256 B row segment.
#include <cstdint>
#include <cstdio>

#include "apk_launch.h"
#include "apk_math.cuh"

namespace APK_NS {

namespace {

constexpr int kThreads = 128;

struct Range3 {
  int il, iu, jl, ju, kl, ku;
};

__device__ __forceinline__ bool decode(long t, const Range3 &r, int nb, int &b, int &k, int &j,
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

// --------------------------------------------------------------------------------------------
// K1-K3 / K4: Hydro::CalculateFluxes (hydro.cpp:1025-1208) and CalculateFluxesTight (:980-1022)
// one thread = one face of sweep direction `dir`
// --------------------------------------------------------------------------------------------
template <int FLUID, int RECON, int RS>
__global__ void __launch_bounds__(kThreads)
k_calc_flux(ApkGeom g, const apk_block_t *__restrict__ blocks, int dir, Range3 r, double gamma,
            double c_h) {
  constexpr int NH = (FLUID == APK_FLUID_GLMMHD) ? 9 : 5;
  int b, k, j, i;
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, r, g.nb, b, k, j, i)) return;
  const apk_block_t &B = blocks[b];
  const long st = (dir == 0) ? 1 : (dir == 1) ? g.n1 : (long)g.n1 * g.n2;
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  const double dx = B.dx[dir];
  const double *__restrict__ prim = B.prim;
  double L[NH], R[NH], F[NH];
#pragma unroll
  for (int n = 0; n < NH; ++n) {
    const int v = var_of_slot(n, dir);
    const double *q = prim + (long)v * g.ncell + c;
    const bool pos = (n == 0 || n == 4);
    double unused;
    recon_cell<RECON>(q - st, st, dx, pos, L[n], unused);  // ql of face i comes from cell i-1
    recon_cell<RECON>(q, st, dx, pos, unused, R[n]);       // qr of face i comes from cell i
  }
  riemann_solve<FLUID, RS>(L, R, gamma, c_h, F);
  double *__restrict__ flux = B.flux[dir];
#pragma unroll
  for (int n = 0; n < NH; ++n) flux[(long)var_of_slot(n, dir) * g.ncell + c] = F[n];
  // passive scalars: upwinded by the sign of the mass flux (hydro.cpp:1089-1097)
  for (int n = NH; n < g.nvar; ++n) {
    const double *q = prim + (long)n * g.ncell + c;
    double wl, wr, unused;
    recon_cell<RECON>(q - st, st, dx, false, wl, unused);
    recon_cell<RECON>(q, st, dx, false, unused, wr);
    flux[(long)n * g.ncell + c] = (F[0] >= 0.0) ? F[0] * wl : F[0] * wr;
  }
}

template <int FLUID, int RECON, int RS>
int run_calc_flux(const ApkGeom &g, const apk_block_t *blocks, bool tight, double gamma,
                  double c_h, cudaStream_t s) {
  for (int dir = 0; dir < g.ndim; ++dir) {
    Range3 r;
    if (tight) {  // hydro.cpp:1006-1009
      r = {g.is, g.ie + 1, g.js, g.je + (g.ndim >= 2), g.ks, g.ke + (g.ndim >= 3)};
    } else {      // hydro.cpp:1031-1039, 1104-1110, 1158
      r = {g.is - 1, g.ie + 1, g.js, g.je, g.ks, g.ke};
      if (g.ndim >= 2) { r.jl = g.js - 1; r.ju = g.je + 1; }
      if (g.ndim >= 3) { r.kl = g.ks - 1; r.ku = g.ke + 1; }
      if (dir == 0) r.il = g.is;
      if (dir == 1) r.jl = g.js;
      if (dir == 2) r.kl = g.ks;
    }
    const long total = (long)(r.iu - r.il + 1) * (r.ju - r.jl + 1) * (r.ku - r.kl + 1) * g.nb;
    const unsigned grid = (unsigned)((total + kThreads - 1) / kThreads);
    k_calc_flux<FLUID, RECON, RS><<<grid, kThreads, 0, s>>>(g, blocks, dir, r, gamma, c_h);
    apk_common::count_launch();
  }
  return 0;
}

template <int FLUID, int RS>
int dispatch_recon(int recon, const ApkGeom &g, const apk_block_t *blocks, double gamma,
                   double c_h, cudaStream_t s) {
  switch (recon) {
  case APK_RECON_DC: return run_calc_flux<FLUID, APK_RECON_DC, RS>(g, blocks, false, gamma, c_h, s);
  case APK_RECON_PLM: return run_calc_flux<FLUID, APK_RECON_PLM, RS>(g, blocks, false, gamma, c_h, s);
  case APK_RECON_PPM: return run_calc_flux<FLUID, APK_RECON_PPM, RS>(g, blocks, false, gamma, c_h, s);
  case APK_RECON_WENOZ: return run_calc_flux<FLUID, APK_RECON_WENOZ, RS>(g, blocks, false, gamma, c_h, s);
  case APK_RECON_WENO3: return run_calc_flux<FLUID, APK_RECON_WENO3, RS>(g, blocks, false, gamma, c_h, s);
  case APK_RECON_LIMO3: return run_calc_flux<FLUID, APK_RECON_LIMO3, RS>(g, blocks, false, gamma, c_h, s);
  }
  apk_common::set_error("apk_calc_fluxes: unknown reconstruction");
  return 1;
}

// --------------------------------------------------------------------------------------------
// K6: Update::UpdateWithFluxDivergence (update.cpp:115-141, FluxDivHelper update.hpp:42-56)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double flux_div(const ApkGeom &g, const apk_block_t &B, int l, long c,
                                           double a1, double a2, double a3, double vol) {
  const long o = (long)l * g.ncell + c;
  double du = (a1 * B.flux[0][o + 1] - a1 * B.flux[0][o]);
  if (g.ndim >= 2) du += (a2 * B.flux[1][o + g.n1] - a2 * B.flux[1][o]);
  if (g.ndim == 3) du += (a3 * B.flux[2][o + (long)g.n1 * g.n2] - a3 * B.flux[2][o]);
  return fdiv(-du, vol);
}

__global__ void __launch_bounds__(kThreads)
k_flux_div_update(ApkGeom g, const apk_block_t *__restrict__ blocks, double g0, double g1,
                  double bdt) {
  int b, k, j, i;
  const Range3 r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, r, g.nb, b, k, j, i)) return;
  const apk_block_t &B = blocks[b];
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  // uniform_coordinates.hpp:37-40
  const double a1 = B.dx[1] * B.dx[2], a2 = B.dx[0] * B.dx[2], a3 = B.dx[0] * B.dx[1];
  const double vol = B.dx[0] * B.dx[1] * B.dx[2];
  for (int l = 0; l < g.nvar; ++l) {
    const long o = (long)l * g.ncell + c;
    B.cons[o] = g0 * B.cons[o] + g1 * B.cons_u1[o] + bdt * flux_div(g, B, l, c, a1, a2, a3, vol);
  }
}

// --------------------------------------------------------------------------------------------
// K5: FirstOrderFluxCorrect (hydro.cpp:1223-1342), Jacobi form: flag pass + face fix pass
// --------------------------------------------------------------------------------------------
template <int FLUID>
__global__ void __launch_bounds__(kThreads)
k_fofc_flag(ApkGeom g, const apk_block_t *__restrict__ blocks, double g0, double g1, double bdt,
            int attempt, unsigned char *__restrict__ flags, unsigned long long *counter) {
  constexpr int NH = (FLUID == APK_FLUID_GLMMHD) ? 9 : 5;
  int b, k, j, i;
  const Range3 r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, r, g.nb, b, k, j, i)) return;
  const apk_block_t &B = blocks[b];
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  const double a1 = B.dx[1] * B.dx[2], a2 = B.dx[0] * B.dx[2], a3 = B.dx[0] * B.dx[1];
  const double vol = B.dx[0] * B.dx[1] * B.dx[2];
  double nc[NH];
#pragma unroll
  for (int v = 0; v < NH; ++v) {
    const long o = (long)v * g.ncell + c;
    nc[v] = g0 * B.cons[o] + g1 * B.cons_u1[o] + bdt * flux_div(g, B, v, c, a1, a2, a3, vol);
  }
  double new_p = nc[4] - fdiv(0.5 * (sqr(nc[1]) + sqr(nc[2]) + sqr(nc[3])), nc[0]);
  if constexpr (FLUID == APK_FLUID_GLMMHD) new_p -= 0.5 * (sqr(nc[5]) + sqr(nc[6]) + sqr(nc[7]));
  if (nc[0] > 0.0 && new_p > 0.0) return;
  if (attempt > 2 && nc[0] > 0.0 && new_p < 0.0) return;
  flags[(long)b * g.ncell + c] = 1;
  atomicAdd(counter, 1ULL);
}

template <int FLUID>
__global__ void __launch_bounds__(kThreads)
k_fofc_fix(ApkGeom g, const apk_block_t *__restrict__ blocks, int dir, double gamma, double c_h,
           const unsigned char *__restrict__ flags) {
  constexpr int NH = (FLUID == APK_FLUID_GLMMHD) ? 9 : 5;
  int b, k, j, i;
  // faces of interior cells in direction dir
  Range3 r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  if (dir == 0) r.iu += 1;
  if (dir == 1) r.ju += 1;
  if (dir == 2) r.ku += 1;
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, r, g.nb, b, k, j, i)) return;
  const long st = (dir == 0) ? 1 : (dir == 1) ? g.n1 : (long)g.n1 * g.n2;
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  const unsigned char *fb = flags + (long)b * g.ncell;
  if (!(fb[c] | fb[c - st])) return;
  const apk_block_t &B = blocks[b];
  double L[NH], R[NH], F[NH];
#pragma unroll
  for (int n = 0; n < NH; ++n) {
    const double *q = B.prim + (long)var_of_slot(n, dir) * g.ncell + c;
    L[n] = q[-st];
    R[n] = q[0];
  }
  riemann_solve<FLUID, APK_RIEMANN_LLF>(L, R, gamma, c_h, F);
  double *__restrict__ flux = B.flux[dir];
#pragma unroll
  for (int n = 0; n < NH; ++n) flux[(long)var_of_slot(n, dir) * g.ncell + c] = F[n];
  for (int n = NH; n < g.nvar; ++n) {
    const double *q = B.prim + (long)n * g.ncell + c;
    flux[(long)n * g.ncell + c] = (F[0] >= 0.0) ? F[0] * q[-st] : F[0] * q[0];
  }
}

// --------------------------------------------------------------------------------------------
// K7: GLMMHD::DednerSource (dedner_source.cpp:17-75); coeff = exp(-alpha c_h bdt / mindx)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_dedner(ApkGeom g, const apk_block_t *__restrict__ blocks, double coeff, double bdt,
         int extended) {
  int b, k, j, i;
  const Range3 r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  if (!decode((long)blockIdx.x * blockDim.x + threadIdx.x, r, g.nb, b, k, j, i)) return;
  const apk_block_t &B = blocks[b];
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  if (extended) {
    const long sj = g.n1, sk = (g.ndim < 3) ? 0 : (long)g.n1 * g.n2;
    const double *pb1 = B.prim + 5 * g.ncell + c, *pb2 = B.prim + 6 * g.ncell + c,
                 *pb3 = B.prim + 7 * g.ncell + c, *pps = B.prim + 8 * g.ncell + c;
    const double divB = 0.5 * (fdiv((pb1[1] - pb1[-1]), B.dx[0]) + fdiv((pb2[sj] - pb2[-sj]), B.dx[1]) +
                               fdiv((pb3[sk] - pb3[-sk]), B.dx[2]));
    B.cons[1 * g.ncell + c] -= bdt * divB * pb1[0];
    B.cons[2 * g.ncell + c] -= bdt * divB * pb2[0];
    B.cons[3 * g.ncell + c] -= bdt * divB * pb3[0];
    B.cons[4 * g.ncell + c] -=
        0.5 * bdt *
        (fdiv(pb1[0] * (pps[1] - pps[-1]), B.dx[0]) + fdiv(pb2[0] * (pps[sj] - pps[-sj]), B.dx[1]) +
         fdiv(pb3[0] * (pps[sk] - pps[-sk]), B.dx[2]));
  }
  B.cons[8 * g.ncell + c] *= coeff;
}

// --------------------------------------------------------------------------------------------
// K10: ConservedToPrimitive over the entire index range (adiabatic_glmmhd.cpp:33-56)
// --------------------------------------------------------------------------------------------
template <int NH>
__global__ void __launch_bounds__(kThreads)
k_cons_to_prim(ApkGeom g, const apk_block_t *__restrict__ blocks, EosPars eos, int *status) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.ncell * g.nb) return;
  const int b = (int)(t / g.ncell);
  const long c = t - (long)b * g.ncell;
  const apk_block_t &B = blocks[b];
  double u[NH], w[NH];
#pragma unroll
  for (int n = 0; n < NH; ++n) u[n] = B.cons[(long)n * g.ncell + c];
  const int err = cons_to_prim<NH>(u, w, eos);
  if (err && status) atomicMax(status, err);
  // the reference writes through references: rho, m, E may have been floored
#pragma unroll
  for (int n = 0; n < 5; ++n) B.cons[(long)n * g.ncell + c] = u[n];
#pragma unroll
  for (int n = 0; n < NH; ++n) B.prim[(long)n * g.ncell + c] = w[n];
  const double di = frcp(u[0]);
  for (int n = NH; n < g.nvar; ++n) B.prim[(long)n * g.ncell + c] = B.cons[(long)n * g.ncell + c] * di;
}

// --------------------------------------------------------------------------------------------
// K11: EstimateHyperbolicTimestep (hydro.cpp:828-910): min over cells/dirs of dx/(|v|+c)
// positive doubles order like their bit patterns -> atomicMin on the uint64 image
// --------------------------------------------------------------------------------------------
template <int FLUID>
__global__ void __launch_bounds__(256)
k_min_dt(ApkGeom g, const apk_block_t *__restrict__ blocks, double gamma,
         unsigned long long *out) {
  double m = DBL_MAX;
  const Range3 r = {g.is, g.ie, g.js, g.je, g.ks, g.ke};
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x;; t += (long)gridDim.x * blockDim.x) {
    int b, k, j, i;
    if (!decode(t, r, g.nb, b, k, j, i)) break;
    const apk_block_t &B = blocks[b];
    const long c = ((long)k * g.n2 + j) * g.n1 + i;
    const double d = B.prim[c], v1 = B.prim[g.ncell + c], v2 = B.prim[2 * g.ncell + c],
                 v3 = B.prim[3 * g.ncell + c], p = B.prim[4 * g.ncell + c];
    double lx, ly, lz;
    if constexpr (FLUID == APK_FLUID_EULER) {
      lx = ly = lz = sound_speed(gamma, d, p);
    } else {
      const double b1 = B.prim[5 * g.ncell + c], b2 = B.prim[6 * g.ncell + c],
                   b3 = B.prim[7 * g.ncell + c];
      lx = fast_speed(gamma, d, p, b1, b2, b3);
      ly = (g.ndim > 1) ? fast_speed(gamma, d, p, b2, b3, b1) : 0.0;
      lz = (g.ndim > 2) ? fast_speed(gamma, d, p, b3, b1, b2) : 0.0;
    }
    m = fmin(m, fdiv(B.dx[0], (fabs(v1) + lx)));
    if (g.ndim > 1) m = fmin(m, fdiv(B.dx[1], (fabs(v2) + ly)));
    if (g.ndim > 2) m = fmin(m, fdiv(B.dx[2], (fabs(v3) + lz)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : DBL_MAX;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMin(out, (unsigned long long)__double_as_longlong(m));
  }
}

__global__ void k_init_u64(unsigned long long *p, unsigned long long v) { *p = v; }

inline unsigned grid_for(long total, int threads) {
  return (unsigned)((total + threads - 1) / threads);
}

}  // namespace

// ============================================================================================
// launchers
// ============================================================================================
int launch_calc_fluxes(const apk_pack_desc_t &d, const apk_block_t *blocks, double c_h,
                       int recon, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  if (recon < 0) recon = d.recon;
  if (d.fluid == APK_FLUID_EULER) {
    switch (d.riemann) {
    case APK_RIEMANN_NONE:
      return run_calc_flux<APK_FLUID_EULER, APK_RECON_DC, APK_RIEMANN_NONE>(g, blocks, false, d.gamma, c_h, s);
    case APK_RIEMANN_HLLE: return dispatch_recon<APK_FLUID_EULER, APK_RIEMANN_HLLE>(recon, g, blocks, d.gamma, c_h, s);
    case APK_RIEMANN_HLLC: return dispatch_recon<APK_FLUID_EULER, APK_RIEMANN_HLLC>(recon, g, blocks, d.gamma, c_h, s);
    case APK_RIEMANN_LLF:
      return run_calc_flux<APK_FLUID_EULER, APK_RECON_DC, APK_RIEMANN_LLF>(g, blocks, true, d.gamma, c_h, s);
    }
  } else if (d.fluid == APK_FLUID_GLMMHD) {
    switch (d.riemann) {
    case APK_RIEMANN_NONE:
      return run_calc_flux<APK_FLUID_GLMMHD, APK_RECON_DC, APK_RIEMANN_NONE>(g, blocks, false, d.gamma, c_h, s);
    case APK_RIEMANN_HLLE: return dispatch_recon<APK_FLUID_GLMMHD, APK_RIEMANN_HLLE>(recon, g, blocks, d.gamma, c_h, s);
    case APK_RIEMANN_HLLD: return dispatch_recon<APK_FLUID_GLMMHD, APK_RIEMANN_HLLD>(recon, g, blocks, d.gamma, c_h, s);
    case APK_RIEMANN_LLF:
      return run_calc_flux<APK_FLUID_GLMMHD, APK_RECON_DC, APK_RIEMANN_LLF>(g, blocks, true, d.gamma, c_h, s);
    }
  }
  apk_common::set_error("apk_calc_fluxes: unsupported fluid/riemann combination");
  return 1;
}

int launch_fofc_flag(const apk_pack_desc_t &d, const apk_block_t *blocks, double g0, double g1,
                     double bdt, int attempt, unsigned char *flags, unsigned long long *counter,
                     cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  const long total = (long)d.nx[0] * d.nx[1] * d.nx[2] * g.nb;
  if (d.fluid == APK_FLUID_GLMMHD)
    k_fofc_flag<APK_FLUID_GLMMHD><<<grid_for(total, kThreads), kThreads, 0, s>>>(g, blocks, g0, g1, bdt, attempt, flags, counter);
  else
    k_fofc_flag<APK_FLUID_EULER><<<grid_for(total, kThreads), kThreads, 0, s>>>(g, blocks, g0, g1, bdt, attempt, flags, counter);
  apk_common::count_launch();
  return 0;
}

int launch_fofc_fix(const apk_pack_desc_t &d, const apk_block_t *blocks, double c_h,
                    const unsigned char *flags, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  for (int dir = 0; dir < g.ndim; ++dir) {
    long n[3] = {d.nx[0], d.nx[1], d.nx[2]};
    n[dir] += 1;
    const long total = n[0] * n[1] * n[2] * g.nb;
    if (d.fluid == APK_FLUID_GLMMHD)
      k_fofc_fix<APK_FLUID_GLMMHD><<<grid_for(total, kThreads), kThreads, 0, s>>>(g, blocks, dir, d.gamma, c_h, flags);
    else
      k_fofc_fix<APK_FLUID_EULER><<<grid_for(total, kThreads), kThreads, 0, s>>>(g, blocks, dir, d.gamma, c_h, flags);
    apk_common::count_launch();
  }
  return 0;
}

int launch_flux_div_update(const apk_pack_desc_t &d, const apk_block_t *blocks, double g0,
                           double g1, double bdt, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  const long total = (long)d.nx[0] * d.nx[1] * d.nx[2] * g.nb;
  k_flux_div_update<<<grid_for(total, kThreads), kThreads, 0, s>>>(g, blocks, g0, g1, bdt);
  apk_common::count_launch();
  return 0;
}

int launch_dedner(const apk_pack_desc_t &d, const apk_block_t *blocks, double coeff, double bdt,
                  int extended, cudaStream_t s) {
  if (d.fluid != APK_FLUID_GLMMHD) return 0;
  const ApkGeom g = apk_make_geom(d);
  const long total = (long)d.nx[0] * d.nx[1] * d.nx[2] * g.nb;
  k_dedner<<<grid_for(total, kThreads), kThreads, 0, s>>>(g, blocks, coeff, bdt, extended);
  apk_common::count_launch();
  return 0;
}

int launch_cons_to_prim(const apk_pack_desc_t &d, const apk_block_t *blocks, int *status,
                        cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  const EosPars eos = {d.gamma, d.dfloor, d.pfloor, d.efloor, d.vceil, d.eceil};
  const long total = g.ncell * g.nb;
  if (d.fluid == APK_FLUID_GLMMHD)
    k_cons_to_prim<9><<<grid_for(total, kThreads), kThreads, 0, s>>>(g, blocks, eos, status);
  else
    k_cons_to_prim<5><<<grid_for(total, kThreads), kThreads, 0, s>>>(g, blocks, eos, status);
  apk_common::count_launch();
  return 0;
}

int launch_min_dt(const apk_pack_desc_t &d, const apk_block_t *blocks, double *d_out,
                  cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  unsigned long long *out = reinterpret_cast<unsigned long long *>(d_out);
  const double big = DBL_MAX;
  unsigned long long bits;
  memcpy(&bits, &big, sizeof(bits));
  k_init_u64<<<1, 1, 0, s>>>(out, bits);
  const long total = (long)d.nx[0] * d.nx[1] * d.nx[2] * g.nb;
  unsigned grid = grid_for(total, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (d.fluid == APK_FLUID_GLMMHD)
    k_min_dt<APK_FLUID_GLMMHD><<<grid, 256, 0, s>>>(g, blocks, d.gamma, out);
  else
    k_min_dt<APK_FLUID_EULER><<<grid, 256, 0, s>>>(g, blocks, d.gamma, out);
  apk_common::count_launch(2);
  return 0;
}

}  // namespace APK_NS
