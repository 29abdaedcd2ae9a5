// apk_amr.cu -- AMR building blocks: restriction (cells and face fluxes),
// AthenaPK's multi-D limited prolongation, pressure-gradient refinement tagging.
// Compiled twice (apk_fast: FMA contraction, apk_strict: -fmad=false = bit-identical to the
// reference's CPU build); operation order follows
//   external/parthenon/src/prolong_restrict/pr_ops.hpp:75-165, src/hydro/prolongation/
//   custom_ops.hpp:49-186, coordinates/uniform_coordinates.hpp:31-58, src/refinement/gradient.cpp.
// Mapping: one thread per coarse cell and variable (grid-stride), blockIdx.y = item / meshblock;
// pure streaming kernels (HBM-bound by construction, a few launches per remesh / exchange).
#include <cfloat>
#include <cmath>

#include "apk_launch.h"

#ifndef APK_NS
#define APK_NS apk_fast
#endif

namespace APK_NS {
namespace {
constexpr int kThreads = 256;

struct AmrCoords {  // uniform_coordinates.hpp:31-58 (fine block coords and their coarsen=2 copy)
  double dx[3], xmin[3], cdx[3], cxmin[3];
  int istart[3];
};

__device__ __forceinline__ AmrCoords make_coords(int ndim, int ng, const apk_amr_item_t &it) {
  AmrCoords c;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    c.istart[d] = (d < ndim) ? ng : 0;
    c.dx[d] = it.dx[d];
    c.xmin[d] = it.xmin[d] - c.istart[d] * c.dx[d];
    c.cxmin[d] = c.xmin[d] + c.istart[d] * c.dx[d] * (1 - 2);
    c.cdx[d] = c.dx[d];
  }
  c.cdx[0] *= 2;
  c.cdx[1] *= (c.istart[1] > 0 ? 2 : 1);
  c.cdx[2] *= (c.istart[2] > 0 ? 2 : 1);
  return c;
}

__device__ __forceinline__ bool decode(const apk_amr_item_t &it, int nvar, long t, int &n, int &k,
                                       int &j, int &i) {
  const int e0 = it.hi[0] - it.lo[0] + 1, e1 = it.hi[1] - it.lo[1] + 1,
            e2 = it.hi[2] - it.lo[2] + 1;
  const long cells = (long)e0 * e1 * e2;
  if (e0 <= 0 || e1 <= 0 || e2 <= 0 || t >= cells * nvar) return false;
  n = (int)(t / cells);
  long r = t - (long)n * cells;
  k = (int)(r / ((long)e0 * e1));
  r -= (long)k * e0 * e1;
  j = (int)(r / e0);
  i = (int)(r - (long)j * e0);
  k += it.lo[2];
  j += it.lo[1];
  i += it.lo[0];
  return true;
}

__global__ void __launch_bounds__(kThreads)
k_restrict_average(int ndim, int ng, int nvar, int el, const apk_amr_item_t *__restrict__ items) {
  const apk_amr_item_t it = items[blockIdx.y];
  const AmrCoords c = make_coords(ndim, ng, it);
  const bool inc1 = (ndim > 0) && (el == 0 || el == 2 || el == 3);  // pr_ops.hpp:118-123
  const bool inc2 = (ndim > 1) && (el == 0 || el == 3 || el == 1);
  const bool inc3 = (ndim > 2) && (el == 0 || el == 1 || el == 2);
  const double v = (el == 0)   ? c.dx[0] * c.dx[1] * c.dx[2]        // Volume<el>
                   : (el == 1) ? c.dx[1] * c.dx[2]
                   : (el == 2) ? c.dx[0] * c.dx[2]
                               : c.dx[0] * c.dx[1];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x;; t += (long)gridDim.x * blockDim.x) {
    int n, ck, cj, ci;
    if (!decode(it, nvar, t, n, ck, cj, ci)) break;
    const int i = (ndim > 0) ? (ci - c.istart[0]) * 2 + c.istart[0] : c.istart[0];
    const int j = (ndim > 1) ? (cj - c.istart[1]) * 2 + c.istart[1] : c.istart[1];
    const int k = (ndim > 2) ? (ck - c.istart[2]) * 2 + c.istart[2] : c.istart[2];
    const double *f = it.fine + (((long)n * it.nf[2] + k) * it.nf[1] + j) * it.nf[0] + i;
    const long sj = it.nf[0], sk = (long)it.nf[0] * it.nf[1];
    double vol[2][2][2], terms[2][2][2];
#pragma unroll
    for (int ok = 0; ok < 2; ++ok)
#pragma unroll
      for (int oj = 0; oj < 2; ++oj)
#pragma unroll
        for (int oi = 0; oi < 2; ++oi) {
          const bool on = (ok == 0 || inc3) && (oj == 0 || inc2) && (oi == 0 || inc1);
          vol[ok][oj][oi] = on ? v : 0.0;
          terms[ok][oj][oi] = on ? vol[ok][oj][oi] * f[ok * sk + oj * sj + oi] : 0.0;
        }
    // pr_ops.hpp:154-161 (off-centred quantities first)
    const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                        ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
    it.coarse[(((long)n * it.nc[2] + ck) * it.nc[1] + cj) * it.nc[0] + ci] =
        (((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
         ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))) /
        tvol;
  }
}

__device__ __forceinline__ double sign_of(double x) { return (x < 0.0) ? -1.0 : 1.0; }

// pr_ops.hpp:94-100
__device__ __forceinline__ double grad_minmod(double fc, double fm, double fp, double dxm,
                                              double dxp) {
  const double gxm = (fc - fm) / dxm;
  const double gxp = (fp - fc) / dxp;
  return 0.5 * (sign_of(gxm) + sign_of(gxp)) * fmin(fabs(gxm), fabs(gxp));
}

// pr_ops.hpp:75-91, cell centred: X<DIM,CC>(i) = xmin + (i + 0.5) dx
__device__ __forceinline__ void spacings(const AmrCoords &c, int d, int i, int fi, double &dxm,
                                         double &dxp, double &dxfm, double &dxfp) {
  const double xm = c.cxmin[d] + ((i - 1) + 0.5) * c.cdx[d];
  const double xc = c.cxmin[d] + (i + 0.5) * c.cdx[d];
  const double xp = c.cxmin[d] + ((i + 1) + 0.5) * c.cdx[d];
  dxm = xc - xm;
  dxp = xp - xc;
  const double fxm = c.xmin[d] + (fi + 0.5) * c.dx[d];
  const double fxp = c.xmin[d] + ((fi + 1) + 0.5) * c.dx[d];
  dxfm = xc - fxm;
  dxfp = fxp - xc;
}

__global__ void __launch_bounds__(kThreads)
k_prolongate_minmod(int ndim, int ng, int nvar, const apk_amr_item_t *__restrict__ items) {
  const apk_amr_item_t it = items[blockIdx.y];
  const AmrCoords c = make_coords(ndim, ng, it);
  const long csj = it.nc[0], csk = (long)it.nc[0] * it.nc[1];
  const long fsj = it.nf[0], fsk = (long)it.nf[0] * it.nf[1];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x;; t += (long)gridDim.x * blockDim.x) {
    int n, k, j, i;
    if (!decode(it, nvar, t, n, k, j, i)) break;
    const int fi = (ndim > 0) ? (i - c.istart[0]) * 2 + c.istart[0] : c.istart[0];
    const int fj = (ndim > 1) ? (j - c.istart[1]) * 2 + c.istart[1] : c.istart[1];
    const int fk = (ndim > 2) ? (k - c.istart[2]) * 2 + c.istart[2] : c.istart[2];
    const double *q = it.coarse + (((long)n * it.nc[2] + k) * it.nc[1] + j) * it.nc[0] + i;
    const double fc = q[0];
    double dx1fm = 0, dx1fp = 0, gx1c = 0, dx2fm = 0, dx2fp = 0, gx2c = 0, dx3fm = 0, dx3fp = 0,
           gx3c = 0, dm, dp;
    spacings(c, 0, i, fi, dm, dp, dx1fm, dx1fp);
    gx1c = grad_minmod(fc, q[-1], q[1], dm, dp);
    if (ndim > 1) {
      spacings(c, 1, j, fj, dm, dp, dx2fm, dx2fp);
      gx2c = grad_minmod(fc, q[-csj], q[csj], dm, dp);
    }
    if (ndim > 2) {
      spacings(c, 2, k, fk, dm, dp, dx3fm, dx3fp);
      gx3c = grad_minmod(fc, q[-csk], q[csk], dm, dp);
    }
    // custom_ops.hpp:121-157
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
#pragma unroll
        for (int ioff = -1; ioff <= 1; ioff++) {
          const double x = q[koff * csk + joff * csj + ioff];
          qmin = fmin(qmin, x);
          qmax = fmax(qmax, x);
        }
    double alpha = 1.0;
    if (dqmax * alpha > (qmax - fc)) alpha = (qmax - fc) / dqmax;
    if (dqmax * alpha > (fc - qmin)) alpha = (fc - qmin) / dqmax;
    gx1c *= alpha;
    gx2c *= alpha;
    gx3c *= alpha;
    double *f = it.fine + (((long)n * it.nf[2] + fk) * it.nf[1] + fj) * it.nf[0] + fi;
    // custom_ops.hpp:159-181
    f[0] = fc - (gx1c * dx1fm + gx2c * dx2fm + gx3c * dx3fm);
    f[1] = fc + (gx1c * dx1fp - gx2c * dx2fm - gx3c * dx3fm);
    if (ndim > 1) {
      f[fsj] = fc - (gx1c * dx1fm - gx2c * dx2fp + gx3c * dx3fm);
      f[fsj + 1] = fc + (gx1c * dx1fp + gx2c * dx2fp - gx3c * dx3fm);
    }
    if (ndim > 2) {
      f[fsk] = fc - (gx1c * dx1fm + gx2c * dx2fm - gx3c * dx3fp);
      f[fsk + 1] = fc + (gx1c * dx1fp - gx2c * dx2fm + gx3c * dx3fp);
      f[fsk + fsj] = fc - (gx1c * dx1fm - gx2c * dx2fp - gx3c * dx3fp);
      f[fsk + fsj + 1] = fc + (gx1c * dx1fp + gx2c * dx2fp + gx3c * dx3fp);
    }
  }
}

// Refinement criteria of src/refinement: per block the maximum over a criterion-specific range of
//   CRIT 0  |grad p| / p, central differences without dx, interior +- 1 (gradient.cpp:18-61)
//   CRIT 1  sqrt((dv1/dx2)^2 + (dv2/dx1)^2) of the x1-x2 velocities, undivided central differences,
//           interior +- 1 in x1 and x2 (gradient.cpp:64-95)
//   CRIT 2  density, interior plus the first ghost cell beyond ie (other.cpp:18-44: `ib.e + 1`)
template <int CRIT>
__global__ void __launch_bounds__(kThreads)
k_refinement_measure(ApkGeom g, const apk_block_t *__restrict__ blocks,
                     unsigned long long *measure_bits) {
  const apk_block_t &B = blocks[blockIdx.y];
  const int lo1 = (CRIT == 2) ? 0 : -1, hi1 = 1;                      // x1: is+lo1 .. ie+hi1
  const int pad2 = (CRIT == 2) ? 0 : 1;                                // x2: js-pad2 .. je+pad2
  const int pad3 = (CRIT == 0 && g.ndim == 3) ? 1 : 0;
  const int e1 = g.ie - g.is + 1 + hi1 - lo1, e2 = g.je - g.js + 1 + 2 * pad2,
            e3 = g.ke - g.ks + 1 + 2 * pad3;
  const long total = (long)e1 * e2 * e3;
  const long sj = g.n1, sk = (long)g.n1 * g.n2;
  double m = 0.0;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(t / ((long)e1 * e2));
    const long r = t - (long)kk * e1 * e2;
    const int jj = (int)(r / e1), ii = (int)(r - (long)jj * e1);
    const int k = g.ks - pad3 + kk, j = g.js - pad2 + jj, i = g.is + lo1 + ii;
    const long c = (long)k * sk + (long)j * sj + i;
    if (CRIT == 0) {
      const double *p = B.prim + 4 * g.ncell + c;  // IPR
      const double a = 0.5 * (p[1] - p[-1]), b = 0.5 * (p[sj] - p[-sj]);
      double s = a * a + b * b;
      if (g.ndim == 3) {
        const double cc = 0.5 * (p[sk] - p[-sk]);
        s = s + cc * cc;
      }
      m = fmax(m, sqrt(s) / p[0]);
    } else if (CRIT == 1) {
      const double *v1 = B.prim + 1 * g.ncell + c, *v2 = B.prim + 2 * g.ncell + c;
      const double vgy = fabs(v2[1] - v2[-1]) * 0.5;
      const double vgx = fabs(v1[sj] - v1[-sj]) * 0.5;
      const double vg = sqrt(vgx * vgx + vgy * vgy);
      if (vg > m) m = vg;
    } else {
      m = fmax(m, B.prim[c]);  // IDN
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  // measure >= 0: the IEEE bit pattern is monotone, atomicMax on the integer image is exact
  if ((threadIdx.x & 31) == 0 && m > 0.0)
    atomicMax(measure_bits + blockIdx.y, (unsigned long long)__double_as_longlong(m));
}

__global__ void k_refinement_tags(int nb, double refine_above, double derefine_below,
                                  const double *measure, int *tags) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const double m = measure[b];
  tags[b] = (m > refine_above) ? 1 : ((m < derefine_below) ? -1 : 0);
}

// ---- history sums (hydro.cpp:145-208).  Level 1: one CTA per (block, slice) reduces its cells in
// a fixed order (thread-strided partial sums -> shuffle tree -> shared-memory tree); level 2: one
// CTA adds the partials in index order.  No atomics: bitwise reproducible.
constexpr int kHistSlices = 8;

template <int FLUID>
__global__ void __launch_bounds__(kThreads)
k_history_partial(ApkGeom g, const apk_block_t *__restrict__ blocks, double *__restrict__ partial) {
  const apk_block_t &B = blocks[blockIdx.y];
  const double *__restrict__ u = B.cons;
  const int e1 = g.ie - g.is + 1, e2 = g.je - g.js + 1, e3 = g.ke - g.ks + 1;
  const long total = (long)e1 * e2 * e3;
  const long per = (total + kHistSlices - 1) / kHistSlices;
  const long lo = (long)blockIdx.x * per, hi = min(lo + per, total);
  const long sj = g.n1, sk = (long)g.n1 * g.n2;
  const double vol = B.dx[0] * B.dx[1] * B.dx[2];
  const double diag = sqrt(B.dx[0] * B.dx[0] + B.dx[1] * B.dx[1] + B.dx[2] * B.dx[2]);
  double s[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) s[q] = 0.0;
  for (long t = lo + threadIdx.x; t < hi; t += blockDim.x) {
    const int kk = (int)(t / ((long)e1 * e2));
    const long r = t - (long)kk * e1 * e2;
    const int jj = (int)(r / e1), ii = (int)(r - (long)jj * e1);
    const long c = (long)(g.ks + kk) * sk + (long)(g.js + jj) * sj + (g.is + ii);
    const double d = u[c], m1 = u[g.ncell + c], m2 = u[2 * g.ncell + c], m3 = u[3 * g.ncell + c];
    s[0] += d * vol;
    s[1] += m1 * vol;
    s[2] += m2 * vol;
    s[3] += m3 * vol;
    s[4] += 0.5 / d * (m1 * m1 + m2 * m2 + m3 * m3) * vol;
    s[5] += u[4 * g.ncell + c] * vol;
    if (FLUID == APK_FLUID_GLMMHD) {  // Hst::emag sums in any dimensionality (hydro.cpp:179-183)
      const double *b1 = u + 5 * g.ncell + c, *b2 = u + 6 * g.ncell + c, *b3 = u + 7 * g.ncell + c;
      s[6] += 0.5 * (b1[0] * b1[0] + b2[0] * b2[0] + b3[0] * b3[0]) * vol;
    }
    if (FLUID == APK_FLUID_GLMMHD && g.ndim >= 2) {  // the j +- 1 neighbours exist from 2-D on
      const double *b1 = u + 5 * g.ncell + c, *b2 = u + 6 * g.ncell + c, *b3 = u + 7 * g.ncell + c;
      double divb = (b1[1] - b1[-1]) / B.dx[0] + (b2[sj] - b2[-sj]) / B.dx[1];
      if (g.ndim == 3) divb += (b3[sk] - b3[-sk]) / B.dx[2];
      const double abs_b = sqrt(b1[0] * b1[0] + b2[0] * b2[0] + b3[0] * b3[0]);
      s[7] += (abs_b != 0) ? 0.5 * diag * fabs(divb) / abs_b * vol : 0;
    }
  }
  __shared__ double sh[8][kThreads / 32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    double v = s[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) v += sh[threadIdx.x][w];
    partial[((long)blockIdx.y * kHistSlices + blockIdx.x) * 8 + threadIdx.x] = v;
  }
}

__global__ void k_history_final(int npartial, const double *__restrict__ partial,
                                double *__restrict__ out) {
  // 8 warps, one per quantity: lanes take strided partials in a fixed order, then a shuffle tree
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = 0.0;
  for (int p = lane; p < npartial; p += 32) v += partial[(long)p * 8 + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) out[q] = v;
}
}  // namespace

size_t history_workspace_bytes(const apk_pack_desc_t &d) {
  return ((size_t)d.nblocks * kHistSlices + 1) * 8 * sizeof(double);
}

int launch_history(const apk_pack_desc_t &d, const apk_block_t *blocks, double *work,
                   cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  double *out = work + (size_t)g.nb * kHistSlices * 8;
  dim3 grid(kHistSlices, g.nb);
  if (d.fluid == APK_FLUID_GLMMHD)
    k_history_partial<APK_FLUID_GLMMHD><<<grid, kThreads, 0, s>>>(g, blocks, work);
  else
    k_history_partial<APK_FLUID_EULER><<<grid, kThreads, 0, s>>>(g, blocks, work);
  k_history_final<<<1, 256, 0, s>>>(g.nb * kHistSlices, work, out);
  apk_common::count_launch(2);
  return 0;
}

int launch_restrict_average(int ndim, int ng, int nvar, int el, const apk_amr_item_t *items,
                            int nitems, cudaStream_t s) {
  if (nitems <= 0) return 0;
  // items of the multilevel exchange are small (a face region of an 8^3 block: 4 x 4 x 2 coarse
  // cells x nvar): two CTAs per item loop over it.  (64 CTAs per item made the exchange of a
  // 600-block mesh launch ~600 k nearly empty CTAs: 1.2 ms per exchange, measured.)
  dim3 grid(2, nitems);
  k_restrict_average<<<grid, kThreads, 0, s>>>(ndim, ng, nvar, el, items);
  apk_common::count_launch();
  return 0;
}

int launch_prolongate_minmod(int ndim, int ng, int nvar, const apk_amr_item_t *items, int nitems,
                             cudaStream_t s) {
  if (nitems <= 0) return 0;
  dim3 grid(2, nitems);
  k_prolongate_minmod<<<grid, kThreads, 0, s>>>(ndim, ng, nvar, items);
  apk_common::count_launch();
  return 0;
}

int launch_refinement_tags(const apk_pack_desc_t &d, const apk_block_t *blocks, int criterion,
                           double refine_above, double derefine_below, int *tags, double *measure,
                           cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  cudaMemsetAsync(measure, 0, sizeof(double) * g.nb, s);
  // gradient.cpp:58-60: 1-D meshes are never tagged by the gradient criteria
  if (g.ndim >= 2 || criterion == 2) {
    dim3 grid(4, g.nb);
    unsigned long long *bits = reinterpret_cast<unsigned long long *>(measure);
    if (criterion == 0) k_refinement_measure<0><<<grid, kThreads, 0, s>>>(g, blocks, bits);
    else if (criterion == 1) k_refinement_measure<1><<<grid, kThreads, 0, s>>>(g, blocks, bits);
    else k_refinement_measure<2><<<grid, kThreads, 0, s>>>(g, blocks, bits);
    k_refinement_tags<<<(g.nb + 127) / 128, 128, 0, s>>>(g.nb, refine_above, derefine_below, measure,
                                                          tags);
    apk_common::count_launch(2);
  } else {
    cudaMemsetAsync(tags, 0, sizeof(int) * g.nb, s);
  }
  return 0;
}

}  // namespace APK_NS
