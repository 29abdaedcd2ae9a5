// apk_rows.cu -- the row-task form of the fused stage kernel (2-D and 3-D meshes; replaces
// K1-K3, K6, K7, K10, K11 and the ghost exchange of SURVEY.md 2.3 with one launch per stage).
//
// Same contract and the same tile decomposition as k_stage_fused (apk_fused.cu: a CTA of 4 warps
// owns a 32 x 4 column of one meshblock and marches along x3), but the work inside a plane is cut
// so that reconstructed states never leave the registers of the warp that consumes them and every
// PPM interface value is computed ONCE:
//   * warp w owns tile row w.  "row passes" (one code path, run for x3 and then x1) reconstruct
//     all variables of the row along the pass direction, solve the Riemann problem of the row's
//     faces straight from registers and
//       x3: keep the flux in registers, FINISH the cells of the previous plane (flux divergence in
//           the reference's association, low-storage update, Dedner damping, cons->prim, dt
//           reduction, ghost push); the left state and the PPM interface value needed by the next
//           plane wait in thread-private shared-memory slots (face k+1 is never recomputed),
//       x1: left states come from the neighbouring lane (warp shuffle); fluxes go to shared memory.
//   * x2 needs data of other rows: per variable one warp marches over rows -1..TY with a rolling
//     register window (each value is loaded once instead of five times), writing L/R states to
//     shared memory; after a barrier the x2 face rows (and the one x1 face per row that has no
//     lane, face 32) are solved; after the second barrier the next plane starts.
//   * PPM (ppm_simple.hpp:38-162): the limited 4th-order interface value between two cells is the
//     same expression for both of them (apk_math.cuh, ppm_ifaces), so a lane evaluates one
//     interface and obtains the other from its neighbour lane (x1), from the previous plane (x3)
//     or from the previous row (x2).  Per cell and direction that is 1 interface + 1 cell part
//     instead of 2 + 1, and the x2 sweep costs TY+3 interfaces + TY+2 cell parts per TY rows
//     instead of 3 (TY+2) parts.
// Results are bit-identical to k_stage_fused in APK_FP_STRICT mode (same expression trees).
// Compiled twice (APK_NS = apk_fast / apk_strict), see apk_launch.h.
#include <cstdint>
#include <cstdlib>
#include <algorithm>

#include "apk_launch.h"
#include "apk_math.cuh"

namespace APK_NS {
namespace {

#ifndef APK_ROWS_MINBLOCKS
#define APK_ROWS_MINBLOCKS 3
#endif
#ifndef APK_ROWS_PREFETCH
#define APK_ROWS_PREFETCH 1
#endif

constexpr int RTX = 32, RTY = 4, RNW = 4, RNT = RNW * 32;
constexpr unsigned FULL = 0xffffffffu;

template <int NH, bool PPM>
struct RowsSmem {
  double wly[NH][RTY + 1][RTX];  // x2: L states of face rows 0..TY (slot order of the x2 sweep)
  double wry[NH][RTY + 1][RTX];  // x2: R states
  double fy[NH][RTY + 1][RTX];   // x2: fluxes
  double fx[NH][RTY][RTX + 1];   // x1: fluxes of faces 0..TX
  double xel[NH][RTY];           // x1: L / R state of the face behind the last lane, per row
  double xer[NH][RTY];
  double lz[NH][RTY][RTX];       // x3: L state of the next face (thread-private slot)
  double iz[PPM ? NH : 1][RTY][RTX];  // x3: PPM interface value shared with the next plane
  double *push_prim[27], *push_cons[27];
};

struct RowsArgs {
  ApkGeom g;
  const apk_block_t *blocks;
  double *const *prim_out;
  double g0, g1, bdt, dedner;
  RsPars rsp;
  EosPars eos;
  int ntx, flags, nchunk;
  unsigned long long *dtmin;
  const apk_push_dst_t *push;
  int *status;
  unsigned long long *fofc;
};

APK_DEV double shfl_d(double x, int src) { return __shfl_sync(FULL, x, src); }
APK_DEV double shfl_up_d(double x) { return __shfl_up_sync(FULL, x, 1); }
APK_DEV double shfl_down_d(double x) { return __shfl_down_sync(FULL, x, 1); }

// Reconstruction of U consecutive variable slots N0.. of one tile row along `dir` (0: x1, lanes
// are neighbours along the sweep; 2: x3, the march direction).  pc points at the lane's cell of
// variable 0; every lane of the warp calls (clamped lanes compute valid duplicates).
template <int NH, int RECON, int U, int N0>
APK_DEV void row_recon_group(const double *__restrict__ pc, long ncell, long stride, int dir,
                             double dxd, int lane, int last, double e_ilo, double *iz_slot,
                             double (&ql)[NH], double (&qr)[NH]) {
  double v[U][5];
#pragma unroll
  for (int u = 0; u < U; ++u)
    load_stencil<RECON>(pc + (long)var_of_slot(N0 + u, dir) * ncell, stride, v[u]);
  if constexpr (RECON == APK_RECON_PPM) {
    // x1: the lane evaluates its LOWER interface (cells i-2..i+1), the upper one is the lower one
    //     of lane+1 (of the cell behind the tile for the last lane: e_ilo of the edge pass);
    // x3: the lane evaluates the UPPER interface (planes k-1..k+2), the lower one was the upper
    //     one of the previous plane (iz_slot)
    double s4[U][4], inew[U], flo[U], fhi[U], qlg[U], qrg[U];
    const int o = (dir == 0) ? 0 : 1;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      s4[u][0] = o ? v[u][1] : v[u][0];
      s4[u][1] = o ? v[u][2] : v[u][1];
      s4[u][2] = o ? v[u][3] : v[u][2];
      s4[u][3] = o ? v[u][4] : v[u][3];
    }
    ppm_ifaces<U>(s4, inew);
    if (dir == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double up = shfl_down_d(inew[u]);
        const double ed = shfl_d(e_ilo, 2 * (N0 + u) + 1);
        flo[u] = inew[u];
        fhi[u] = (lane == last) ? ed : up;
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double *slot = iz_slot + (N0 + u) * (RTY * RTX);
        flo[u] = *slot;
        fhi[u] = inew[u];
        *slot = inew[u];
      }
    }
    ppm_cells<U>(v, flo, fhi, qlg, qrg);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ql[N0 + u] = qlg[u];
      qr[N0 + u] = qrg[u];
    }
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u)
      recon_stencil<RECON>(v[u], dxd, (N0 + u == 0 || N0 + u == 4), ql[N0 + u], qr[N0 + u]);
  }
}

template <int NH, int RECON>
APK_DEV void row_recon(const double *__restrict__ pc, long ncell, long stride, int dir, double dxd,
                       int lane, int last, double e_ilo, double *iz_slot, double (&ql)[NH],
                       double (&qr)[NH]) {
  row_recon_group<NH, RECON, 3, 0>(pc, ncell, stride, dir, dxd, lane, last, e_ilo, iz_slot, ql, qr);
  if constexpr (NH == 9) {
    row_recon_group<NH, RECON, 3, 3>(pc, ncell, stride, dir, dxd, lane, last, e_ilo, iz_slot, ql, qr);
    row_recon_group<NH, RECON, 3, 6>(pc, ncell, stride, dir, dxd, lane, last, e_ilo, iz_slot, ql, qr);
  } else {
    row_recon_group<NH, RECON, 2, 3>(pc, ncell, stride, dir, dxd, lane, last, e_ilo, iz_slot, ql, qr);
  }
}

// x1 edge pass: the two cells next to the tile row (cell -1 and cell txe) for all variables in
// one warp pass, lane = 2 * variable + side.  e_ql: L state of face 0 (from cell -1); e_qr: R
// state of face txe (from cell txe); e_ilo (PPM): lower interface of the lane's cell -- for
// side 1 that is the upper interface of the row's last cell.  prow points at cell 0 of the row.
template <int NH, int RECON>
APK_DEV void row_edge(const double *__restrict__ prow, long ncell, int txe, double dx, int lane,
                      double &e_ql, double &e_qr, double &e_ilo) {
  const int ne = min(lane >> 1, NH - 1), side = lane & 1;
  const double *q = prow + (long)ne * ncell + (side ? txe : -1);  // var_of_slot(n, 0) == n
  double v[1][5];
  load_stencil<RECON>(q, 1L, v[0]);
  e_ilo = 0.0;
  if constexpr (RECON == APK_RECON_PPM) {
    double s4[2][4], f2[2], flo[1], fhi[1], qlg[1], qrg[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s4[0][k] = v[0][k];
      s4[1][k] = v[0][k + 1];
    }
    ppm_ifaces<2>(s4, f2);
    flo[0] = f2[0];
    fhi[0] = f2[1];
    ppm_cells<1>(v, flo, fhi, qlg, qrg);
    e_ql = qlg[0];
    e_qr = qrg[0];
    e_ilo = f2[0];
  } else {
    recon_stencil<RECON>(v[0], dx, (ne == 0 || ne == 4), e_ql, e_qr);
  }
}

// x2 column task: U variable slots ns[] of the x2 sweep, lanes along x1, rows r = -1..tye with a
// rolling register window; writes L states of face rows 1..tye... (ql of row r is the L state of
// face r+1, qr of row r the R state of face r).  pcol points at the lane's cell of tile row 0 of
// memory variable 0.
template <int NH, int RECON, bool PPMS, int U>
APK_DEV void col_task(const double *__restrict__ pcol, long ncell, long sj, int tye, double dx2,
                      int lane, bool lane_ok, const int (&ns)[U], RowsSmem<NH, PPMS> &sm) {
  constexpr int R = ReconTraits<RECON>::radius;
  double w[U][5], flo[U];
  const double *q[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    q[u] = pcol + (long)var_of_slot(ns[u], 1) * ncell;
#pragma unroll
    for (int k = 0; k < 5; ++k) w[u][k] = 0.0;
#pragma unroll
    for (int d = -R; d <= R; ++d) w[u][2 + d] = q[u][(long)(-1 + d) * sj];
    flo[u] = 0.0;
  }
  if constexpr (RECON == APK_RECON_PPM) {
    double s4[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k) s4[u][k] = w[u][k];
    ppm_ifaces<U>(s4, flo);
  }
#pragma unroll 1
  for (int r = -1; r <= tye; ++r) {
    double ql[U], qr[U];
    if constexpr (RECON == APK_RECON_PPM) {
      double s4[U][4], fhi[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) s4[u][k] = w[u][k + 1];
      ppm_ifaces<U>(s4, fhi);
      ppm_cells<U>(w, flo, fhi, ql, qr);
#pragma unroll
      for (int u = 0; u < U; ++u) flo[u] = fhi[u];
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        recon_stencil<RECON>(w[u], dx2, (ns[u] == 0 || ns[u] == 4), ql[u], qr[u]);
    }
    if (lane_ok) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r < tye) sm.wly[ns[u]][r + 1][lane] = ql[u];
        if (r >= 0) sm.wry[ns[u]][r][lane] = qr[u];
      }
    }
    if (r < tye) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) w[u][k] = w[u][k + 1];
        w[u][2 + R] = q[u][(long)(r + 1 + R) * sj];
      }
    }
  }
}

template <int FLUID, int RECON, int RS, int NDIM>
__global__ void __launch_bounds__(RNT, APK_ROWS_MINBLOCKS)
k_stage_rows(const __grid_constant__ RowsArgs a) {
  constexpr int NH = (FLUID == APK_FLUID_GLMMHD) ? 9 : 5;
  constexpr bool PPMS = (RECON == APK_RECON_PPM);
  constexpr int R = ReconTraits<RECON>::radius;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SmemT = RowsSmem<NH, PPMS>;
  SmemT &sm = *reinterpret_cast<SmemT *>(smem_raw);
  const ApkGeom &g = a.g;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const apk_block_t &B = a.blocks[blockIdx.y];
  const int tix = blockIdx.x % a.ntx, tiy = blockIdx.x / a.ntx;
  const int i0 = g.is + tix * RTX, j0 = g.js + tiy * RTY;
  const int txe = min(RTX, g.ie + 1 - i0);  // valid cells of the tile in x1
  const int tye = min(RTY, g.je + 1 - j0);  // ... in x2
  const long sj = g.n1, sk = (long)g.n1 * g.n2;
  const long ncell = g.ncell;
  const double *__restrict__ prim = B.prim;
  double *__restrict__ cons = B.cons;
  double *__restrict__ cons_u1 = B.cons_u1;
  const int flags = a.flags;
  const bool save_u1 = (flags & APK_FUSED_SAVE_U1) != 0;
  const bool prim_only = (flags & APK_FUSED_PRIM_ONLY) != 0;
  const bool u1_is_cons = (flags & APK_FUSED_U1_FROM_CONS) != 0;
  double *__restrict__ po = a.prim_out[blockIdx.y];
  const double dx1 = B.dx[0], dx2 = B.dx[1], dx3 = B.dx[2];
  // uniform_coordinates.hpp:37-40
  const double a1 = dx2 * dx3, a2 = dx1 * dx3, a3 = dx1 * dx2, vol = dx1 * dx2 * dx3;
#ifndef APK_STRICT
  const double ivol = frcp(vol);
#endif
  const double g0 = a.g0, g1 = a.g1, bdt = a.bdt;

  // the warp's tile row and the thread's cell
  const int rj = warp, ci = lane;
  const bool row_ok = rj < tye, lane_ok = lane < txe;
  const bool owner = row_ok && lane_ok;
  const int rjc = min(rj, tye - 1), lc = min(lane, txe - 1), last = txe - 1;

  const int nx3c = g.ke - g.ks + 1;
  const int kbeg = (NDIM == 3) ? g.ks + ((int)blockIdx.z * nx3c) / a.nchunk : g.ks;
  const int kend = (NDIM == 3) ? g.ks + (((int)blockIdx.z + 1) * nx3c) / a.nchunk - 1 : g.ke;
  const int kfirst = (NDIM == 3) ? kbeg - 1 : kbeg;
  const int klast = (NDIM == 3) ? kend + 1 : kend;

  if (flags & (APK_FUSED_PUSH_PRIM | APK_FUSED_PUSH_CONS)) {
    if (threadIdx.x < 27) {
      double *pp = nullptr, *pcn = nullptr;
      if (a.push) {  // explicit destinations (may be peer-GPU memory: halo exchange as NVLink stores)
        pp = a.push[blockIdx.y * 27 + threadIdx.x].prim;
        pcn = a.push[blockIdx.y * 27 + threadIdx.x].cons;
      } else {
        const int nb = B.nbr[threadIdx.x];
        if (nb >= 0) {
          pp = a.prim_out[nb];
          pcn = a.blocks[nb].cons;
        }
      }
      sm.push_prim[threadIdx.x] = (flags & APK_FUSED_PUSH_PRIM) ? pp : nullptr;
      sm.push_cons[threadIdx.x] = (flags & APK_FUSED_PUSH_CONS) ? pcn : nullptr;
    }
    __syncthreads();
  }

  double fzlo[NH];
#pragma unroll
  for (int v = 0; v < NH; ++v) fzlo[v] = 0.0;
  double my_min = DBL_MAX;

  // tile-relative base of the lane's (clamped) cell in plane 0, variable 0
  const long cell0 = (long)(j0 + rjc) * sj + (i0 + lc);
  double *const lz_slot = &sm.lz[0][rj][lane];
  double *const iz_slot = &sm.iz[0][rj][lane];

  if constexpr (NDIM == 3 && PPMS) {
    // prime the carried x3 interface: lower interface of plane kfirst (planes kfirst-2..kfirst+1)
    const double *pc = prim + (long)kfirst * sk + cell0;
#pragma unroll 1
    for (int n0 = 0; n0 < NH; n0 += 3) {
      double s4[3][4], f3[3];
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int n = min(n0 + u, NH - 1);
        const double *q = pc + (long)var_of_slot(n, 2) * ncell;
        s4[u][0] = q[-2 * sk];
        s4[u][1] = q[-sk];
        s4[u][2] = q[0];
        s4[u][3] = q[sk];
      }
      ppm_ifaces<3>(s4, f3);
#pragma unroll
      for (int u = 0; u < 3; ++u)
        if (n0 + u < NH) iz_slot[(n0 + u) * (RTY * RTX)] = f3[u];
    }
  }

  // ---- finishing a cell: flux divergence, update, Dedner, cons->prim, stores, push, dt ----------
  auto finalize = [&](int kfin, const double (&fz_hi)[NH]) {
    const long c = (long)kfin * sk + (long)(j0 + rj) * sj + (i0 + ci);
    double u[NH], w[NH], c0[NH], c1[NH];
    const bool src_cons = save_u1 || prim_only || u1_is_cons;  // u1 term comes from cons
    const bool need_u0 = src_cons || (g0 != 0.0);
    double cc[NH];
#pragma unroll
    for (int v = 0; v < NH; ++v) cc[v] = need_u0 ? cons[(long)v * ncell + c] : 0.0;
#pragma unroll
    for (int v = 0; v < NH; ++v) c0[v] = (g0 != 0.0 || save_u1 || prim_only) ? cc[v] : 0.0;
    if (src_cons) {
#pragma unroll
      for (int v = 0; v < NH; ++v) {
        c1[v] = cc[v];
        if (save_u1 && !prim_only) cons_u1[(long)v * ncell + c] = cc[v];
      }
    } else {
#pragma unroll
      for (int v = 0; v < NH; ++v) c1[v] = cons_u1[(long)v * ncell + c];
    }
#pragma unroll
    for (int v = 0; v < NH; ++v) {
      const int sx = slot_of_var(v, 0), sy = slot_of_var(v, 1);
      double d = (a1 * sm.fx[sx][rj][ci + 1] - a1 * sm.fx[sx][rj][ci]);
      d += (a2 * sm.fy[sy][rj + 1][ci] - a2 * sm.fy[sy][rj][ci]);
      if constexpr (NDIM == 3) {
        const int sz = slot_of_var(v, 2);
        d += (a3 * fz_hi[sz] - a3 * fzlo[sz]);
      }
      // update.cpp:135-137 / update.hpp:55
#ifdef APK_STRICT
      u[v] = g0 * c0[v] + g1 * c1[v] + bdt * fdiv(-d, vol);
#else
      u[v] = g0 * c0[v] + g1 * c1[v] + bdt * (-d * ivol);  // 1/V once per thread (<= 1 ulp)
#endif
    }
    bool fofc_bad = false;
    if (flags & APK_FUSED_FOFC_CHECK) {
      // first attempt of FirstOrderFluxCorrect on this very trial update (hydro.cpp:1287-1306)
      fofc_bad = fofc_needs_correction<NH>(u);
      if (fofc_bad) atomicAdd(a.fofc, 1ULL);
    }
    if constexpr (FLUID == APK_FLUID_GLMMHD) u[8] *= a.dedner;  // dedner_source.cpp:72
    const int err = cons_to_prim<NH>(u, w, a.eos);
    // PARTHENON_REQUIRE of ConsToPrim (a cell FOFC flagged is redone by the caller instead)
    if (err && a.status && !fofc_bad) atomicMax(a.status, err);
#pragma unroll
    for (int v = 0; v < NH; ++v) {
      const long o = (long)v * ncell + c;
      if (!prim_only) cons[o] = u[v];
      po[o] = w[v];
    }
    if (flags & (APK_FUSED_PUSH_PRIM | APK_FUSED_PUSH_CONS)) {
      // ghost exchange folded into the epilogue ("push", sender side of bnd_info.cpp:136-230):
      // same as k_stage_fused
      const int gi = i0 + ci, gj = j0 + rj;
      const bool lo1 = gi < g.is + g.ng, hi1 = gi > g.ie - g.ng;
      const bool lo2 = gj < g.js + g.ng, hi2 = gj > g.je - g.ng;
      const bool lo3 = NDIM == 3 && kfin < g.ks + g.ng, hi3 = NDIM == 3 && kfin > g.ke - g.ng;
      if (lo1 | hi1 | lo2 | hi2 | lo3 | hi3) {
        const long e1 = g.ie - g.is + 1, e2 = (long)(g.je - g.js + 1) * sj,
                   e3 = (long)(g.ke - g.ks + 1) * sk;
#pragma unroll 1
        for (int o3 = (NDIM == 3 ? -1 : 0); o3 <= (NDIM == 3 ? 1 : 0); ++o3) {
          if ((o3 < 0 && !lo3) || (o3 > 0 && !hi3)) continue;
#pragma unroll 1
          for (int o2 = -1; o2 <= 1; ++o2) {
            if ((o2 < 0 && !lo2) || (o2 > 0 && !hi2)) continue;
#pragma unroll 1
            for (int o1 = -1; o1 <= 1; ++o1) {
              if ((o1 < 0 && !lo1) || (o1 > 0 && !hi1) || (o1 | o2 | o3) == 0) continue;
              const int o = (o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1);
              const long cd = c - (o3 * e3 + o2 * e2 + o1 * e1);
              double *__restrict__ pd = sm.push_prim[o];
              double *__restrict__ cdst = sm.push_cons[o];
              if (pd) {
#pragma unroll
                for (int v = 0; v < NH; ++v) pd[(long)v * ncell + cd] = w[v];
              }
              if (cdst) {
#pragma unroll
                for (int v = 0; v < NH; ++v) cdst[(long)v * ncell + cd] = u[v];
              }
            }
          }
        }
      }
    }
    if (a.dtmin) {  // hydro.cpp:845-895
      double lx, ly, lz;
      if constexpr (FLUID == APK_FLUID_EULER) {
        lx = ly = lz = sound_speed(a.rsp.gamma, w[0], w[4]);
      } else {
        lx = fast_speed(a.rsp.gamma, w[0], w[4], w[5], w[6], w[7]);
        ly = fast_speed(a.rsp.gamma, w[0], w[4], w[6], w[7], w[5]);
        lz = (NDIM > 2) ? fast_speed(a.rsp.gamma, w[0], w[4], w[7], w[5], w[6]) : 0.0;
      }
      my_min = dmin(my_min, qdiv(dx1, (fabs(w[1]) + lx)));
      my_min = dmin(my_min, qdiv(dx2, (fabs(w[2]) + ly)));
      if (NDIM > 2) my_min = dmin(my_min, qdiv(dx3, (fabs(w[3]) + lz)));
    }
  };

  for (int kp = kfirst; kp <= klast; ++kp) {
    const bool inplane = (kp >= kbeg && kp <= kend);
    const long plane = (long)kp * sk;
    // ------------------------------------------------------------------ phase A: row passes
#pragma unroll 1
    for (int pass = (NDIM == 3) ? 0 : 1; pass < 2; ++pass) {
      const int dir = pass ? 0 : 2;
      if (pass == 1 && !inplane) break;
      if (!row_ok) continue;  // warp-uniform: rows beyond a ragged tile
      const double *pc = prim + plane + cell0;
      double e_ql = 0.0, e_qr = 0.0, e_ilo = 0.0;
      if (dir == 0)
        row_edge<NH, RECON>(prim + plane + (long)(j0 + rj) * sj + i0, ncell, txe, dx1, lane, e_ql,
                            e_qr, e_ilo);
      double ql[NH], qr[NH];
      row_recon<NH, RECON>(pc, ncell, dir == 0 ? 1L : sk, dir, dir == 0 ? dx1 : dx3, lane, last,
                           e_ilo, iz_slot, ql, qr);
      double L[NH], F[NH];
      if (dir == 0) {
#pragma unroll
        for (int n = 0; n < NH; ++n) {
          const double up = shfl_up_d(ql[n]);
          const double ed = shfl_d(e_ql, 2 * n);
          L[n] = (lane == 0) ? ed : up;
          if (lane == last) sm.xel[n][rj] = ql[n];
        }
        if (lane < 2 * NH && (lane & 1)) sm.xer[lane >> 1][rj] = e_qr;
      } else {
#pragma unroll
        for (int n = 0; n < NH; ++n) {
          L[n] = lz_slot[n * (RTY * RTX)];
          lz_slot[n * (RTY * RTX)] = ql[n];
        }
      }
      if (dir == 0 || kp >= kbeg) {
        riemann_solve<FLUID, RS>(L, qr, a.rsp, F);
        if (dir == 0) {
          if (lane_ok) {
#pragma unroll
            for (int n = 0; n < NH; ++n) sm.fx[n][rj][lane] = F[n];
          }
        } else {
          // the cells of plane kp-1 now have all six fluxes (fx / fy of plane kp-1 are still in
          // shared memory: this warp overwrites its fx row only in the x1 pass below, fy is
          // rewritten after the next barrier)
          if (lane_ok && kp - 1 >= kbeg) finalize(kp - 1, F);
#pragma unroll
          for (int n = 0; n < NH; ++n) fzlo[n] = F[n];
          __syncwarp();
        }
      }
    }
    if (inplane) {
      // ---------------------------------------------------------------- phase A: x2 columns
      const double *pcol = prim + plane + (long)j0 * sj + (i0 + lc);
      if constexpr (NH == 9) {
        if (warp == 0) {
          const int ns[3] = {0, 4, 8};
          col_task<NH, RECON, PPMS, 3>(pcol, ncell, sj, tye, dx2, lane, lane_ok, ns, sm);
        } else {
          const int ns[2] = {warp, warp + 4};
          col_task<NH, RECON, PPMS, 2>(pcol, ncell, sj, tye, dx2, lane, lane_ok, ns, sm);
        }
      } else {
        if (warp == 0) {
          const int ns[2] = {0, 4};
          col_task<NH, RECON, PPMS, 2>(pcol, ncell, sj, tye, dx2, lane, lane_ok, ns, sm);
        } else {
          const int ns[1] = {warp};
          col_task<NH, RECON, PPMS, 1>(pcol, ncell, sj, tye, dx2, lane, lane_ok, ns, sm);
        }
      }
      __syncthreads();
      // ---------------------------------------------------------------- phase B: x2 faces + x1 edge
      if constexpr (APK_ROWS_PREFETCH && NDIM == 3) {
        // pull the lines the next row passes touch for the first time into L1: the new top plane
        // of the x3 stencils (kp+1+R) of this warp's row
        if (row_ok && kp + 1 <= klast) {
          const double *pn = prim + (long)(kp + 1 + R) * sk + (long)(j0 + rj) * sj + i0;
          for (int p = lane; p < NH * 3; p += 32) {
            const int vv = p / 3, seg = p % 3;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pn + (long)vv * ncell +
                                                         min(seg * 16, txe - 1)));
          }
        }
      }
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        // tasks: face rows 0..3 on warps 0..3, then face row 4 on warp 2 and the x1 edge faces
        // (one per tile row, lanes = rows) on warp 3
        const bool edge = (q == 1 && warp == 3);
        const int r = (q == 0) ? warp : RTY;
        if (q == 1 && warp < 2) break;
        if (!edge && r > tye) continue;
        double L[NH], Rr[NH], F[NH];
        const int le = min(lane, tye - 1);
#pragma unroll
        for (int n = 0; n < NH; ++n) {
          L[n] = edge ? sm.xel[n][le] : sm.wly[n][r][lc];
          Rr[n] = edge ? sm.xer[n][le] : sm.wry[n][r][lc];
        }
        riemann_solve<FLUID, RS>(L, Rr, a.rsp, F);
        if (edge) {
          if (lane < tye) {
#pragma unroll
            for (int n = 0; n < NH; ++n) sm.fx[n][lane][txe] = F[n];
          }
        } else if (lane_ok) {
#pragma unroll
          for (int n = 0; n < NH; ++n) sm.fy[n][r][lane] = F[n];
        }
      }
      __syncthreads();
      if constexpr (NDIM == 2) {
        if (owner) finalize(kp, fzlo);
      }
    }
  }
  if (a.dtmin) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_min = fmin(my_min, __shfl_xor_sync(FULL, my_min, o));
    if (lane == 0 && my_min < DBL_MAX)
      atomicMin(a.dtmin, (unsigned long long)__double_as_longlong(my_min));
  }
}

template <int FLUID, int RECON, int RS>
int run_rows(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
             double *const *prim_out, const apk_stage_args_t &sa, double coeff, cudaStream_t s) {
  constexpr int NH = (FLUID == APK_FLUID_GLMMHD) ? 9 : 5;
  constexpr bool PPMS = (RECON == APK_RECON_PPM);
  RowsArgs a;
  a.g = g;
  a.blocks = blocks;
  a.prim_out = prim_out;
  a.g0 = sa.gam0;
  a.g1 = sa.gam1;
  a.bdt = sa.beta_dt;
  a.dedner = coeff;
  // IEEE quotients evaluated once on the host (== frcp / fdiv on the device)
  a.rsp = {d.gamma, 1.0 / (d.gamma - 1.0), sa.c_h, 0.5 / sa.c_h};
  a.eos = {d.gamma, d.dfloor, d.pfloor, d.efloor, d.vceil, d.eceil};
  a.ntx = (d.nx[0] + RTX - 1) / RTX;
  a.flags = sa.flags;
  a.dtmin = reinterpret_cast<unsigned long long *>(sa.d_dtmin);
  a.push = sa.d_push;
  a.status = sa.d_status;
  a.fofc = sa.d_fofc_count;
  const int nty = (d.nx[1] + RTY - 1) / RTY;
  // x3 chunks: aim for >= 16 waves of CTAs (tail < 6 %) while chunks stay >= 8 planes
  int nchunk = 1;
  if (g.ndim == 3) {
    static const char *env = std::getenv("APK_B200_KCHUNKS");
    if (env && std::atoi(env) > 0) {
      nchunk = std::min(std::atoi(env), d.nx[2]);
    } else {
      const long slots = (long)apk_common::sm_count() * APK_ROWS_MINBLOCKS;
      while (nchunk < 8 && (long)a.ntx * nty * g.nb * nchunk < 16 * slots &&
             d.nx[2] / (nchunk * 2) >= 4)
        nchunk *= 2;
    }
  }
  a.nchunk = nchunk;
  dim3 grid(a.ntx * nty, g.nb, nchunk);
  const size_t smem = sizeof(RowsSmem<NH, PPMS>);
  const int dev = apk_common::current_device();
#define APK_LAUNCH_ROWS(ND)                                                                    \
  {                                                                                            \
    auto kern = k_stage_rows<FLUID, RECON, RS, ND>;                                            \
    static unsigned long long attr_set = 0;                                                    \
    if (dev >= 64 || !(attr_set >> dev & 1ULL)) {                                              \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
      if (dev < 64) attr_set |= 1ULL << dev;                                                   \
    }                                                                                          \
    kern<<<grid, RNT, smem, s>>>(a);                                                           \
  }
  if (g.ndim == 3) APK_LAUNCH_ROWS(3)
  else APK_LAUNCH_ROWS(2)
#undef APK_LAUNCH_ROWS
  apk_common::count_launch();
  return 0;
}

template <int FLUID, int RS>
int rows_recon(int recon, const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
               double *const *prim_out, const apk_stage_args_t &a, double coeff, cudaStream_t s) {
  switch (recon) {
  case APK_RECON_DC: return run_rows<FLUID, APK_RECON_DC, RS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_PLM: return run_rows<FLUID, APK_RECON_PLM, RS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_PPM: return run_rows<FLUID, APK_RECON_PPM, RS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_WENOZ: return run_rows<FLUID, APK_RECON_WENOZ, RS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_WENO3: return run_rows<FLUID, APK_RECON_WENO3, RS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_LIMO3: return run_rows<FLUID, APK_RECON_LIMO3, RS>(d, g, blocks, prim_out, a, coeff, s);
  }
  apk_common::set_error("apk_stage_fused: unknown reconstruction");
  return 1;
}

}  // namespace

#ifndef APK_FUSED_PART
#define APK_FUSED_PART 0
#endif
#if APK_FUSED_PART == 0
int launch_stage_rows_euler(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                            double *const *prim_out, const apk_stage_args_t &a, double coeff,
                            int recon, cudaStream_t s) {
  switch (d.riemann) {
  case APK_RIEMANN_NONE: return run_rows<APK_FLUID_EULER, APK_RECON_DC, APK_RIEMANN_NONE>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RIEMANN_HLLE: return rows_recon<APK_FLUID_EULER, APK_RIEMANN_HLLE>(recon, d, g, blocks, prim_out, a, coeff, s);
  case APK_RIEMANN_HLLC: return rows_recon<APK_FLUID_EULER, APK_RIEMANN_HLLC>(recon, d, g, blocks, prim_out, a, coeff, s);
  }
  apk_common::set_error("apk_stage_fused: unsupported fluid/riemann combination");
  return 1;
}
#else
int launch_stage_rows_mhd(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                          double *const *prim_out, const apk_stage_args_t &a, double coeff,
                          int recon, cudaStream_t s) {
  switch (d.riemann) {
  case APK_RIEMANN_NONE: return run_rows<APK_FLUID_GLMMHD, APK_RECON_DC, APK_RIEMANN_NONE>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RIEMANN_HLLE: return rows_recon<APK_FLUID_GLMMHD, APK_RIEMANN_HLLE>(recon, d, g, blocks, prim_out, a, coeff, s);
  case APK_RIEMANN_HLLD: return rows_recon<APK_FLUID_GLMMHD, APK_RIEMANN_HLLD>(recon, d, g, blocks, prim_out, a, coeff, s);
  }
  apk_common::set_error("apk_stage_fused: unsupported fluid/riemann combination");
  return 1;
}
#endif

}  // namespace APK_NS
