// apk_rows.cu -- the row-pass form of the fused stage kernel (2-D and 3-D meshes; replaces K1-K3,
// K6, K7, K10, K11 and the ghost exchange of SURVEY.md 2.3 with one launch per stage).
//
// Same contract and the same tile decomposition as k_stage_fused (apk_fused.cu: a CTA of 4 warps
// owns a 32 x 4 column of one meshblock and marches along x3), but the work inside a plane follows
// OWNERSHIP instead of a task table: warp w owns tile row w, lane i owns cell i of that row for the
// whole march.  Per plane a warp runs four steps through ONE copy of the reconstruction code and
// ONE copy of the Riemann solver (runtime sweep direction; the loop body stays ~40 KB of SASS):
//   step 0 (x3): reconstruct the own row along x3 -> Riemann problem of face k (left state: the
//                previous plane's, kept in a thread-private shared-memory slot) -> the cells of
//                plane k-1 now have all six fluxes: FINISH them (flux divergence in the reference's
//                association, low-storage update, Dedner damping, cons->prim, dt reduction, ghost
//                push) -- all of it warp-private, no barrier;
//   step 1 (x1): reconstruct the own row along x1 (+ the two cells beside the row, one lane per
//                (variable, side)) -> faces 0..31 of the row -> fx;
//   step 2 (x2): reconstruct the own row along x2 and a share of rows -1 / TY -> __syncthreads ->
//                face row w -> fy;
//   step 3     : warp 2: face row TY, warp 3: the x1 face behind the last lane of every row;
//                warps 0/1 prefetch the next plane -> __syncthreads.
// Reconstructed states travel through warp-private staging slots (x3, x1) or the x2 arrays; no
// (direction, variable, row) decode, two barriers per plane instead of three.
// PPM (ppm_simple.hpp:38-162): every lane evaluates the limiter-free part (recon_ppm_fastpath);
// cells that need the interface or the extremum limiter are appended to a per-warp work list and
// re-evaluated with the full recon_ppm() 32 at a time before the step's Riemann solves.
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

template <int NH>
struct RowsSmem {
  double stL[RNW][NH][RTX];      // warp-private staging: L states of the own row's faces (x3, x1)
  double stR[RNW][NH][RTX];      // ... R states
  double wly[NH][RTY + 1][RTX];  // x2: L states of face rows 0..TY (slot order of the x2 sweep)
  double wry[NH][RTY + 1][RTX];  // x2: R states
  double fx[NH][RTY][RTX + 1];   // x1: fluxes of faces 0..TX
  double fy[NH][RTY + 1][RTX];   // x2: fluxes
  double xel[NH][RTY];           // x1: L / R state of the face behind the last lane, per row
  double xer[NH][RTY];
  double lz[NH][RTY][RTX];       // x3: L state of the next face (thread-private slot)
  double *push_prim[27], *push_cons[27];
  // per-warp work list of cells that need a PPM limiter: kind << 9 | index << 5 | lane; a step
  // appends at most NH own-row + 5 extra tasks of 32 cells
  static constexpr int WL = (NH + 5) * 32;
  unsigned short wl[RNW][WL];
};
static_assert(3 * (sizeof(RowsSmem<9>) + 1024) <= 233472 || APK_ROWS_MINBLOCKS < 3,
              "three CTAs per SM must fit");

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

// U independent reconstructions per lane: q[u] points at the lane's cell, st = element stride along
// the sweep, dl / dr = shared-memory index (in doubles) of the L / R destination or -1, ent[u] =
// work-list code of the task (lane bits clear).  Must be called by all 32 lanes of the warp.
template <int RECON, int U>
APK_DEV void recon_group(double *smd, const double *const (&q)[U], int st, const int (&dl)[U],
                         const int (&dr)[U], const int (&ent)[U], const bool (&pos)[U], double dxd,
                         unsigned short *wl, int &nlist, int lane) {
  double v[U][5], ql[U], qr[U];
#pragma unroll
  for (int u = 0; u < U; ++u) load_stencil<RECON>(q[u], st, v[u]);
  if constexpr (RECON == APK_RECON_PPM) {
    bool fl[U];
    recon_ppm_fastpath<U>(v, ql, qr, fl);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool f = fl[u] && (dl[u] >= 0 || dr[u] >= 0);
      if (dl[u] >= 0 && !f) smd[dl[u]] = ql[u];
      if (dr[u] >= 0 && !f) smd[dr[u]] = qr[u];
      const unsigned bal = __ballot_sync(FULL, f);
      if (bal) {
        if (f) wl[nlist + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)(ent[u] | lane);
        nlist += __popc(bal);
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      recon_stencil<RECON>(v[u], dxd, pos[u], ql[u], qr[u]);
      if (dl[u] >= 0) smd[dl[u]] = ql[u];
      if (dr[u] >= 0) smd[dr[u]] = qr[u];
    }
  }
}

template <int FLUID, int RECON, int RS, int NDIM>
__global__ void __launch_bounds__(RNT, APK_ROWS_MINBLOCKS)
k_stage_rows(const __grid_constant__ RowsArgs a) {
  constexpr int NH = (FLUID == APK_FLUID_GLMMHD) ? 9 : 5;
  constexpr int R = ReconTraits<RECON>::radius;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SmemT = RowsSmem<NH>;
  SmemT &sm = *reinterpret_cast<SmemT *>(smem_raw);
  double *const smd = reinterpret_cast<double *>(smem_raw);
  const ApkGeom &g = a.g;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const apk_block_t &B = a.blocks[blockIdx.y];
  const int tix = blockIdx.x % a.ntx, tiy = blockIdx.x / a.ntx;
  const int i0 = g.is + tix * RTX, j0 = g.js + tiy * RTY;
  const int txe = min(RTX, g.ie + 1 - i0);  // valid cells of the tile in x1
  const int tye = min(RTY, g.je + 1 - j0);  // ... in x2
  const int sj = g.n1;
  const long sk = (long)g.n1 * g.n2;
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
  const int rjc = min(rj, tye - 1), lc = min(lane, txe - 1);

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

  // shared-memory indices (in doubles) of the arrays
  const int iStL = (int)(&sm.stL[warp][0][0] - smd), iStR = (int)(&sm.stR[warp][0][0] - smd);
  const int iWly = (int)(&sm.wly[0][0][0] - smd), iWry = (int)(&sm.wry[0][0][0] - smd);
  const int iFx = (int)(&sm.fx[0][0][0] - smd), iFy = (int)(&sm.fy[0][0][0] - smd);
  const int iXel = (int)(&sm.xel[0][0] - smd), iXer = (int)(&sm.xer[0][0] - smd);
  const int iLz = (int)(&sm.lz[0][0][0] - smd);
  constexpr int VS_ST = RTX, VS_Y = (RTY + 1) * RTX, VS_FX = RTY * (RTX + 1), VS_Z = RTY * RTX;
  unsigned short *const wl = sm.wl[warp];
  int nlist = 0;  // entries in the warp's work list (warp uniform)

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
    const double *const pplane = prim + (long)kp * sk;
    // own-row task of step `step` (sweep direction dir), slot n, for lane ln (< txe)
    auto row_task = [&](int dir, int n, int ln, const double *&q, int &dl, int &dr) {
      q = pplane + (long)var_of_slot(n, dir) * ncell + (long)(j0 + rjc) * sj + (i0 + ln);
      if (dir == 2) {
        dl = iStL + n * VS_ST + ln;
        dr = iStR + n * VS_ST + ln;
      } else if (dir == 0) {
        // ql of cell c is the L state of face c+1 (the face behind the last cell has no lane)
        dl = (ln == txe - 1) ? iXel + n * RTY + rj : iStL + n * VS_ST + ln + 1;
        dr = iStR + n * VS_ST + ln;
      } else {
        dl = iWly + n * VS_Y + (rj + 1) * RTX + ln;
        dr = iWry + n * VS_Y + rj * RTX + ln;
      }
    };
    // extra task m of the step for lane ln.  x1 (dir 0), m == 0: the two cells beside the own row,
    // lane = 2 * slot + side; x2 (dir 1): rows -1 and tye of slot (warp + 4 m) >> 1
    auto misc_task = [&](int dir, int m, int ln, const double *&q, int &st, int &dl, int &dr) {
      dl = dr = -1;
      if (dir == 0) {
        const int n = min(ln >> 1, NH - 1), side = ln & 1;
        q = pplane + (long)n * ncell + (long)(j0 + rjc) * sj + (i0 + (side ? txe : -1));
        st = 1;
        if (ln < 2 * NH) {
          if (side == 0) dl = iStL + n * VS_ST;          // L state of face 0
          else dr = iXer + n * RTY + rj;                  // R state of face txe
        }
      } else {
        const int e = warp + 4 * m, n = min(e >> 1, NH - 1), which = e & 1;
        q = pplane + (long)var_of_slot(n, 1) * ncell + (long)(j0 + (which ? tye : -1)) * sj +
            (i0 + min(ln, txe - 1));
        st = sj;
        if (e < 2 * NH && ln < txe) {
          if (which == 0) dl = iWly + n * VS_Y + ln;      // L state of face row 0
          else dr = iWry + n * VS_Y + tye * RTX + ln;     // R state of face row tye
        }
      }
    };
    const int s0 = (NDIM == 3) ? 0 : 1, s1 = inplane ? 4 : 1;
#pragma unroll 1
    for (int step = s0; step < s1; ++step) {
      const int dir = (step == 0) ? 2 : (step == 1 ? 0 : 1);
      // ------------------------------------------------------------------ reconstruction
      if (step < 3) {
        const int st = (dir == 0) ? 1 : (dir == 1 ? sj : (int)sk);
        const double dxd = (dir == 0) ? dx1 : (dir == 1 ? dx2 : dx3);
        if (row_ok) {
          const double *q0;
          int dl0, dr0, dl1, dr1;
          row_task(dir, 0, lc, q0, dl0, dr0);
          row_task(dir, 1, lc, q0, dl1, dr1);   // q0 is reset below; dl1 - dl0 = variable stride
          const int dls = dl1 - dl0, drs = dr1 - dr0;
          q0 = pplane + (long)(j0 + rjc) * sj + (i0 + lc);
          const long o1 = (long)(1 + dir) * ncell, o2 = (long)(1 + (dir + 1) % 3) * ncell,
                     o3 = (long)(1 + (dir + 2) % 3) * ncell;
          auto D = [&](int base, int vs, int n) { return lane_ok ? base + n * vs : -1; };
          if constexpr (NH == 9) {
            {
              const double *const q[3] = {q0, q0 + 4 * ncell, q0 + 8 * ncell};
              const int dl[3] = {D(dl0, dls, 0), D(dl0, dls, 4), D(dl0, dls, 8)};
              const int dr[3] = {D(dr0, drs, 0), D(dr0, drs, 4), D(dr0, drs, 8)};
              const int ent[3] = {0 << 5, 4 << 5, 8 << 5};
              const bool pos[3] = {true, true, false};
              recon_group<RECON, 3>(smd, q, st, dl, dr, ent, pos, dxd, wl, nlist, lane);
            }
            {
              const double *const q[3] = {q0 + o1, q0 + o2, q0 + o3};
              const int dl[3] = {D(dl0, dls, 1), D(dl0, dls, 2), D(dl0, dls, 3)};
              const int dr[3] = {D(dr0, drs, 1), D(dr0, drs, 2), D(dr0, drs, 3)};
              const int ent[3] = {1 << 5, 2 << 5, 3 << 5};
              const bool pos[3] = {false, false, false};
              recon_group<RECON, 3>(smd, q, st, dl, dr, ent, pos, dxd, wl, nlist, lane);
            }
            {
              const double *const q[3] = {q0 + o1 + 4 * ncell, q0 + o2 + 4 * ncell, q0 + o3 + 4 * ncell};
              const int dl[3] = {D(dl0, dls, 5), D(dl0, dls, 6), D(dl0, dls, 7)};
              const int dr[3] = {D(dr0, drs, 5), D(dr0, drs, 6), D(dr0, drs, 7)};
              const int ent[3] = {5 << 5, 6 << 5, 7 << 5};
              const bool pos[3] = {false, false, false};
              recon_group<RECON, 3>(smd, q, st, dl, dr, ent, pos, dxd, wl, nlist, lane);
            }
          } else {
            {
              const double *const q[2] = {q0, q0 + 4 * ncell};
              const int dl[2] = {D(dl0, dls, 0), D(dl0, dls, 4)};
              const int dr[2] = {D(dr0, drs, 0), D(dr0, drs, 4)};
              const int ent[2] = {0 << 5, 4 << 5};
              const bool pos[2] = {true, true};
              recon_group<RECON, 2>(smd, q, st, dl, dr, ent, pos, dxd, wl, nlist, lane);
            }
            {
              const double *const q[3] = {q0 + o1, q0 + o2, q0 + o3};
              const int dl[3] = {D(dl0, dls, 1), D(dl0, dls, 2), D(dl0, dls, 3)};
              const int dr[3] = {D(dr0, drs, 1), D(dr0, drs, 2), D(dr0, drs, 3)};
              const int ent[3] = {1 << 5, 2 << 5, 3 << 5};
              const bool pos[3] = {false, false, false};
              recon_group<RECON, 3>(smd, q, st, dl, dr, ent, pos, dxd, wl, nlist, lane);
            }
          }
        }
        // extra tasks, two at a time
        const int nmisc = (step == 1) ? (row_ok ? 1 : 0) : (step == 2 ? (2 * NH - warp + 3) / 4 : 0);
#pragma unroll 1
        for (int m0 = 0; m0 < nmisc; m0 += 2) {
          const double *qa, *qb;
          int sta, stb, dla, dra, dlb, drb;
          misc_task(dir, m0, lane, qa, sta, dla, dra);
          misc_task(dir, min(m0 + 1, nmisc - 1), lane, qb, stb, dlb, drb);
          if (m0 + 1 >= nmisc) dlb = drb = -1;
          const double *const q[2] = {qa, qb};
          const int dl[2] = {dla, dlb}, dr[2] = {dra, drb};
          const int ent[2] = {(1 << 9) | (m0 << 5), (1 << 9) | ((m0 + 1) << 5)};
          // positivity (LimO3): slots 0 and 4 (density, pressure)
          const int na = (dir == 0) ? (lane >> 1) : ((warp + 4 * m0) >> 1);
          const int nb = (dir == 0) ? (lane >> 1) : ((warp + 4 * (m0 + 1)) >> 1);
          const bool pos[2] = {na == 0 || na == 4, nb == 0 || nb == 4};
          recon_group<RECON, 2>(smd, q, sta, dl, dr, ent, pos, dxd, wl, nlist, lane);
        }
        if constexpr (RECON == APK_RECON_PPM) {
          // the cells that need a limiter: full PPM, one cell per lane
          __syncwarp();
#pragma unroll 1
          for (int b = 0; b < nlist; b += 32) {
            if (b + lane < nlist) {
              const unsigned e = wl[b + lane];
              const int ln = e & 31, idx = (e >> 5) & 15;
              const double *q;
              int es = st, dl, dr;
              if (e >> 9) misc_task(dir, idx, ln, q, es, dl, dr);
              else row_task(dir, idx, ln, q, dl, dr);
              double s5[5], ql, qr;
              load_stencil<RECON>(q, es, s5);
              recon_ppm(s5[0], s5[1], s5[2], s5[3], s5[4], ql, qr);
              if (dl >= 0) smd[dl] = ql;
              if (dr >= 0) smd[dr] = qr;
            }
          }
          nlist = 0;
        }
        __syncwarp();
      }
      if (step == 2) __syncthreads();  // the x2 states (and the x1 edge states) of all rows are in place
      // ------------------------------------------------------------------ Riemann problems
      bool act, fst = lane_ok;
      int Li, Ls, Ri, Rs, Fi = 0, Fs = 0;
      if (step == 0) {
        act = row_ok;
        Li = iLz + rj * RTX + lane; Ls = VS_Z;
        Ri = iStR + lc; Rs = VS_ST;
      } else if (step == 1) {
        act = row_ok;
        Li = iStL + lc; Ls = VS_ST;
        Ri = iStR + lc; Rs = VS_ST;
        Fi = iFx + rj * (RTX + 1) + lane; Fs = VS_FX;
      } else {
        const int r = (step == 2) ? warp : RTY;
        act = (step == 2) ? (warp <= tye) : (warp == 2 && tye == RTY);
        Li = iWly + r * RTX + lc; Ls = VS_Y;
        Ri = iWry + r * RTX + lc; Rs = VS_Y;
        Fi = iFy + r * RTX + lane; Fs = VS_Y;
        if (step == 3 && warp == 3) {  // the x1 faces behind the last lane, lane = tile row
          const int le = min(lane, tye - 1);
          act = true;
          Li = iXel + le; Ls = RTY;
          Ri = iXer + le; Rs = RTY;
          Fi = iFx + lane * (RTX + 1) + txe; Fs = VS_FX;
          fst = lane < tye;
        }
      }
      if (act) {
        double L[NH], Rr[NH], F[NH];
#pragma unroll
        for (int n = 0; n < NH; ++n) {
          L[n] = smd[Li + n * Ls];
          Rr[n] = smd[Ri + n * Rs];
        }
        if (step == 0) {
          // the L state of the next x3 face replaces the one just read (thread-private slot)
#pragma unroll
          for (int n = 0; n < NH; ++n) smd[Li + n * Ls] = smd[iStL + lc + n * VS_ST];
        }
        if (step != 0 || kp >= kbeg) {
          riemann_solve<FLUID, RS>(L, Rr, a.rsp, F);
          if (step == 0) {
            // the cells of plane kp-1 now have all six fluxes (fx / fy of plane kp-1 are still in
            // shared memory: fx is rewritten in step 1, fy after the next barrier)
            if (lane_ok && kp - 1 >= kbeg) finalize(kp - 1, F);
#pragma unroll
            for (int n = 0; n < NH; ++n) fzlo[n] = F[n];
          } else if (fst) {
#pragma unroll
            for (int n = 0; n < NH; ++n) smd[Fi + n * Fs] = F[n];
          }
        }
      } else if (step == 3 && APK_ROWS_PREFETCH && NDIM == 3 && warp < 2) {
        // idle warps pull what the next plane touches for the first time into L1: warp 0 the new
        // top plane of the x3 stencils (kp+1+R), warp 1 the x2 halo rows of plane kp+1
        if (kp + 1 <= klast) {
          const int nrow = (warp == 0) ? tye : 2 * (R + 1);
          const long pl = (warp == 0) ? (long)(kp + 1 + R) * sk : (long)(kp + 1) * sk;
          for (int p = lane; p < NH * nrow * 3; p += 32) {
            const int seg = p % 3, rv = p / 3, rr = rv % nrow, vv = rv / nrow;
            const int jj = (warp == 0) ? rr : (rr <= R ? rr - (R + 1) : tye + rr - (R + 1));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(prim + (long)vv * ncell + pl +
                                                         (long)(j0 + jj) * sj + i0 +
                                                         min(seg * 16, txe - 1)));
          }
        }
      }
      if (step == 3) __syncthreads();
    }
    if constexpr (NDIM == 2) {
      if (owner) finalize(kp, fzlo);
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
  const size_t smem = sizeof(RowsSmem<NH>);
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
