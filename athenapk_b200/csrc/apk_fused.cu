// apk_fused.cu -- the fused stage kernel: reconstruct -> Riemann -> flux divergence -> low-storage
// update -> Dedner damping -> cons->prim (+ dt reduction) in ONE pass over the pack, without
// materialising cons.flux (replaces K1-K3, K6, K7, K10 and K11 of SURVEY.md 2.3).
//
// Work decomposition (B200: 148 SMs; defaults NW = 4 warps, TY = 4, 3 CTAs per SM, 168 regs):
//  * a CTA owns a TX x TY = 32 x 4 column of one meshblock and MARCHES along x3 (one of `nchunk`
//    x3 chunks, blockIdx.z).  x3-reconstruction of plane k yields the right state of face k and
//    the left state of face k+1, so every x3 face is reconstructed and solved exactly once;
//    x1/x2 faces of a plane are solved inside the plane tile (33/32 and 5/4 faces per cell).
//    Three independent CTAs per SM interleave their phases (hides barrier and latency stalls).
//  * per plane three phases separated by __syncthreads():
//      recon   : warp tasks (variable, row) from a per-CTA task table -> 32 lanes along x1
//                (coalesced 256 B rows); L/R states go to shared memory in sweep-local slot order;
//                PPM runs warp-vote guarded (recon_ppm_warp).  DC stages instead stage halo'd
//                prim planes with cp.async and have no recon phase,
//      riemann : warp tasks (direction, face row); the flux overwrites the L state in place,
//      finalize: one thread per cell; the x1+x2 divergence and the lower x3 flux stay in
//                REGISTERS across the march, the cell is finished one plane later when its upper
//                x3 flux exists (association order of the reference's FluxDivHelper), cons->prim,
//                dt reduction, and the ghost PUSH into same-device / peer-GPU neighbour blocks.
// Compiled twice (APK_NS = apk_fast / apk_strict), see apk_launch.h.
#include <cstdint>
#include <cstdlib>
#include <algorithm>

#include "apk_launch.h"
#include "apk_math.cuh"

namespace APK_NS {
namespace {

#ifndef APK_FUSED_NWARPS
#define APK_FUSED_NWARPS 4
#endif
// The register file is split over 4 SM sub-partitions (16 K regs each): <= 12 resident warps cap at
// 168 regs/thread, 13..16 at 128.  Measured (profiles/r01_notes.md): 3 CTAs x 4 warps (32x4 tiles)
// beat 1 x 16, 1 x 12, 2 x 8, 2 x 6 and 3 x 5 warps -- independent CTAs fill each other's barrier
// and latency bubbles.
#ifndef APK_FUSED_UNROLL
#define APK_FUSED_UNROLL 2
#endif
#ifndef APK_FUSED_MINBLOCKS
#define APK_FUSED_MINBLOCKS 3  // CTAs per SM the register budget is sized for
#endif
#ifndef APK_FUSED_PREFETCH
#define APK_FUSED_PREFETCH 1
#endif
#ifndef APK_FUSED_SWPIPE
#define APK_FUSED_SWPIPE 0  // measured: no gain at 256^3 (the loads are not the limiter), costs registers
#endif
#ifndef APK_FUSED_PPM_VOTE
#define APK_FUSED_PPM_VOTE 1
#endif
#ifndef APK_FUSED_TY
#define APK_FUSED_TY 4
#endif
// APK_FUSED_PPM_LIST=1 (default): the recon phase evaluates only the limiter-free part of PPM
// (recon_ppm_fastpath) in the warp tasks; the few cells that need the interface or the extremum
// limiter go to a per-warp work list in shared memory and are re-evaluated with the full
// recon_ppm() 32 at a time.  With the warp vote (APK_FUSED_PPM_LIST=0) one extremum makes all
// 32 x UNROLL cells of a task pay for both limiters: on the Orszag-Tang state 56-60 % of the warp
// tasks did, for ~2 % of the cells.  Same expressions per cell, bit-identical results.
#ifndef APK_FUSED_PPM_LIST
#define APK_FUSED_PPM_LIST 1
#endif
// APK_FUSED_TMA=1 (A/B variant, namespace apk_fast_tma, selected with APK_B200_FUSED_V=3): the halo'd
// prim tile of the plane being swept -- (TX+6) x (TY+6) cells x all variables -- is staged in
// shared memory by ONE cp.async.bulk.tensor (TMA, 4-D tensor map over [var][k][j][i]) per plane,
// issued by one thread when the previous recon phase has released the buffer and awaited on an
// mbarrier; the x1 / x2 reconstructions then read their stencils from shared memory with
// compile-time strides (no per-lane global address arithmetic, no L2 round trips).  The x3
// stencils keep their global loads (a 5-plane ring does not fit).  The tile costs 27 KB, so only
// two CTAs fit per SM: the variant is built with 6 warps per CTA.
#ifndef APK_FUSED_TMA
#define APK_FUSED_TMA 0
#endif
constexpr int TX = 32, TY = APK_FUSED_TY, NWARPS = APK_FUSED_NWARPS, NTHREADS = NWARPS * 32;
static_assert(NTHREADS >= TX * TY, "finalize phase needs one thread per tile cell");

// One reconstruction warp task = one variable x one row of 32 cells.  The (direction, variable,
// row) -> (global offset, stride, shared-memory destinations) decode is plane independent, so it
// is done ONCE per CTA into a shared table; the per-plane loop then costs one LDS.128 and a few
// integer instructions per task instead of ~90 (divisions by TY / TY+2, slot permutation, 64-bit
// address arithmetic, row / tile-edge clamps) -- measured 45 % of the recon loop before.
struct ReconTask {
  int goff;    // element offset of the (clamped) row start relative to prim + plane offset
  int stride;  // element stride along the sweep | dir << 28 | positivity << 30
  int sl, sr;  // shared-memory element index of the L / R destination of lane 0, or -1
};

// NH = number of variables the arrays hold: the hydro / GLM-MHD variables plus the passive scalars
template <int NH, int NDIM, bool DC>
struct Smem {
  static constexpr int NZ = (NDIM == 3) ? NH * TY : 0, NX = NH * TY,
                       NY = (NDIM >= 2) ? NH * (TY + 2) : 0, NT = NZ + NX + NY;
  // donor cell (DC) stages need no L/R state arrays (the states are the cell values): they keep the
  // three flux arrays and instead stage three halo'd prim planes (k-1, k and the one in flight)
  static constexpr int NR = DC ? 1 : NH, ZB = DC ? 1 : 2, NP = DC ? 3 : 1, NPV = DC ? NH : 1;
  static constexpr int RY = DC ? 1 : TY, RX = DC ? 1 : TX;  // R-state arrays collapse for DC
  double wlx[NH][TY][TX + 1], wrx[NR][RY][RX + 1];
  double wly[NH][TY + 1][TX], wry[NR][RY + 1][RX];
  double wrz[NR][RY][RX];
  // x3 L states / fluxes, double buffered.  Each buffer ends in a trash row: the destination of
  // the stores of warp tasks without an L or R consumer (lean PPM path) -- the x3 double-buffer
  // offset then maps the trash row of buffer 0 onto the trash row of buffer 1.
  struct ZBuf {
    double a[NH][TY][TX];
    double trash[TX];
  } wlz[ZB];
  double pl[NP][NPV][TY + 2][TX + 2];  // DC: prim planes, cells -1..TX x rows -1..TY
  // TMA variant: halo'd prim tile of the current plane, cells -3..TX+2 x rows -3..TY+2
  static constexpr bool TMA = APK_FUSED_TMA && !DC && NDIM == 3;
  alignas(128) double tile[TMA ? NH : 1][TMA ? TY + 6 : 1][TMA ? TX + 6 : 2];
  alignas(8) unsigned long long tma_bar;
  alignas(16) ReconTask tab[(!DC && NT > 0) ? NT + 1 : 1];  // + one dummy task (lean PPM path)
  // per-warp work list of cells that need a PPM limiter: (task << 5 | lane); a batch appends at most
  // 32 * UNROLL entries to at most 31 left over
  static constexpr int WL = 32 * (APK_FUSED_UNROLL + 1);
  unsigned short wl[DC ? 1 : NWARPS][DC ? 2 : WL];
  double *push_prim[27], *push_cons[27];  // ghost-push destinations per neighbour offset (or null)
};

// NS passive scalars ride along (hydro.cpp:1089-1097): reconstructed like every other primitive,
// flux = mass flux x upwind state, s = (rho s) / rho in cons -> prim.  NS = 0 is the headline kernel;
// NS > 0 instantiations live in their own translation units (APK_FUSED_NS).
template <int FLUID, int RECON, int RS, int NDIM, int NS = 0>
__global__ void __launch_bounds__(NTHREADS, APK_FUSED_MINBLOCKS)
k_stage_fused(ApkGeom g, const apk_block_t *__restrict__ blocks, double *const *__restrict__ prim_out,
              double g0, double g1, double bdt, RsPars rsp, double dedner_coeff,
              EosPars eos, int ntx, int flags, unsigned long long *dtmin,
              const apk_push_dst_t *__restrict__ push, int nchunk, int *status,
              unsigned long long *fofc_count, const unsigned char *__restrict__ tmaps) {
  constexpr int NHY = (FLUID == APK_FLUID_GLMMHD) ? 9 : 5;  // hydro / GLM-MHD variables
  constexpr int NH = NHY + NS;                               // all variables of the pack
  // sweep-local slot <-> memory variable (scalars are not permuted; for Euler, slots 5..7 would
  // otherwise be taken for magnetic field components)
  auto var_of_slot = [](int n, int dir) { return n >= NHY ? n : APK_NS::var_of_slot(n, dir); };
  auto slot_of_var = [](int v, int dir) { return v >= NHY ? v : APK_NS::slot_of_var(v, dir); };
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool DC = (RECON == APK_RECON_DC);
  using SmemT = Smem<NH, NDIM, DC>;
  constexpr bool TMA = SmemT::TMA;
  constexpr int TW = TX + 6, TILE_V = (TY + 6) * TW;  // tile row / variable strides (elements)
  SmemT &sm = *reinterpret_cast<SmemT *>(smem_raw);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const apk_block_t &B = blocks[blockIdx.y];
  const int tix = blockIdx.x % ntx, tiy = blockIdx.x / ntx;
  const int i0 = g.is + tix * TX, j0 = g.js + tiy * TY;
  const int txe = min(TX, g.ie + 1 - i0);                      // valid cells of the tile in x1
  const int tye = (NDIM >= 2) ? min(TY, g.je + 1 - j0) : 1;    // ... in x2
  const long sj = g.n1, sk = (long)g.n1 * g.n2;
  const double *__restrict__ prim = B.prim;
  double *__restrict__ cons = B.cons;
  double *__restrict__ cons_u1 = B.cons_u1;
  const bool save_u1 = (flags & APK_FUSED_SAVE_U1) != 0;
  const bool prim_only = (flags & APK_FUSED_PRIM_ONLY) != 0;      // no cons / u1 stores
  const bool u1_is_cons = (flags & APK_FUSED_U1_FROM_CONS) != 0;  // u1 lives in cons
  double *__restrict__ po = prim_out[blockIdx.y];
  const double dx1 = B.dx[0], dx2 = B.dx[1], dx3 = B.dx[2];
  // uniform_coordinates.hpp:37-40
  const double a1 = dx2 * dx3, a2 = dx1 * dx3, a3 = dx1 * dx2, vol = dx1 * dx2 * dx3;
#ifndef APK_STRICT
  const double ivol = frcp(vol);
#endif

  // finalize-phase ownership: thread -> cell (cj, ci) of the tile
  const int ci = threadIdx.x & 31, cj = threadIdx.x >> 5;
  const bool owner = (threadIdx.x < TX * TY) && ci < txe && cj < tye;
  double du[NH], fzlo[NH];
#pragma unroll
  for (int v = 0; v < NH; ++v) du[v] = fzlo[v] = 0.0;
  double my_min = DBL_MAX;

  // x3 range of this CTA: the march is cut into `nchunk` balanced chunks (blockIdx.z) so that small
  // packs (strong scaling: 64 meshblocks per GPU) still fill every SM for several waves; a chunk
  // re-does one x3-only reconstruction plane and one x3 face row of its lower neighbour (~0.3
  // plane of work per chunk)
  const int nx3c = g.ke - g.ks + 1;  // chunk c covers planes [c nx3 / nchunk, (c+1) nx3 / nchunk)
  const int kbeg = (NDIM == 3) ? g.ks + ((int)blockIdx.z * nx3c) / nchunk : g.ks;
  const int kend = (NDIM == 3) ? g.ks + (((int)blockIdx.z + 1) * nx3c) / nchunk - 1 : g.ke;
  const int kfirst = (NDIM == 3) ? kbeg - 1 : kbeg;
  const int klast = (NDIM == 3) ? kend + 1 : kend;
  // recon work of one plane as a flat list of warp tasks (one task = one variable x one row of
  // 32 cells): [0,NZ) x3 rows, [NZ,NZ+NX) x1 rows, [NZ+NX, NZ+NX+NY) x2 rows -1..TY.
  // UNROLL tasks are fetched before the first one is evaluated (memory-level parallelism; with
  // 4 warps per scheduler there are too few warps to hide L2 latency otherwise).
  constexpr int NZ = SmemT::NZ, NX = SmemT::NX, NT = SmemT::NT;
  constexpr int UNROLL = (RECON == APK_RECON_DC) ? 4 : APK_FUSED_UNROLL;
  constexpr int ZBUF = NH * TY * TX + TX;  // elements between wlz[0] and wlz[1]
  // lean recon phase (PPM with work lists, global-memory stencils): see the recon phase below
  constexpr bool LEAN = (RECON == APK_RECON_PPM) && APK_FUSED_PPM_LIST && !TMA;
  [[maybe_unused]] const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
  double *const smd = reinterpret_cast<double *>(smem_raw);
  if constexpr (RECON != APK_RECON_DC) {
    for (int t = threadIdx.x; t < NT; t += NTHREADS) {
      int dir, n, r;
      if (t < NZ) { dir = 2; n = t / TY; r = t % TY; }
      else if (t < NZ + NX) { dir = 0; n = (t - NZ) / TY; r = (t - NZ) % TY; }
      else { dir = 1; n = (t - NZ - NX) / (TY + 2); r = (t - NZ - NX) % (TY + 2) - 1; }
      // out-of-range rows evaluate a clamped (valid) stencil and skip the store, so the UNROLL
      // reconstructions stay straight-line code the scheduler can interleave
      const bool rok = (dir == 1) ? r <= tye : r < tye;
      const int rc = min(r, dir == 1 ? tye : tye - 1);
      ReconTask e;
      e.goff = var_of_slot(n, dir) * (int)g.ncell + rc * (int)sj;
      e.stride = (dir == 0 ? 1 : (dir == 1 ? (int)sj : (int)sk)) | (dir << 28) |
                 ((n == 0 || n == 4) ? (1 << 30) : 0);
      if (TMA && dir != 2) {  // x1 / x2 stencils come from the staged tile
        e.goff = var_of_slot(n, dir) * TILE_V + (rc + 3) * TW + 3;
        e.stride = (dir == 0 ? 1 : TW) | (dir << 28) | ((n == 0 || n == 4) ? (1 << 30) : 0);
      }
      e.sl = e.sr = -1;
      if (rok) {
        if (dir == 2) {
          e.sl = (int)(&sm.wlz[0].a[n][r][0] - smd);
          e.sr = (int)(&sm.wrz[n][r][0] - smd);
        } else if (dir == 0) {
          e.sl = (int)(&sm.wlx[n][r][1] - smd);
          e.sr = (int)(&sm.wrx[n][r][0] - smd);
        } else {
          if (r + 1 <= tye) e.sl = (int)(&sm.wly[n][r + 1][0] - smd);
          if (r >= 0) e.sr = (int)(&sm.wry[n][r][0] - smd);
        }
      }
      if constexpr (LEAN) {
        // lean PPM path: pure stride, destinations as BYTE offsets in shared memory (rows without
        // a consumer store into the trash row, so the loop body needs no destination tests)
        const int tr8 = (int)(reinterpret_cast<unsigned char *>(&sm.wlz[0].trash[0]) - smem_raw);
        e.stride &= 0x0fffffff;
        e.sl = (e.sl >= 0) ? e.sl * 8 : tr8;
        e.sr = (e.sr >= 0) ? e.sr * 8 : tr8;
      }
      sm.tab[t] = e;
    }
    if constexpr (LEAN) {
      if (threadIdx.x == 0) {
        const int tr8 = (int)(reinterpret_cast<unsigned char *>(&sm.wlz[0].trash[0]) - smem_raw);
        sm.tab[NT] = ReconTask{0, 1, tr8, tr8};  // what out-of-range task slots of a batch evaluate
      }
    }
    __syncthreads();
  }
  if (flags & (APK_FUSED_PUSH_PRIM | APK_FUSED_PUSH_CONS)) {
    if (threadIdx.x < 27) {
      double *pp = nullptr, *pc = nullptr;
      if (push) {  // explicit destinations (may be peer-GPU memory: halo exchange as NVLink stores)
        pp = push[blockIdx.y * 27 + threadIdx.x].prim;
        pc = push[blockIdx.y * 27 + threadIdx.x].cons;
      } else {
        const int nb = B.nbr[threadIdx.x];
        if (nb >= 0) {
          pp = prim_out[nb];
          pc = blocks[nb].cons;
        }
      }
      sm.push_prim[threadIdx.x] = (flags & APK_FUSED_PUSH_PRIM) ? pp : nullptr;
      sm.push_cons[threadIdx.x] = (flags & APK_FUSED_PUSH_CONS) ? pc : nullptr;
    }
    __syncthreads();
  }
  const int lc = min(lane, txe - 1);
  const bool lane_ok = lane < txe;
  // ---- TMA staging of the halo'd prim tile (one bulk tensor copy per plane) ----
  [[maybe_unused]] unsigned tma_phase = 0;
  [[maybe_unused]] const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&sm.tma_bar);
  [[maybe_unused]] auto tma_issue = [&](int k) {
    if constexpr (TMA) {
      if (threadIdx.x == 0) {
        constexpr unsigned BYTES = NH * TILE_V * sizeof(double);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&sm.tile[0][0][0]);
        const unsigned long long map = (unsigned long long)(tmaps + (size_t)blockIdx.y * 128);
        // the generic-proxy reads of the previous plane are ordered before the async-proxy write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr),
                     "r"(BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
            "[%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"(map), "r"(i0 - 3), "r"(j0 - 3), "r"(k),
            "r"(0), "r"(bar_addr) : "memory");
      }
    }
  };
  [[maybe_unused]] auto tma_wait = [&]() {
    if constexpr (TMA) {
      unsigned done = 0;
      while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar_addr), "r"(tma_phase) : "memory");
      }
      tma_phase ^= 1;
    }
  };
  if constexpr (TMA) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    tma_issue(kbeg);  // the first in-plane iteration's tile
  }
  // DC: asynchronous copy (cp.async, 8 B granules: always aligned) of one halo'd prim plane into
  // the staging ring; indices beyond the tile are clamped to valid cells and never used.  The
  // plane needed by iteration k+1 is requested at the start of iteration k and has the whole
  // iteration to land: the Riemann tasks then read L/R from shared memory instead of issuing 18
  // long-latency global loads per face (3.3 faces per cell, every value fetched ~6 times).
  // (element slots of one variable's halo'd plane are decoded once per thread: in-plane source
  //  offset and destination index; per plane and variable a slot then costs one address + cp.async)
  constexpr int SP_ROWS = (NDIM >= 2) ? TY + 2 : 1, SP_PER = SP_ROWS * (TX + 2);
  constexpr int SP_N = DC ? (SP_PER + NTHREADS - 1) / NTHREADS : 1;
  [[maybe_unused]] int sp_src[SP_N], sp_dst[SP_N];
  if constexpr (DC) {
#pragma unroll
    for (int q = 0; q < SP_N; ++q) {
      const int e = threadIdx.x + q * NTHREADS;
      const int r = e / (TX + 2), c = e % (TX + 2);
      const int jj = (NDIM >= 2) ? min(r - 1, tye) : 0, ii = min(c - 1, txe);
      sp_src[q] = (j0 + jj) * (int)sj + (i0 + ii);
      sp_dst[q] = (e < SP_PER) ? ((NDIM >= 2) ? r : 1) * (TX + 2) + c : -1;
    }
  }
  auto stage_plane = [&](int k) {
    if constexpr (DC) {
      const int buf = k % 3;
      const double *src = prim + (long)k * sk;
      const unsigned dst0 = (unsigned)__cvta_generic_to_shared(&sm.pl[buf][0][0][0]);
#pragma unroll
      for (int v = 0; v < NH; ++v) {
#pragma unroll
        for (int q = 0; q < SP_N; ++q) {
          if (sp_dst[q] >= 0) {
            const unsigned dst = dst0 + (unsigned)(v * (TY + 2) * (TX + 2) + sp_dst[q]) * 8u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst),
                         "l"(src + (long)v * g.ncell + sp_src[q]) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  if constexpr (DC) {
    stage_plane(kfirst);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
  // prefetch targets of a plane (see the Riemann phase): one 128 B line each, decoded once --
  // the rows of plane pn + R the x3 stencils touch for the first time and the halo rows / columns
  // of plane pn itself; element offsets relative to prim + pn * sk (negative: none)
  constexpr int PFN = 4;
  [[maybe_unused]] int pf_off[PFN];
  [[maybe_unused]] unsigned pf_zrow = 0;
  if constexpr (RECON != APK_RECON_DC && APK_FUSED_PREFETCH) {
    constexpr int R = ReconTraits<RECON>::radius;
    const int nzr = (NDIM == 3) ? tye : 0;                          // rows of plane pn+R
    const int nyr = TMA ? 0 : (NDIM >= 2 ? tye + 2 + 2 * R : 1);    // rows of plane pn
    const int ylo = (NDIM >= 2) ? -1 - R : 0;
    const int nrow = nzr + nyr;
    const int lo = -(1 + R), hi = txe + R;                          // cell range along x1 (inclusive)
    const int nseg = ((hi - lo + 1) * 8 + 127) / 128 + 1;
#pragma unroll
    for (int i = 0; i < PFN; ++i) {
      const int p = threadIdx.x + i * NTHREADS;
      pf_off[i] = -1;
      if (p < NH * nrow * nseg) {
        const int seg = p % nseg, rv = p / nseg, rr = rv % nrow, v = rv / nrow;
        const long row = (rr < nzr) ? (long)R * sk + (long)(j0 + rr) * sj
                                    : (long)(j0 + ylo + (rr - nzr)) * sj;
        pf_off[i] = (int)((long)v * g.ncell + row + i0 + min(lo + seg * 16, hi));
        if (rr < nzr) pf_zrow |= 1u << i;
      }
    }
  }
  int cur = 0;
  for (int kp = kfirst; kp <= klast; ++kp) {
    const bool inplane = (kp >= kbeg && kp <= kend);
    if constexpr (DC) {
      if (kp < klast) stage_plane(kp + 1);
    }
    // ------------------------------------------------------------------ recon phase
    // (donor cell: L/R states are the neighbouring cell values themselves; the Riemann tasks read
    //  them straight from global memory, no staging pass and no barrier)
    if constexpr (RECON != APK_RECON_DC) {
      const int t_end = inplane ? NT : NZ;
      const long plane = (long)kp * sk + (long)j0 * sj + i0;
      const double *__restrict__ pbase = prim + plane + lc;
      [[maybe_unused]] const double *tile = &sm.tile[0][0][0] + lc;
      if (TMA && inplane) tma_wait();  // the tile of plane kp has landed
      const int zoff = (cur ^ 1) * ZBUF;
      // Software pipeline over batches of UNROLL tasks: the stencil loads of batch b+1 are in
      // flight while batch b is evaluated (two register buffers A/B, loop unrolled by two so no
      // register copies).  With <= 4 warps per scheduler and ~40 KB of L1 next to the shared
      // arrays of three resident CTAs, most of these loads are L2 hits (~300-600 cycles) that
      // nothing else can hide (measured 22 % of the recon phase in long_scoreboard before).
      constexpr int STEP = UNROLL * NWARPS;
      if constexpr (LEAN) {
        // ---- lean PPM recon: limiter-free fast path in the warp tasks, flagged cells to a per-warp
        // work list (recon_ppm_fastpath, apk_math.cuh); stores through 32-bit shared addresses
        int nlist = 0;  // entries in this warp's work list (warp uniform)
        unsigned short *const wl = sm.wl[warp];
        const unsigned zoff8 = (unsigned)zoff * 8u;
        const unsigned lane8 = sbase + (unsigned)lane * 8u;
        const double *pb = pbase;
        asm volatile("" : "+l"(pb));  // keep the plane base as one 64-bit register pair
#ifdef APK_FUSED_NO_TRASH_STORES
        // sanitizer build: stores into the trash rows (unsynchronised write-after-write BY DESIGN:
        // nobody reads them) are skipped so that racecheck only sees real hazards
        const unsigned trash_lo = sbase + (unsigned)(reinterpret_cast<unsigned char *>(&sm.wlz[0].trash[0]) - smem_raw);
        const unsigned trash_sz = (unsigned)sizeof(sm.wlz[0]);
        auto sts = [&](unsigned addr, double x) {
          const unsigned o = (addr - trash_lo) % trash_sz;
          if (addr >= trash_lo && o < TX * sizeof(double)) return;
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(x) : "memory");
        };
#else
        auto sts = [](unsigned addr, double x) {
          asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(x) : "memory");
        };
#endif
        constexpr int NEC = NH * TY * 2;  // x1 end cells: (slot, row, side)
        // full PPM for one listed cell per lane (entries [base, base + n) of the list)
        auto drain = [&](int base, int n) {
          __syncwarp();
          if (lane < n) {
            const unsigned ent = wl[base + lane];
            const int tr = ent >> 5, ln = ent & 31;
            const double *q;
            int st;
            unsigned al, ar;
            if (tr <= NT) {
              const int4 e = *reinterpret_cast<const int4 *>(&sm.tab[tr]);
              q = prim + plane + ln + e.x;
              st = e.y;
              al = sbase + ln * 8 + e.z + (tr < NZ ? zoff8 : 0u);
              ar = sbase + ln * 8 + e.w;
            } else {
              const int ec = ((tr - NT - 1) << 5) | ln;
              const int side = ec & 1, j = (ec >> 1) % TY, n2 = (ec >> 1) / TY;
              q = prim + (long)var_of_slot(n2, 0) * g.ncell + plane + (long)j * sj + (side ? txe : -1);
              st = 1;
              const unsigned tr8 = sbase + (unsigned)(reinterpret_cast<unsigned char *>(&sm.wlz[0].trash[0]) - smem_raw);
              al = side ? tr8 : sbase + (unsigned)(reinterpret_cast<unsigned char *>(&sm.wlx[n2][j][0]) - smem_raw);
              ar = side ? sbase + (unsigned)(reinterpret_cast<unsigned char *>(&sm.wrx[n2][j][txe]) - smem_raw) : tr8;
            }
            double s5[5], ql, qr;
            load_stencil<RECON>(q, st, s5);
            recon_ppm(s5[0], s5[1], s5[2], s5[3], s5[4], ql, qr);
            sts(al, ql);
            sts(ar, qr);
          }
          __syncwarp();
        };
        for (int t0 = warp; t0 < t_end; t0 += STEP) {
          double v[UNROLL][5], qls[UNROLL], qrs[UNROLL];
          unsigned al[UNROLL], ar[UNROLL];
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            const int tr = t0 + u * NWARPS;
            const int4 e = *reinterpret_cast<const int4 *>(&sm.tab[tr < t_end ? tr : NT]);
            al[u] = lane8 + e.z + (tr < NZ ? zoff8 : 0u);
            ar[u] = lane8 + e.w;
            load_stencil<RECON>(pb + e.x, e.y, v[u]);
          }
          bool fl[UNROLL];
          bool any = recon_ppm_fastpath<UNROLL>(v, qls, qrs, fl);
          // (lanes beyond a ragged tile and out-of-range task slots neither store nor list)
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            if (lane_ok && !fl[u]) {
              sts(al[u], qls[u]);
              sts(ar[u], qrs[u]);
            }
          }
          if (__any_sync(0xffffffffu, any && lane_ok)) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
              const bool f = fl[u] && lane_ok && (t0 + u * NWARPS < t_end);
              const unsigned bal = __ballot_sync(0xffffffffu, f);
              if (f) wl[nlist + __popc(bal & ((1u << lane) - 1u))] =
                         (unsigned short)(((t0 + u * NWARPS) << 5) | lane);
              nlist += __popc(bal);
            }
            while (nlist >= 32) {
              nlist -= 32;
              drain(nlist, 32);
            }
          }
        }
        if (inplane && warp * 32 < NEC) {
          // x1 end cells: cell -1 gives the L state of face 0, cell txe the R state of face txe
          const int ec = min(threadIdx.x, NEC - 1);
          const int side = ec & 1, j = (ec >> 1) % TY, n = (ec >> 1) / TY;
          const bool ok = threadIdx.x < NEC && j < tye;
          const double *q = prim + (long)var_of_slot(n, 0) * g.ncell + plane + (long)min(j, tye - 1) * sj +
                            (side ? txe : -1);
          double v1[1][5], ql1[1], qr1[1];
          load_stencil<RECON>(q, 1, v1[0]);
          bool f1[1];
          const bool f = recon_ppm_fastpath<1>(v1, ql1, qr1, f1) && ok;
          if (ok && !f) {
            if (side) sm.wrx[n][j][txe] = qr1[0];
            else sm.wlx[n][j][0] = ql1[0];
          }
          const unsigned bal = __ballot_sync(0xffffffffu, f);
          if (f) wl[nlist + __popc(bal & ((1u << lane) - 1u))] =
                     (unsigned short)(((NT + 1 + warp) << 5) | lane);
          nlist += __popc(bal);
        }
        while (nlist > 0) {
          const int n = min(nlist, 32);
          nlist -= n;
          drain(nlist, n);
        }
      } else {
      struct Batch {
        double v[UNROLL][5];
        int sl[UNROLL], sr[UNROLL], st[UNROLL];
      };
      constexpr bool LIST = (RECON == APK_RECON_PPM) && APK_FUSED_PPM_LIST;
      [[maybe_unused]] int nlist = 0;  // entries in this warp's work list (warp uniform)
      [[maybe_unused]] unsigned short *const wl = sm.wl[DC ? 0 : warp];
      // full PPM for one listed cell per lane (entries [base, base + 32) of the list)
      [[maybe_unused]] auto drain = [&](int base, int n) {
        __syncwarp();
        if (lane < n) {
          const unsigned ent = wl[base + lane];
          const int tr = ent >> 5, ln = ent & 31;
          const int4 e = *reinterpret_cast<const int4 *>(&sm.tab[tr]);
          const double *q = prim + plane + ln + e.x;
          if constexpr (TMA) {
            if (tr >= NZ) q = &sm.tile[0][0][0] + ln + e.x;
          }
          double s5[5], ql, qr;
          load_stencil<RECON>(q, e.y & 0x0fffffff, s5);
          recon_ppm(s5[0], s5[1], s5[2], s5[3], s5[4], ql, qr);
          if (e.z >= 0) smd[e.z + (tr < NZ ? zoff : 0) + ln] = ql;
          if (e.w >= 0) smd[e.w + ln] = qr;
        }
        __syncwarp();
      };
      auto fetch = [&](Batch &b, int t0) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int tr = t0 + u * NWARPS;
          const int4 e = *reinterpret_cast<const int4 *>(&sm.tab[min(tr, t_end - 1)]);
          const bool ok = lane_ok && tr < t_end;
          // (rows beyond the tile have e.z == -1: the x3 double-buffer offset must not revive them)
          b.sl[u] = (ok && e.z >= 0) ? e.z + (tr < NZ ? zoff : 0) : -1;
          b.sr[u] = ok ? e.w : -1;
          b.st[u] = e.y;
          if constexpr (TMA) {
            // (the clamped task decides: on x3-only planes out-of-range slots repeat an x3 task)
            if (min(tr, t_end - 1) < NZ) load_stencil<RECON>(pbase + e.x, e.y & 0x0fffffff, b.v[u]);
            else load_stencil<RECON>(tile + e.x, e.y & 0x0fffffff, b.v[u]);
          } else {
            load_stencil<RECON>(pbase + e.x, e.y & 0x0fffffff, b.v[u]);
          }
        }
      };
      auto evaluate = [&](const Batch &b, [[maybe_unused]] int t0) {
        double qls[UNROLL], qrs[UNROLL];
        if constexpr (LIST) {
          bool fl[UNROLL];
          recon_ppm_fastpath<UNROLL>(b.v, qls, qrs, fl);
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            // only cells whose result is stored can be listed (flagged lanes are < txe: lane == ln)
            const bool f = fl[u] && (b.sl[u] >= 0 || b.sr[u] >= 0);
            if (b.sl[u] >= 0 && !f) smd[b.sl[u] + lane] = qls[u];
            if (b.sr[u] >= 0 && !f) smd[b.sr[u] + lane] = qrs[u];
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (bal) {
              if (f) wl[nlist + __popc(bal & ((1u << lane) - 1u))] =
                         (unsigned short)(((t0 + u * NWARPS) << 5) | lane);
              nlist += __popc(bal);
            }
          }
          while (nlist >= 32) {
            nlist -= 32;
            drain(nlist, 32);
          }
          return;
        }
        if constexpr (RECON == APK_RECON_PPM && APK_FUSED_PPM_VOTE) {
          recon_ppm_warp<UNROLL>(b.v, qls, qrs);  // all 32 lanes are here: the loop is warp uniform
        } else {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            const int dir = (b.st[u] >> 28) & 3;
            recon_stencil<RECON>(b.v[u], dir == 0 ? dx1 : (dir == 1 ? dx2 : dx3),
                                 (b.st[u] >> 30) & 1, qls[u], qrs[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          if (b.sl[u] >= 0) smd[b.sl[u] + lane] = qls[u];
          if (b.sr[u] >= 0) smd[b.sr[u] + lane] = qrs[u];
        }
      };
      if constexpr (APK_FUSED_SWPIPE) {
        Batch A, Bb;
        int t0 = warp;
        if (t0 < t_end) fetch(A, t0);
        while (t0 < t_end) {
          const bool more1 = t0 + STEP < t_end;
          if (more1) fetch(Bb, t0 + STEP);
          evaluate(A, t0);
          if (!more1) break;
          const bool more2 = t0 + 2 * STEP < t_end;
          if (more2) fetch(A, t0 + 2 * STEP);
          evaluate(Bb, t0 + STEP);
          t0 += 2 * STEP;
          if (!more2) break;
        }
      } else {
        for (int t0 = warp; t0 < t_end; t0 += STEP) {
          Batch A;
          fetch(A, t0);
          evaluate(A, t0);
        }
      }
      if constexpr (LIST) {
        if (nlist > 0) drain(0, nlist);
      }
      if (inplane) {
        // x1 end cells: cell -1 gives the L state of face 0, cell txe the R state of face txe
        for (int t = threadIdx.x; t < NH * TY * 2; t += NTHREADS) {
          const int side = t & 1, j = (t >> 1) % TY, n = (t >> 1) / TY;
          if (j < tye) {
            const int c = side ? txe : -1;
            const double *q = prim + (long)var_of_slot(n, 0) * g.ncell + plane + (long)j * sj + c;
            if constexpr (TMA) q = &sm.tile[var_of_slot(n, 0)][j + 3][c + 3];
            double ql, qr;
            recon_cell<RECON>(q, 1, dx1, (n == 0 || n == 4), ql, qr);
            if (side) sm.wrx[n][j][txe] = qr;
            else sm.wlx[n][j][0] = ql;
          }
        }
      }
      }
    }
    if constexpr (RECON != APK_RECON_DC) __syncthreads();
    if constexpr (TMA) {
      // every warp has finished reading the tile of plane kp: refill it for the next swept plane
      // while the Riemann and finalize phases run
      if (kp + 1 >= kbeg && kp + 1 <= kend && kp >= kbeg) tma_issue(kp + 1);
    }
    // ------------------------------------------------------------------ riemann phase
    if constexpr (RECON != APK_RECON_DC && APK_FUSED_PREFETCH) {
      // The Riemann phase touches no global memory: use it to pull everything the NEXT recon phase
      // reads for the first time into L1 -- the new top plane of the x3 stencils (kp+1+R) and,
      // if plane kp+1 is swept in x1/x2, its halo rows/columns (one prefetch per 128 B line;
      // ~2 per thread).  Measured before: 19 % of the recon loop was spent waiting for these
      // L2/HBM misses (first use after the stencil loads).
      const int pn = kp + 1;
      if (pn <= klast) {
        // (line offsets relative to plane pn were decoded once per CTA: pf_off)
        const double *pl = prim + (long)pn * sk;
        const bool next_in = (pn >= kbeg && pn <= kend);
#pragma unroll
        for (int i = 0; i < PFN; ++i) {
          // x3 stencil rows always; halo rows of the plane itself only if it is swept in x1 / x2
          if (pf_off[i] >= 0 && (next_in || (pf_zrow >> i & 1)))
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pl + pf_off[i]));
        }
      }
    }
    if (owner) {
      // the finalize phase of this iteration streams cons / u1 of plane kfin exactly once: start
      // the HBM fetch now so it overlaps the Riemann solves (12-16 warps/SM cannot hide it later)
      const int kfin = (NDIM == 3) ? kp - 1 : kp;
      if (kfin >= kbeg && kfin <= kend) {
        const long c = (long)kfin * sk + (long)(j0 + cj) * sj + (i0 + ci);
        const bool need_u0 = save_u1 || prim_only || u1_is_cons || (g0 != 0.0);
        const bool need_u1 = !(save_u1 || prim_only || u1_is_cons);
#pragma unroll
        for (int v = 0; v < NH; ++v) {
          if (need_u0) asm volatile("prefetch.global.L1 [%0];" ::"l"(cons + (long)v * g.ncell + c));
          if (need_u1) asm volatile("prefetch.global.L1 [%0];" ::"l"(cons_u1 + (long)v * g.ncell + c));
        }
      }
    }
    {
      // warp-task list: [0,TY) x1 rows, TY: x1 last faces, [TY+1, 2TY+2) x2 face rows,
      // [2TY+2, 3TY+2) x3 rows
      const int ntask = 3 * TY + 2;
      for (int t = warp; t < ntask; t += NWARPS) {
        double L[NH], R[NH], F[NH];
        double *dst = nullptr;
        long dstride = 0;
        if (t <= TY) {
          if (!inplane) continue;
          int j, f;
          if (t < TY) { j = t; f = lane; if (f > txe || f >= TX) continue; }
          else { j = lane; f = TX; if (txe < TX || j >= TY) continue; }
          if (j >= tye) continue;
          if constexpr (RECON == APK_RECON_DC) {
            const int pb = kp % 3, pr = (NDIM >= 2) ? j + 1 : 1;
#pragma unroll
            for (int n = 0; n < NH; ++n) {
              const int vs = var_of_slot(n, 0);
              L[n] = sm.pl[pb][vs][pr][f];
              R[n] = sm.pl[pb][vs][pr][f + 1];
            }
          } else {
#pragma unroll
            for (int n = 0; n < NH; ++n) { L[n] = sm.wlx[n][j][f]; R[n] = sm.wrx[n][j][f]; }
          }
          dst = &sm.wlx[0][j][f];
          dstride = TY * (TX + 1);
        } else if (t < 2 * TY + 2) {
          if (!inplane || NDIM < 2) continue;
          const int r = t - (TY + 1);
          if (r > tye || lane >= txe) continue;
          if constexpr (RECON == APK_RECON_DC) {
            const int pb = kp % 3;
#pragma unroll
            for (int n = 0; n < NH; ++n) {
              const int vs = var_of_slot(n, 1);
              L[n] = sm.pl[pb][vs][r][lane + 1];
              R[n] = sm.pl[pb][vs][r + 1][lane + 1];
            }
          } else {
#pragma unroll
            for (int n = 0; n < NH; ++n) { L[n] = sm.wly[n][r][lane]; R[n] = sm.wry[n][r][lane]; }
          }
          dst = &sm.wly[0][r][lane];
          dstride = (TY + 1) * TX;
        } else {
          if (NDIM < 3 || kp < kbeg) continue;
          const int j = t - (2 * TY + 2);
          if (j >= tye || lane >= txe) continue;
          if constexpr (RECON == APK_RECON_DC) {
            const int pb = kp % 3, pa = (kp + 2) % 3, pr = (NDIM >= 2) ? j + 1 : 1;  // planes k, k-1
#pragma unroll
            for (int n = 0; n < NH; ++n) {
              const int vs = var_of_slot(n, 2);
              L[n] = sm.pl[pa][vs][pr][lane + 1];
              R[n] = sm.pl[pb][vs][pr][lane + 1];
            }
          } else {
#pragma unroll
            for (int n = 0; n < NH; ++n) { L[n] = sm.wlz[cur].a[n][j][lane]; R[n] = sm.wrz[n][j][lane]; }
          }
          dst = &sm.wlz[cur].a[0][j][lane];
          dstride = TY * TX;
        }
        if constexpr (NS == 0) {
          riemann_solve<FLUID, RS>(L, R, rsp, F);
        } else {
          double Lh[NHY], Rh[NHY], Fh[NHY];
#pragma unroll
          for (int n = 0; n < NHY; ++n) { Lh[n] = L[n]; Rh[n] = R[n]; }
          riemann_solve<FLUID, RS>(Lh, Rh, rsp, Fh);
#pragma unroll
          for (int n = 0; n < NHY; ++n) F[n] = Fh[n];
          // passive scalars: upwinded by the sign of the mass flux (hydro.cpp:1089-1097)
#pragma unroll
          for (int n = NHY; n < NH; ++n) F[n] = (Fh[0] >= 0.0) ? Fh[0] * L[n] : Fh[0] * R[n];
        }
#pragma unroll
        for (int n = 0; n < NH; ++n) dst[n * dstride] = F[n];
      }
    }
    __syncthreads();
    // ------------------------------------------------------------------ finalize phase
    if (owner) {
      const int kfin = (NDIM == 3) ? kp - 1 : kp;  // plane whose cells can be finished now
      if (NDIM < 3) {
#pragma unroll
        for (int v = 0; v < NH; ++v) {
          const int sx = slot_of_var(v, 0);
          du[v] = (a1 * sm.wlx[sx][cj][ci + 1] - a1 * sm.wlx[sx][cj][ci]);
          if (NDIM >= 2) {
            const int sy = slot_of_var(v, 1);
            du[v] += (a2 * sm.wly[sy][cj + 1][ci] - a2 * sm.wly[sy][cj][ci]);
          }
        }
      }
      if (kfin >= kbeg && kfin <= kend) {
        const long c = (long)kfin * sk + (long)(j0 + cj) * sj + (i0 + ci);
        double u[NH], w[NH], c0[NH], c1[NH];
        // batch the global loads (uniform branch: stage 1 saves u1 <- u0 instead of reading it,
        // hydro_driver.cpp:474-495 folded into this pass)
        // gam0 == 0 (every VL2 stage, first stage of all integrators): 0*u0 contributes an exact
        // zero, so u0 need not be read at all unless it is also the source of the u1 copy
        // VL2 special case (APK_FUSED_PRIM_ONLY / APK_FUSED_U1_FROM_CONS): the predictor's cons is
        // never read again (gam0 == 0 in the corrector), so stage 1 stores only prim_out and leaves
        // cons == u^n in place, and stage 2 takes u1 from cons: 18 stores per cell fewer
        const bool src_cons = save_u1 || prim_only || u1_is_cons;  // u1 term comes from cons
        const bool need_u0 = src_cons || (g0 != 0.0);
        double cc[NH];
#pragma unroll
        for (int v = 0; v < NH; ++v) cc[v] = need_u0 ? cons[(long)v * g.ncell + c] : 0.0;
#pragma unroll
        for (int v = 0; v < NH; ++v) c0[v] = (g0 != 0.0 || save_u1 || prim_only) ? cc[v] : 0.0;
        if (src_cons) {
#pragma unroll
          for (int v = 0; v < NH; ++v) {
            c1[v] = cc[v];
            if (save_u1 && !prim_only) cons_u1[(long)v * g.ncell + c] = cc[v];
          }
        } else {
#pragma unroll
          for (int v = 0; v < NH; ++v) c1[v] = cons_u1[(long)v * g.ncell + c];
        }
#pragma unroll
        for (int v = 0; v < NH; ++v) {
          double d = du[v];
          if (NDIM == 3) {
            const int sz = slot_of_var(v, 2);
            d += (a3 * sm.wlz[cur].a[sz][cj][ci] - a3 * fzlo[v]);
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
          if constexpr (NS == 0) fofc_bad = fofc_needs_correction<NH>(u);  // (host: no FOFC with scalars)
          if (fofc_bad) atomicAdd(fofc_count, 1ULL);
        }
        if constexpr (FLUID == APK_FLUID_GLMMHD) u[8] *= dedner_coeff;  // dedner_source.cpp:72
        int err;
        if constexpr (NS == 0) {
          err = cons_to_prim<NH>(u, w, eos);
        } else {
          double uh[NHY], wh[NHY];
#pragma unroll
          for (int v = 0; v < NHY; ++v) uh[v] = u[v];
          err = cons_to_prim<NHY>(uh, wh, eos);
#pragma unroll
          for (int v = 0; v < NHY; ++v) { u[v] = uh[v]; w[v] = wh[v]; }
          // scalars: s = (rho s) / rho with the floored density (adiabatic_hydro.hpp:137-141)
          const double di = frcp(uh[0]);
#pragma unroll
          for (int v = NHY; v < NH; ++v) w[v] = u[v] * di;
        }
        // PARTHENON_REQUIRE of ConsToPrim (a cell FOFC flagged is redone by the caller instead)
        if (err && status && !fofc_bad) atomicMax(status, err);
#pragma unroll
        for (int v = 0; v < NH; ++v) {
          const long o = (long)v * g.ncell + c;
          if (!prim_only) cons[o] = u[v];
          po[o] = w[v];
        }
        if (flags & (APK_FUSED_PUSH_PRIM | APK_FUSED_PUSH_CONS)) {
          // Same-device ghost exchange folded into the epilogue ("push"): a finished cell within
          // ng of a block face is also the source of ghost cells of up to 7 neighbour blocks
          // (sender side of bnd_info.cpp:136-230).  The neighbours' prim_out / cons ghost cells
          // are not read by any CTA of this launch, so the stores are race free and the separate
          // fill-ghosts pass (2 x 1 ms per cycle at 256^3) disappears.
          const int gi = i0 + ci, gj = j0 + cj;
          const bool lo1 = gi < g.is + g.ng, hi1 = gi > g.ie - g.ng;
          const bool lo2 = NDIM >= 2 && gj < g.js + g.ng, hi2 = NDIM >= 2 && gj > g.je - g.ng;
          const bool lo3 = NDIM == 3 && kfin < g.ks + g.ng, hi3 = NDIM == 3 && kfin > g.ke - g.ng;
          if (lo1 | hi1 | lo2 | hi2 | lo3 | hi3) {
            const long e1 = g.ie - g.is + 1, e2 = (long)(g.je - g.js + 1) * sj,
                       e3 = (long)(g.ke - g.ks + 1) * sk;
#pragma unroll 1
            for (int o3 = (NDIM == 3 ? -1 : 0); o3 <= (NDIM == 3 ? 1 : 0); ++o3) {
              if ((o3 < 0 && !lo3) || (o3 > 0 && !hi3)) continue;
#pragma unroll 1
              for (int o2 = (NDIM >= 2 ? -1 : 0); o2 <= (NDIM >= 2 ? 1 : 0); ++o2) {
                if ((o2 < 0 && !lo2) || (o2 > 0 && !hi2)) continue;
#pragma unroll 1
                for (int o1 = -1; o1 <= 1; ++o1) {
                  if ((o1 < 0 && !lo1) || (o1 > 0 && !hi1) || (o1 | o2 | o3) == 0) continue;
                  const int o = (o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1);
                  const long cd = c - (o3 * e3 + o2 * e2 + o1 * e1);
                  double *__restrict__ pd = sm.push_prim[o];
                  double *__restrict__ cdst = sm.push_cons[o];
                  if (pd) {  // null: physical boundary (k_bc), remote (k_halo_copy) or not requested
#pragma unroll
                    for (int v = 0; v < NH; ++v) pd[(long)v * g.ncell + cd] = w[v];
                  }
                  if (cdst) {
#pragma unroll
                    for (int v = 0; v < NH; ++v) cdst[(long)v * g.ncell + cd] = u[v];
                  }
                }
              }
            }
          }
        }
        if (dtmin) {  // hydro.cpp:845-895
          double lx, ly, lz;
          if constexpr (FLUID == APK_FLUID_EULER) {
            lx = ly = lz = sound_speed(rsp.gamma, w[0], w[4]);
          } else {
            lx = fast_speed(rsp.gamma, w[0], w[4], w[5], w[6], w[7]);
            ly = (NDIM > 1) ? fast_speed(rsp.gamma, w[0], w[4], w[6], w[7], w[5]) : 0.0;
            lz = (NDIM > 2) ? fast_speed(rsp.gamma, w[0], w[4], w[7], w[5], w[6]) : 0.0;
          }
          my_min = dmin(my_min, qdiv(dx1, (fabs(w[1]) + lx)));
          if (NDIM > 1) my_min = dmin(my_min, qdiv(dx2, (fabs(w[2]) + ly)));
          if (NDIM > 2) my_min = dmin(my_min, qdiv(dx3, (fabs(w[3]) + lz)));
        }
      }
      if (NDIM == 3 && inplane) {
#pragma unroll
        for (int v = 0; v < NH; ++v) {
          const int sx = slot_of_var(v, 0), sy = slot_of_var(v, 1), sz = slot_of_var(v, 2);
          du[v] = (a1 * sm.wlx[sx][cj][ci + 1] - a1 * sm.wlx[sx][cj][ci]);
          du[v] += (a2 * sm.wly[sy][cj + 1][ci] - a2 * sm.wly[sy][cj][ci]);
          fzlo[v] = sm.wlz[cur].a[sz][cj][ci];
        }
      }
    }
    if constexpr (DC) {
      asm volatile("cp.async.wait_all;" ::: "memory");  // the next plane has landed
    } else {
      cur ^= 1;
    }
    __syncthreads();
  }
  if (dtmin) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_min = fmin(my_min, __shfl_xor_sync(0xffffffffu, my_min, o));
    if (lane == 0 && my_min < DBL_MAX)
      atomicMin(dtmin, (unsigned long long)__double_as_longlong(my_min));
  }
}

// x3 chunks (blockIdx.z): the march of a tile along x3 is cut into `c` chunks so that small packs
// (strong scaling: 64 meshblocks per GPU at 8 GPUs) fill every SM for several waves; a chunk re-does
// one x3 reconstruction plane and one face row of its lower neighbour.  Rule fitted to measurements
// on one B200 with 32^3 meshblocks (ms per cycle; T = tiles, 444 resident CTA slots):
//   T =  512 (8-GPU share of 256^3): c = 2: ~1.30   4: 1.209   5: 1.239   6: 1.235   8: 1.248
//   T = 1000 (~4-GPU share):         c = 2: 2.659   4: 2.661   8: 2.795
//   T = 1728 (~2-GPU share):         c = 1: 4.411   2: 4.369   4: 4.430
//   T = 4096 (one GPU):              c = 2: 8.599   3: 8.698
// -> equal chunks only (c divides nx3; 3, 5, 6 chunks of 32 planes are unequal and slower), at least
// 8 planes per chunk, at least 2 chunks when that allows, and the smallest such c that gives
// >= 7.5 waves of CTAs (else the largest).  The previous rule chose 8 chunks for T <= 1024
// (3-5 % slower).  APK_B200_KCHUNKS overrides.
inline int pick_x3_chunks(long tiles, long slots, int nx3) {
  static const char *env = std::getenv("APK_B200_KCHUNKS");
  if (env && std::atoi(env) > 0) return std::min(std::atoi(env), nx3);
  int cmax = 1;
  for (int c = 2; c <= 8; c *= 2)
    if (nx3 % c == 0 && nx3 / c >= 8) cmax = c;
  for (int c = (cmax >= 2 ? 2 : 1); c < cmax; c *= 2)
    if (2 * tiles * c >= 15 * slots) return c;
  return cmax;
}

template <int FLUID, int RECON, int RS, int NS = 0>
int run_fused(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
              double *const *prim_out, const apk_stage_args_t &a, double coeff, cudaStream_t s) {
  const double g0 = a.gam0, g1 = a.gam1, bdt = a.beta_dt, c_h = a.c_h;
  const int flags = a.flags;
  double *d_dtmin = a.d_dtmin;
  const apk_push_dst_t *push = a.d_push;
  constexpr int NH = ((FLUID == APK_FLUID_GLMMHD) ? 9 : 5) + NS;
  const EosPars eos = {d.gamma, d.dfloor, d.pfloor, d.efloor, d.vceil, d.eceil};
  // IEEE quotients evaluated once on the host (== frcp / fdiv on the device)
  const RsPars rsp = {d.gamma, 1.0 / (d.gamma - 1.0), c_h, 0.5 / c_h};
  const int ntx = (d.nx[0] + TX - 1) / TX;
  const int nty = (g.ndim >= 2) ? (d.nx[1] + TY - 1) / TY : 1;
  const int sm_count = apk_common::sm_count();
  constexpr bool DC = (RECON == APK_RECON_DC);
  static_assert(!DC || 3 * (sizeof(Smem<9, 3, DC>) + 1024) <= 233472 || APK_FUSED_MINBLOCKS < 3 ||
                    TY != 4, "DC stage must keep 3 CTAs per SM");
  const size_t smem = (g.ndim == 3) ? sizeof(Smem<NH, 3, DC>) : (g.ndim == 2 ? sizeof(Smem<NH, 2, DC>) : sizeof(Smem<NH, 1, DC>));
  unsigned long long *dt = reinterpret_cast<unsigned long long *>(d_dtmin);
#define APK_LAUNCH_FUSED(ND)                                                                     \
  {                                                                                              \
    auto kern = k_stage_fused<FLUID, RECON, RS, ND, NS>;                                         \
    /* per-device attribute (a process may drive several GPUs): set once per device */          \
    static unsigned long long attr_set = 0;                                                      \
    const int dev = apk_common::current_device();                                                \
    if (dev >= 64 || !(attr_set >> dev & 1ULL)) {                                                \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
      if (dev < 64) attr_set |= 1ULL << dev;                                                     \
    }                                                                                            \
    static int occ_cache[64] = {0};                                                              \
    int occ = (dev < 64) ? occ_cache[dev] : 0;                                                   \
    if (occ == 0) {                                                                              \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHREADS, smem);                 \
      if (occ < 1) occ = 1;                                                                      \
      if (dev < 64) occ_cache[dev] = occ;                                                        \
    }                                                                                            \
    const int nchunk = (ND == 3) ? pick_x3_chunks((long)ntx * nty * g.nb, (long)sm_count * occ,  \
                                                  d.nx[2]) : 1;                                  \
    dim3 grid(ntx * nty, g.nb, nchunk);                                                          \
    kern<<<grid, NTHREADS, smem, s>>>(g, blocks, prim_out, g0, g1, bdt, rsp, coeff, eos,         \
                                      ntx, flags, dt, push, nchunk, a.d_status,                  \
                                      a.d_fofc_count,                                            \
                                      static_cast<const unsigned char *>(a.d_tma_maps));         \
  }
  if (g.ndim == 3) APK_LAUNCH_FUSED(3)
  else if (g.ndim == 2) APK_LAUNCH_FUSED(2)
  else APK_LAUNCH_FUSED(1)
#undef APK_LAUNCH_FUSED
  apk_common::count_launch();
  return 0;
}

template <int FLUID, int RS, int NS = 0>
int fused_recon(int recon, const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                double *const *prim_out, const apk_stage_args_t &a, double coeff, cudaStream_t s) {
  switch (recon) {
  case APK_RECON_DC: return run_fused<FLUID, APK_RECON_DC, RS, NS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_PLM: return run_fused<FLUID, APK_RECON_PLM, RS, NS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_PPM: return run_fused<FLUID, APK_RECON_PPM, RS, NS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_WENOZ: return run_fused<FLUID, APK_RECON_WENOZ, RS, NS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_WENO3: return run_fused<FLUID, APK_RECON_WENO3, RS, NS>(d, g, blocks, prim_out, a, coeff, s);
  case APK_RECON_LIMO3: return run_fused<FLUID, APK_RECON_LIMO3, RS, NS>(d, g, blocks, prim_out, a, coeff, s);
  }
  apk_common::set_error("apk_stage_fused: unknown reconstruction");
  return 1;
}

}  // namespace

#ifndef APK_FUSED_PART
#define APK_FUSED_PART 0
#endif
int launch_stage_rows_euler(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                            double *const *prim_out, const apk_stage_args_t &a, double coeff,
                            int recon, cudaStream_t s);
int launch_stage_rows_mhd(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                          double *const *prim_out, const apk_stage_args_t &a, double coeff,
                          int recon, cudaStream_t s);
int launch_stage_fused_ns1(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                           double *const *prim_out, const apk_stage_args_t &a, double coeff,
                           int recon, cudaStream_t s);
int launch_stage_fused_ns2(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                           double *const *prim_out, const apk_stage_args_t &a, double coeff,
                           int recon, cudaStream_t s);
#if defined(APK_FUSED_NS) && APK_FUSED_NS > 0
// translation unit of the kernels with APK_FUSED_NS passive scalars (both fluids)
#define APK_CAT2(a, b) a##b
#define APK_CAT(a, b) APK_CAT2(a, b)
int APK_CAT(launch_stage_fused_ns, APK_FUSED_NS)(const apk_pack_desc_t &d, const ApkGeom &g,
                                                 const apk_block_t *blocks, double *const *prim_out,
                                                 const apk_stage_args_t &a, double coeff, int recon,
                                                 cudaStream_t s) {
  constexpr int NS = APK_FUSED_NS;
  if (d.fluid == APK_FLUID_EULER) {
    switch (d.riemann) {
    case APK_RIEMANN_NONE: return run_fused<APK_FLUID_EULER, APK_RECON_DC, APK_RIEMANN_NONE, NS>(d, g, blocks, prim_out, a, coeff, s);
    case APK_RIEMANN_HLLE: return fused_recon<APK_FLUID_EULER, APK_RIEMANN_HLLE, NS>(recon, d, g, blocks, prim_out, a, coeff, s);
    case APK_RIEMANN_HLLC: return fused_recon<APK_FLUID_EULER, APK_RIEMANN_HLLC, NS>(recon, d, g, blocks, prim_out, a, coeff, s);
    case APK_RIEMANN_LLF: return run_fused<APK_FLUID_EULER, APK_RECON_DC, APK_RIEMANN_LLF, NS>(d, g, blocks, prim_out, a, coeff, s);
    }
  } else {
    switch (d.riemann) {
    case APK_RIEMANN_NONE: return run_fused<APK_FLUID_GLMMHD, APK_RECON_DC, APK_RIEMANN_NONE, NS>(d, g, blocks, prim_out, a, coeff, s);
    case APK_RIEMANN_HLLE: return fused_recon<APK_FLUID_GLMMHD, APK_RIEMANN_HLLE, NS>(recon, d, g, blocks, prim_out, a, coeff, s);
    case APK_RIEMANN_HLLD: return fused_recon<APK_FLUID_GLMMHD, APK_RIEMANN_HLLD, NS>(recon, d, g, blocks, prim_out, a, coeff, s);
    case APK_RIEMANN_LLF: return run_fused<APK_FLUID_GLMMHD, APK_RECON_DC, APK_RIEMANN_LLF, NS>(d, g, blocks, prim_out, a, coeff, s);
    }
  }
  apk_common::set_error("apk_stage_fused: unsupported fluid/riemann combination");
  return 1;
}
#elif APK_FUSED_TMA
static_assert(TX == APK_TMA_TX && TY == APK_TMA_TY, "tensor-map box and kernel tile must agree");
// the A/B variant is built for the headline method only: GLM-MHD, PPM + HLLD (3-D meshes)
int launch_stage_fused(const apk_pack_desc_t &d, const apk_block_t *blocks,
                       double *const *prim_out, const apk_stage_args_t &a, double dedner_coeff,
                       cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  return run_fused<APK_FLUID_GLMMHD, APK_RECON_PPM, APK_RIEMANN_HLLD>(d, g, blocks, prim_out, a,
                                                                     dedner_coeff, s);
}
#elif APK_FUSED_PART == 0
int launch_stage_fused_mhd(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                           double *const *prim_out, const apk_stage_args_t &a,
                           double dedner_coeff, int recon, cudaStream_t s);

int launch_stage_fused(const apk_pack_desc_t &d, const apk_block_t *blocks,
                       double *const *prim_out, const apk_stage_args_t &a, double dedner_coeff,
                       cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  const int recon = (a.recon_override < 0) ? d.recon : a.recon_override;
  if (d.nscalars != 0) {
    // kernels with passive scalars are built for 1 and 2 scalars (own translation units)
    if (a.flags & APK_FUSED_FOFC_CHECK) {
      apk_common::set_error("apk_stage_fused: FOFC with passive scalars needs the un-fused tasks");
      return 1;
    }
    if (d.nscalars == 1) return launch_stage_fused_ns1(d, g, blocks, prim_out, a, dedner_coeff, recon, s);
    if (d.nscalars == 2) return launch_stage_fused_ns2(d, g, blocks, prim_out, a, dedner_coeff, recon, s);
    apk_common::set_error("apk_stage_fused: more than 2 passive scalars need the un-fused tasks");
    return 1;
  }
  // APK_B200_FUSED_V=2 selects the row-task kernel (apk_rows.cu, 2-D / 3-D meshes) for A/B runs:
  // bit-identical results, 3 % fewer instructions, but 34 % slower at 256^3 -- its 115 KB of
  // straight-line SASS per plane overflows the 32 KB L1.5 instruction cache (ncu: 2.1 warps stalled
  // on `no_instruction` per issue against 0.11 here; profiles/r02_notes.md)
  static const char *ver = std::getenv("APK_B200_FUSED_V");
  if (g.ndim >= 2 && ver && (ver[0] == '2' || (ver[0] == '4' && recon != APK_RECON_DC))) {
    if (d.fluid == APK_FLUID_EULER)
      return launch_stage_rows_euler(d, g, blocks, prim_out, a, dedner_coeff, recon, s);
    return launch_stage_rows_mhd(d, g, blocks, prim_out, a, dedner_coeff, recon, s);
  }
  if (d.fluid == APK_FLUID_EULER) {
    switch (d.riemann) {
    case APK_RIEMANN_NONE: return run_fused<APK_FLUID_EULER, APK_RECON_DC, APK_RIEMANN_NONE>(d, g, blocks, prim_out, a, dedner_coeff, s);
    case APK_RIEMANN_HLLE: return fused_recon<APK_FLUID_EULER, APK_RIEMANN_HLLE>(recon, d, g, blocks, prim_out, a, dedner_coeff, s);
    case APK_RIEMANN_HLLC: return fused_recon<APK_FLUID_EULER, APK_RIEMANN_HLLC>(recon, d, g, blocks, prim_out, a, dedner_coeff, s);
    // CalculateFluxesTight (hydro.cpp:980-1022): LLF fluxes of the neighbouring cell values = donor cell + LLF
    case APK_RIEMANN_LLF: return run_fused<APK_FLUID_EULER, APK_RECON_DC, APK_RIEMANN_LLF>(d, g, blocks, prim_out, a, dedner_coeff, s);
    }
  } else {
    return launch_stage_fused_mhd(d, g, blocks, prim_out, a, dedner_coeff, recon, s);
  }
  apk_common::set_error("apk_stage_fused: unsupported fluid/riemann combination");
  return 1;
}
#else
int launch_stage_fused_mhd(const apk_pack_desc_t &d, const ApkGeom &g, const apk_block_t *blocks,
                           double *const *prim_out, const apk_stage_args_t &a,
                           double dedner_coeff, int recon, cudaStream_t s) {
  switch (d.riemann) {
  case APK_RIEMANN_NONE: return run_fused<APK_FLUID_GLMMHD, APK_RECON_DC, APK_RIEMANN_NONE>(d, g, blocks, prim_out, a, dedner_coeff, s);
  case APK_RIEMANN_HLLE: return fused_recon<APK_FLUID_GLMMHD, APK_RIEMANN_HLLE>(recon, d, g, blocks, prim_out, a, dedner_coeff, s);
  case APK_RIEMANN_HLLD: return fused_recon<APK_FLUID_GLMMHD, APK_RIEMANN_HLLD>(recon, d, g, blocks, prim_out, a, dedner_coeff, s);
  case APK_RIEMANN_LLF: return run_fused<APK_FLUID_GLMMHD, APK_RECON_DC, APK_RIEMANN_LLF>(d, g, blocks, prim_out, a, dedner_coeff, s);
  }
  apk_common::set_error("apk_stage_fused: unsupported fluid/riemann combination");
  return 1;
}
#endif

}  // namespace APK_NS
