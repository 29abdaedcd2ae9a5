// apk_halo.cu -- ghost-cell machinery (pure data movement, independent of the FP mode).
//
// Replaces Parthenon's per-(block,neighbour,variable) pack -> CommBuffer -> unpack round trip
// (bvals/comms/boundary_communication.cpp:48-348, utils/communication_buffer.hpp) by
//  * k_fill_ghosts: same-device neighbours are read IN PLACE from the neighbour block's
//    interior (index ranges of bnd_info.cpp:105-230: the sender's ng interior layers adjacent
//    to the neighbour land in the receiver's ng exterior layers), one launch for all blocks,
//  * k_bc: outflow / reflecting physical BCs in Parthenon's face order over the full
//    transverse extent (bvals/boundary_conditions.cpp:47-56, boundary_conditions_generic.hpp
//    :174-262, src/bvals/boundary_conditions_apk.hpp:38-85),
//  * k_halo_pack / k_halo_unpack: ONE contiguous buffer per peer GPU holding every outgoing
//    slab (all variables), moved by one ncclSend/ncclRecv pair per peer.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "apk_launch.h"

namespace apk_common {

static std::atomic<long long> g_launches{0};
static std::mutex g_err_mutex;
static std::string g_err;

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
int sm_count() {
  static int cache[64] = {0};
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return 148;
  if (!cache[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = (n > 0) ? n : 148;
  }
  return cache[dev];
}
long long launches() { return g_launches.load(); }
void reset_launches() { g_launches.store(0); }
void set_error(const char *msg) {
  std::lock_guard<std::mutex> l(g_err_mutex);
  g_err = msg;
}
const char *last_error() { return g_err.c_str(); }

namespace {
constexpr int kThreads = 256;

// ghost cell -> neighbour offset in one dimension: -1 below, 0 interior, +1 above
__device__ __forceinline__ int side(int idx, int s, int e) { return (idx < s) ? -1 : (idx > e) ? 1 : 0; }

// One thread per GHOST cell (interior cells get no thread).  Ghost cells of a block are enumerated
// in memory order in two groups so that warps stay dense: (a) rows (k,j) that lie entirely in the
// ghost region (all n1 cells), (b) the 2*ng end cells of the rows that cross the interior.
// Every thread first LOADS all variables of its cell, then stores them (the per-variable planes
// are ncell doubles apart: 9..18 independent 256 B-coalesced streams in flight per warp).
struct GhostEnum {
  int full_rows;      // rows entirely ghost
  int kghost_rows;    // of which: rows of the k-ghost planes (2*ng*n2), 0 below 3-D
  int part_rows;      // rows crossing the interior (nx2*nx3)
  long per_block;     // ghost cells per block
};

__device__ __forceinline__ void ghost_decode(const ApkGeom &g, const GhostEnum &e, int nx2, long gi,
                                             int &k, int &j, int &i) {
  const long nfull = (long)e.full_rows * g.n1;
  if (gi < nfull) {
    const int fr = (int)(gi / g.n1);
    i = (int)(gi - (long)fr * g.n1);
    if (fr < e.kghost_rows) {  // planes k < ks or k > ke, all rows
      int kk = fr / g.n2;
      j = fr - kk * g.n2;
      k = (kk < g.ng) ? kk : g.ke + 1 + (kk - g.ng);
    } else {                   // interior planes, rows j < js or j > je
      const int r = fr - e.kghost_rows;
      const int per_plane = 2 * g.ng;
      const int kk = r / per_plane, jj = r - kk * per_plane;
      k = g.ks + kk;
      j = (jj < g.ng) ? jj : g.je + 1 + (jj - g.ng);
    }
  } else {
    const long p = gi - nfull;
    const int row = (int)(p / (2 * g.ng)), ii = (int)(p - (long)row * 2 * g.ng);
    i = (ii < g.ng) ? ii : g.ie + 1 + (ii - g.ng);
    const int kk = row / nx2;
    k = g.ks + kk;
    j = g.js + (row - kk * nx2);
  }
}

__global__ void __launch_bounds__(kThreads)
k_fill_ghosts(ApkGeom g, GhostEnum e, const apk_block_t *__restrict__ blocks, int nx1, int nx2,
              int nx3, int with_prim) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= e.per_block * g.nb) return;
  const int b = (int)(t / e.per_block);
  int k, j, i;
  ghost_decode(g, e, nx2, t - (long)b * e.per_block, k, j, i);
  const int o1 = side(i, g.is, g.ie);
  const int o2 = (g.ndim >= 2) ? side(j, g.js, g.je) : 0;
  const int o3 = (g.ndim >= 3) ? side(k, g.ks, g.ke) : 0;
  const apk_block_t &B = blocks[b];
  const int nb = B.nbr[(o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1)];
  if (nb < 0) return;  // physical boundary or remote (filled by k_halo_unpack)
  const apk_block_t &N = blocks[nb];
  const long c = ((long)k * g.n2 + j) * g.n1 + i;
  const long cs = ((long)(k - o3 * nx3) * g.n2 + (j - o2 * nx2)) * g.n1 + (i - o1 * nx1);
  // with_prim: 0 = cons only, 1 = cons + prim, 2 = prim only
  for (int f = (with_prim == 2) ? 1 : 0; f < ((with_prim == 0) ? 1 : 2); ++f) {
    const double *__restrict__ src = (f ? N.prim : N.cons) + cs;
    double *__restrict__ dst = (f ? B.prim : B.cons) + c;
    for (int n0 = 0; n0 < g.nvar; n0 += 9) {
      double tmp[9];
#pragma unroll
      for (int q = 0; q < 9; ++q)
        if (n0 + q < g.nvar) tmp[q] = src[(long)(n0 + q) * g.ncell];
#pragma unroll
      for (int q = 0; q < 9; ++q)
        if (n0 + q < g.nvar) dst[(long)(n0 + q) * g.ncell] = tmp[q];
    }
  }
}

// physical BC of one face (0..5) for all blocks whose bc[face] is outflow/reflect
__global__ void __launch_bounds__(kThreads)
k_bc(ApkGeom g, const apk_block_t *__restrict__ blocks, int face, int with_prim) {
  const int d = face >> 1, outer = face & 1;
  int lo[3] = {0, 0, 0}, ext[3] = {g.n1, g.n2, g.n3};
  const int s[3] = {g.is, g.js, g.ks}, e[3] = {g.ie, g.je, g.ke};
  lo[d] = outer ? e[d] + 1 : 0;
  ext[d] = g.ng;
  const long per = (long)ext[0] * ext[1] * ext[2];
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= per * g.nb) return;
  const int b = (int)(t / per);
  const apk_block_t &B = blocks[b];
  const int type = B.bc[face];
  if (type != APK_BC_OUTFLOW && type != APK_BC_REFLECT) return;
  long rem = t - (long)b * per;
  int idx[3];
  idx[2] = lo[2] + (int)(rem / ((long)ext[0] * ext[1]));
  rem -= (long)(idx[2] - lo[2]) * ext[0] * ext[1];
  idx[1] = lo[1] + (int)(rem / ext[0]);
  idx[0] = lo[0] + (int)(rem - (long)(idx[1] - lo[1]) * ext[0]);
  const int ref = outer ? e[d] : s[d];
  int src[3] = {idx[0], idx[1], idx[2]};
  src[d] = (type == APK_BC_REFLECT) ? (2 * ref + (outer ? 1 : -1)) - idx[d] : ref;
  const long c = ((long)idx[2] * g.n2 + idx[1]) * g.n1 + idx[0];
  const long cs = ((long)src[2] * g.n2 + src[1]) * g.n1 + src[0];
  for (int n = 0; n < g.nvar; ++n) {
    // AthenaPK ReflectBC flips variable index == DIR, i.e. the normal momentum / velocity
    const double sgn = (type == APK_BC_REFLECT && n == d + 1) ? -1.0 : 1.0;
    if (with_prim != 2) B.cons[(long)n * g.ncell + c] = sgn * B.cons[(long)n * g.ncell + cs];
    if (with_prim) B.prim[(long)n * g.ncell + c] = sgn * B.prim[(long)n * g.ncell + cs];
  }
}

// slab <-> contiguous buffer. Layout inside one item: [field(cons,prim)][var][k][j][i]
template <bool PACK>
__global__ void __launch_bounds__(kThreads)
k_halo_copy(ApkGeom g, const apk_block_t *__restrict__ blocks,
            const apk_halo_item_t *__restrict__ items, int with_prim, double *buf) {
  const apk_halo_item_t it = items[blockIdx.y];
  const long cells = (long)it.ext[0] * it.ext[1] * it.ext[2];
  const int nf = (with_prim == 1) ? 2 : 1;
  const long total = cells * g.nvar * nf;
  const apk_block_t &B = blocks[it.block];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const int f = (int)(t / (cells * g.nvar));
    long rem = t - (long)f * cells * g.nvar;
    const int n = (int)(rem / cells);
    rem -= (long)n * cells;
    const int kk = (int)(rem / ((long)it.ext[0] * it.ext[1]));
    rem -= (long)kk * it.ext[0] * it.ext[1];
    const int jj = (int)(rem / it.ext[0]);
    const int ii = (int)(rem - (long)jj * it.ext[0]);
    const long c = ((long)(it.lo[2] + kk) * g.n2 + (it.lo[1] + jj)) * g.n1 + (it.lo[0] + ii);
    double *field = (f || with_prim == 2) ? B.prim : B.cons;
    // item offsets are sized for two fields; a single-field exchange packs densely at offset/2
    const long base = (nf == 1) ? it.offset / 2 : it.offset;
    if (PACK) buf[base + t] = field[(long)n * g.ncell + c];
    else field[(long)n * g.ncell + c] = buf[base + t];
  }
}

// box copies between arrays of different extents (multilevel exchange, flux correction)
__global__ void __launch_bounds__(kThreads)
k_box_copy(const apk_box_item_t *__restrict__ items, int nvar) {
  const apk_box_item_t it = items[blockIdx.y];
  const long cells = (long)it.ext[0] * it.ext[1] * it.ext[2];
  const long total = cells * nvar;
  const long ncs = (long)it.ns[0] * it.ns[1] * it.ns[2], ncd = (long)it.nd[0] * it.nd[1] * it.nd[2];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long)gridDim.x * blockDim.x) {
    const int n = (int)(t / cells);
    long rem = t - (long)n * cells;
    const int kk = (int)(rem / ((long)it.ext[0] * it.ext[1]));
    rem -= (long)kk * it.ext[0] * it.ext[1];
    const int jj = (int)(rem / it.ext[0]);
    const int ii = (int)(rem - (long)jj * it.ext[0]);
    const long cs = ((long)(it.s0[2] + kk) * it.ns[1] + (it.s0[1] + jj)) * it.ns[0] + (it.s0[0] + ii);
    const long cd = ((long)(it.d0[2] + kk) * it.nd[1] + (it.d0[1] + jj)) * it.nd[0] + (it.d0[0] + ii);
    it.dst[(long)n * ncd + cd] = it.src[(long)n * ncs + cs];
  }
}

__global__ void __launch_bounds__(kThreads)
k_save_u1(ApkGeom g, const apk_block_t *__restrict__ blocks) {
  const long per = g.ncell * g.nvar;
  const apk_block_t &B = blocks[blockIdx.y];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < per;
       t += (long)gridDim.x * blockDim.x)
    B.cons_u1[t] = B.cons[t];
}

// FP64 pipe probe: 8 independent DFMA chains per thread
__global__ void k_fp64_probe(int niter, double *out) {
  double a[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) a[q] = 1.0 + 1e-9 * (threadIdx.x + q);
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < niter; ++it) {
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = fma(a[q], m, c);
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += a[q];
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}

inline unsigned grid_for(long total, int threads) {
  return (unsigned)((total + threads - 1) / threads);
}
}  // namespace

int launch_fill_ghosts(const apk_pack_desc_t &d, const apk_block_t *blocks, int with_prim,
                       cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  GhostEnum e;
  const int nx2 = (g.ndim >= 2) ? d.nx[1] : 1, nx3 = (g.ndim >= 3) ? d.nx[2] : 1;
  e.kghost_rows = (g.ndim >= 3) ? 2 * g.ng * g.n2 : 0;
  e.full_rows = e.kghost_rows + ((g.ndim >= 2) ? 2 * g.ng * nx3 : 0);
  e.part_rows = nx2 * nx3;
  e.per_block = (long)e.full_rows * g.n1 + (long)e.part_rows * 2 * g.ng;
  k_fill_ghosts<<<grid_for(e.per_block * g.nb, kThreads), kThreads, 0, s>>>(
      g, e, blocks, d.nx[0], nx2, nx3, with_prim);
  count_launch();
  return 0;
}

int launch_physical_bcs(const apk_pack_desc_t &d, const apk_block_t *blocks,
                        unsigned face_mask, int with_prim, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  for (int face = 0; face < 2 * g.ndim; ++face) {
    if (!(face_mask & (1u << face))) continue;
    long ext[3] = {g.n1, g.n2, g.n3};
    ext[face >> 1] = g.ng;
    k_bc<<<grid_for(ext[0] * ext[1] * ext[2] * g.nb, kThreads), kThreads, 0, s>>>(g, blocks, face,
                                                                                 with_prim);
    count_launch();
  }
  return 0;
}

int launch_halo_pack(const apk_pack_desc_t &d, const apk_block_t *blocks,
                     const apk_halo_item_t *items, int nitems, int with_prim, double *buf,
                     cudaStream_t s) {
  if (nitems <= 0) return 0;
  const ApkGeom g = apk_make_geom(d);
  dim3 grid(8, nitems);
  k_halo_copy<true><<<grid, kThreads, 0, s>>>(g, blocks, items, with_prim, buf);
  count_launch();
  return 0;
}

int launch_halo_unpack(const apk_pack_desc_t &d, const apk_block_t *blocks,
                       const apk_halo_item_t *items, int nitems, int with_prim,
                       const double *buf, cudaStream_t s) {
  if (nitems <= 0) return 0;
  const ApkGeom g = apk_make_geom(d);
  dim3 grid(8, nitems);
  k_halo_copy<false><<<grid, kThreads, 0, s>>>(g, blocks, items, with_prim,
                                               const_cast<double *>(buf));
  count_launch();
  return 0;
}

int launch_box_copy(const apk_box_item_t *items, int nitems, int nvar, cudaStream_t s) {
  if (nitems <= 0) return 0;
  dim3 grid(4, nitems);
  k_box_copy<<<grid, kThreads, 0, s>>>(items, nvar);
  count_launch();
  return 0;
}

int launch_save_u1(const apk_pack_desc_t &d, const apk_block_t *blocks, cudaStream_t s) {
  const ApkGeom g = apk_make_geom(d);
  dim3 grid(grid_for(g.ncell * g.nvar, kThreads * 4), g.nb);
  k_save_u1<<<grid, kThreads, 0, s>>>(g, blocks);
  count_launch();
  return 0;
}

int launch_fp64_probe(int nb, int nt, int niter, double *out, cudaStream_t s) {
  k_fp64_probe<<<nb, nt, 0, s>>>(niter, out);
  count_launch();
  return 0;
}

}  // namespace apk_common
