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

// One thread per ghost cell of the entire (ghost-inflated) index range of every block; copies
// all variables of cons (and optionally prim) from the owning neighbour's interior.
__global__ void __launch_bounds__(kThreads)
k_fill_ghosts(ApkGeom g, const apk_block_t *__restrict__ blocks, int nx1, int nx2, int nx3,
              int with_prim) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.ncell * g.nb) return;
  const int b = (int)(t / g.ncell);
  const long c = t - (long)b * g.ncell;
  const int k = (int)(c / ((long)g.n1 * g.n2));
  const int j = (int)((c / g.n1) % g.n2);
  const int i = (int)(c % g.n1);
  const int o1 = side(i, g.is, g.ie);
  const int o2 = (g.ndim >= 2) ? side(j, g.js, g.je) : 0;
  const int o3 = (g.ndim >= 3) ? side(k, g.ks, g.ke) : 0;
  if (!(o1 | o2 | o3)) return;
  const apk_block_t &B = blocks[b];
  const int nb = B.nbr[(o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1)];
  if (nb < 0) return;  // physical boundary or remote (filled by k_halo_unpack)
  const apk_block_t &N = blocks[nb];
  const long cs = ((long)(k - o3 * nx3) * g.n2 + (j - o2 * nx2)) * g.n1 + (i - o1 * nx1);
  // with_prim: 0 = cons only, 1 = cons + prim, 2 = prim only
  if (with_prim != 2)
    for (int n = 0; n < g.nvar; ++n) B.cons[(long)n * g.ncell + c] = N.cons[(long)n * g.ncell + cs];
  if (with_prim)
    for (int n = 0; n < g.nvar; ++n) B.prim[(long)n * g.ncell + c] = N.prim[(long)n * g.ncell + cs];
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
    if (PACK) buf[it.offset + t] = field[(long)n * g.ncell + c];
    else field[(long)n * g.ncell + c] = buf[it.offset + t];
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
  k_fill_ghosts<<<grid_for(g.ncell * g.nb, kThreads), kThreads, 0, s>>>(g, blocks, d.nx[0],
                                                                       d.nx[1], d.nx[2], with_prim);
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
