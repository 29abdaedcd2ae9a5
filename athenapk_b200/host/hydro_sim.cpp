// hydro_sim.cpp -- see hydro_sim.hpp.  Host C++ only: device work goes through the C-ABI.
#include "hydro_sim.hpp"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <tuple>
#include <unordered_map>

namespace apk_host {

namespace {
void cuda_check(cudaError_t e, const char *what) {
  if (e != cudaSuccess) {
    // product path must fail loudly (no CPU fallback)
    throw std::runtime_error(std::string("apk_host: CUDA failure in ") + what + ": " +
                             cudaGetErrorString(e));
  }
}
void abi_check(int rc, const char *what) {
  if (rc != 0)
    throw std::runtime_error(std::string("apk_host: ") + what + " failed: " + apk_last_error());
}
cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// mesh/forest/logical_location.cpp:61-74
double IndexToSymmetrizedCoordinate(int index, int bloc, int nrange) {
  const std::int64_t noffset = index - (nrange) / 2;
  const std::int64_t noffset_ceil = index - (nrange + 1) / 2;
  return static_cast<double>(noffset + noffset_ceil + static_cast<std::int64_t>(bloc)) /
         (2.0 * nrange);
}
// defs.hpp:106-109
double SymmetrizedLogicalToActualPosition(double u, double xmin, double xmax) {
  return 0.5 * (xmin + xmax) + (u * xmax - u * xmin);
}
int MaximumPowerOf2Divisor(int in) { return in & (~(in - 1)); }  // utils/bit_hacks.hpp:91

std::uint64_t Morton(int x, int y, int z) {
  std::uint64_t m = 0;
  for (int bit = 0; bit < 20; ++bit) {
    m |= (std::uint64_t)((x >> bit) & 1) << (3 * bit);
    m |= (std::uint64_t)((y >> bit) & 1) << (3 * bit + 1);
    m |= (std::uint64_t)((z >> bit) & 1) << (3 * bit + 2);
  }
  return m;
}
}  // namespace

LowStorageIntegrator::LowStorageIntegrator(Integrator which) {
  switch (which) {
  case Integrator::rk1:
    nstages = 1; gam0 = {0.0}; gam1 = {1.0}; beta = {1.0};
    break;
  case Integrator::rk2:
    nstages = 2; gam0 = {0.0, 0.5}; gam1 = {1.0, 0.5}; beta = {1.0, 0.5};
    break;
  case Integrator::vl2:
    nstages = 2; gam0 = {0.0, 0.0}; gam1 = {1.0, 1.0}; beta = {0.5, 1.0};
    break;
  case Integrator::rk3:
    nstages = 3; gam0 = {0.0, 0.25, 2.0 / 3.0}; gam1 = {1.0, 0.75, 1.0 / 3.0};
    beta = {1.0, 0.25, 2.0 / 3.0};
    break;
  }
}

TaskStatus TaskList::Execute() {
  // Parthenon re-enqueues tasks that return `incomplete` (tasks/tasks.cpp:119-153); with a
  // single in-order list that is a poll loop on the head task.
  for (auto &t : tasks) {
    TaskStatus st;
    long polls = 0;
    while ((st = t.second()) == TaskStatus::incomplete) {
      if (++polls > 2000000000L) return TaskStatus::fail;
    }
    if (st == TaskStatus::fail) return st;
  }
  return TaskStatus::complete;
}

HydroSim::HydroSim(const SimConfig &cfg)
    : cfg_(cfg), integrator_(static_cast<Integrator>(cfg.integrator)) {
  if (!cfg_.plan_only) {
    int ndev = 0;
    cuda_check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount (is there a GPU?)");
    if (ndev == 0) throw std::runtime_error("apk_host: no CUDA device");
  }
  nhydro_ = (cfg_.fluid == APK_FLUID_GLMMHD) ? 9 : 5;
  nvar_ = nhydro_ + cfg_.nscalars;
  ndim_ = (cfg_.nx[2] > 1) ? 3 : ((cfg_.nx[1] > 1) ? 2 : 1);
  // hydro.cpp:317-339,444-447: ghost zones required by the reconstruction
  const int need = (cfg_.recon == APK_RECON_DC) ? 1
                   : (cfg_.recon == APK_RECON_PPM || cfg_.recon == APK_RECON_WENOZ) ? 3 : 2;
  if (cfg_.ng < need)
    throw std::runtime_error("AthenaPK hydro: Need more ghost zones for chosen reconstruction.");
  if (cfg_.riemann == APK_RIEMANN_LLF && cfg_.recon != APK_RECON_DC)
    throw std::runtime_error("LLF Riemann solver only implemented with DC reconstruction.");
  if (cfg_.fluid == APK_FLUID_EULER && cfg_.riemann == APK_RIEMANN_HLLD)
    throw std::runtime_error("HLLD requires fluid = glmmhd");
  if (cfg_.fluid == APK_FLUID_GLMMHD && cfg_.riemann == APK_RIEMANN_HLLC)
    throw std::runtime_error("HLLC requires fluid = euler");
  for (int f = 0; f < 6; ++f)  // src/bvals/boundary_conditions_apk.hpp:50-52
    if (cfg_.bc[f] == APK_BC_REFLECT && cfg_.fluid == APK_FLUID_GLMMHD)
      throw std::runtime_error("Reflecting boundary conditions for MHD need special treatment.");
  for (int d = 0; d < 3; ++d) {
    if (cfg_.nx[d] % cfg_.mb[d] != 0)
      throw std::runtime_error("Block size is not evenly divisible into the base mesh size.");
    nb_[d] = cfg_.nx[d] / cfg_.mb[d];
    n_[d] = (cfg_.nx[d] == 1) ? 1 : cfg_.mb[d] + 2 * cfg_.ng;
  }
  nblocks_global_ = nb_[0] * nb_[1] * nb_[2];
  multilevel_ = cfg_.nref > 0 || cfg_.adaptive;
  if (cfg_.adaptive) {
    if (cfg_.nranks != 1)
      throw std::runtime_error("apk_host: adaptive remeshing runs on one rank (static refinement "
                               "runs on several)");
    if (ndim_ < 2 || cfg_.numlevel < 1)
      throw std::runtime_error("apk_host: refinement = adaptive needs a 2-D or 3-D mesh and numlevel >= 1");
  }
  if (multilevel_) {
    // mesh_refinement.cpp:60-70
    if (cfg_.ng % 2 != 0)
      throw std::runtime_error("Selected nghost is incompatible with mesh refinement because it is "
                               "not a multiple of 2.");
    if (cfg_.nref > 4)
      throw std::runtime_error("apk_host: at most 4 static refinement regions");
    if (cfg_.nranks > 1)
      for (int f = 0; f < 2 * ndim_; ++f)
        if (cfg_.bc[f] != APK_BC_PERIODIC)
          throw std::runtime_error("apk_host: refined meshes on non-periodic domains run on one rank");
    for (int d = 0; d < ndim_; ++d)
      if (cfg_.mb[d] % 2 != 0 || cfg_.mb[d] / 2 < cfg_.ng)
        throw std::runtime_error("apk_host: meshblock too small for refinement (nx/2 >= nghost)");
  }
  // FOFC decks run fused too: the stage kernel evaluates FirstOrderFluxCorrect's first attempt on
  // its own trial update, and only a stage in which a cell actually fails is redone with the
  // flux-materialising tasks (FusedFofcCheck)
  // multilevel meshes need materialised fluxes (flux correction, hydro_driver.cpp:527-531)
  // diffusion (either integrator) runs on the flux-materialising tasks: unsplit adds to the
  // hyperbolic fluxes (hydro.cpp:1202-1205), RKL2 needs the flux arrays for its sub-steps
  // passive scalars: the fused kernels are built for 1 and 2 scalars (more run un-fused, as do
  // scalars with FOFC, whose acceptance test the fused kernel evaluates on the hydro variables only)
  const bool scalars_ok = cfg_.nscalars == 0 || (cfg_.nscalars <= 2 && !cfg_.fofc);
  fused_ = cfg_.fused && !cfg_.dedner_extended && scalars_ok && !multilevel_ && cfg_.diffint == 0;
  if (cfg_.diffint != 0) {
    if (cfg_.fofc)
      throw std::runtime_error("apk_host: first_order_flux_correct with diffusion is not supported");
    if ((cfg_.conduction == APK_CONDUCTION_ANISOTROPIC || cfg_.resistivity) &&
        cfg_.fluid != APK_FLUID_GLMMHD)
      throw std::runtime_error("Anisotropic thermal conduction / resistivity require fluid = glmmhd");
    if (cfg_.conduction == APK_CONDUCTION_ANISOTROPIC && ndim_ < 2)
      throw std::runtime_error("apk_host: anisotropic conduction needs a 2-D or 3-D mesh");
    diff_.conduction = cfg_.conduction;
    diff_.viscosity = cfg_.viscosity;
    diff_.resistivity = cfg_.resistivity;
    diff_.thermal_diff_coeff = cfg_.thermal_diff_coeff;
    diff_.mom_diff_coeff = cfg_.mom_diff_coeff;
    diff_.ohm_diff_coeff = cfg_.ohm_diff_coeff;
    diff_.conduction_sat_prefac = 5.0 * cfg_.conduction_sat_phi;  // hydro.cpp:604
  }
  BuildMesh();
  BuildNeighbors();
  if (!cfg_.plan_only) {
    cudaStream_t st;
    cuda_check(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
    stream_ = st;
    UploadBlocks();
  } else if (multilevel_) {
    BuildMultilevelPlan();  // descriptors only (ml_desc_): CPU tests of the multi-rank partition
  }
  BuildHaloPlan();
}

HydroSim::~HydroSim() {
  if (cfg_.plan_only) return;
  cudaDeviceSynchronize();
  if (comm_) ncclCommDestroy(reinterpret_cast<ncclComm_t>(comm_));
  apk_tma_maps_destroy(d_tma_[0]);
  apk_tma_maps_destroy(d_tma_[1]);
  d_tma_[0] = d_tma_[1] = nullptr;
  ReleaseMeshArrays();
  SlabFree(d_cons_);
  for (Arena &A : arena_) {
    for (void *p : A.extra) cudaFree(p);
    cudaFree(A.base);
  }
  cudaFree(d_tags_);
  cudaFree(d_maxeps_);
  cudaFree(d_scalars_);
  if (h_scalars_) cudaFreeHost(h_scalars_);
  cudaFree(d_status_);
  apk_halo_plan_destroy(halo_plan_);
  if (!peer_mapped_.empty()) apk_p2p_unmap(peer_mapped_.data(), (int)peer_mapped_.size());
  cudaFree(d_push_[0]);
  cudaFree(d_push_[1]);
  cudaFree(d_barrier_);
  cudaFree(d_stage_);
  cudaFree(d_hist_);
  cudaFree(d_int_items_);
  if (stream_) cudaStreamDestroy(S(stream_));
}

void HydroSim::Synchronize() { cuda_check(cudaStreamSynchronize(S(stream_)), "sync"); }

// Block list in Morton (Z) order like Parthenon's forest (forest.cpp:106-160); contiguous gid
// ranges go to ranks (what Parthenon's load balancer does for equal-cost blocks).
void HydroSim::BuildMesh() {
  nblocks_global_ = nb_[0] * nb_[1] * nb_[2];  // (a remesh comes back here with the old leaf count)
  std::vector<std::pair<std::uint64_t, std::array<int, 3>>> order;
  for (int z = 0; z < nb_[2]; ++z)
    for (int y = 0; y < nb_[1]; ++y)
      for (int x = 0; x < nb_[0]; ++x) order.push_back({Morton(x, y, z), {x, y, z}});
  std::sort(order.begin(), order.end());
  gid_of_loc_.assign(nblocks_global_, -1);
  loc_of_gid_.resize(nblocks_global_);
  for (int g = 0; g < nblocks_global_; ++g) {
    loc_of_gid_[g] = order[g].second;
    const auto &l = order[g].second;
    gid_of_loc_[l[0] + nb_[0] * (l[1] + nb_[1] * l[2])] = g;
  }
  if (multilevel_) {
    if (leaves_.empty()) leaves_ = BuildLeaves();  // (a remesh installs the new leaf list itself)
    nblocks_global_ = (int)leaves_.size();
    for (int d = 0; d < 3; ++d) nc_[d] = (d < ndim_) ? cfg_.mb[d] / 2 + 2 * cfg_.ng : 1;
  }
  // equal split (remainder to the first ranks)
  gid_lo_ = GidLo(cfg_.rank);
  gid_hi_ = GidLo(cfg_.rank + 1);
  if (gid_hi_ <= gid_lo_) throw std::runtime_error("apk_host: rank without meshblocks");
  blocks_.clear();
  for (int g = gid_lo_; g < gid_hi_; ++g) {
    if (multilevel_) blocks_.push_back(MakeBlockInfo(g, leaves_[g].level, leaves_[g].lx));
    else blocks_.push_back(MakeBlockInfo(g, 0, loc_of_gid_[g].data()));
  }
  if (multilevel_) {
    // hydro.cpp:65-95 + the MPI_Allreduce(MIN) of hydro.cpp:127: every rank knows every leaf
    Real m = std::numeric_limits<Real>::max();
    for (int g = 0; g < nblocks_global_; ++g) {
      const BlockInfo B = MakeBlockInfo(g, leaves_[g].level, leaves_[g].lx);
      m = std::fmin(m, B.dx[0]);
      if (n_[1] > 1) m = std::fmin(m, B.dx[1]);
      if (n_[2] > 1) m = std::fmin(m, B.dx[2]);
    }
    global_mindx_ = m;
  }

  desc_.nblocks = (int)blocks_.size();
  desc_.nvar = nvar_;
  desc_.nscalars = cfg_.nscalars;
  for (int d = 0; d < 3; ++d) desc_.nx[d] = cfg_.mb[d];
  desc_.ng = cfg_.ng;
  desc_.ndim = ndim_;
  desc_.fluid = cfg_.fluid;
  desc_.recon = cfg_.recon;
  desc_.riemann = cfg_.riemann;
  desc_.fp_mode = cfg_.fp_mode;
  desc_.gamma = cfg_.gamma;
  desc_.dfloor = cfg_.dfloor;
  desc_.pfloor = cfg_.pfloor;
  desc_.efloor = cfg_.efloor;
  desc_.vceil = cfg_.vceil;
  desc_.eceil = cfg_.eceil;
}

int HydroSim::GidLo(int rank) const {
  const int base = nblocks_global_ / cfg_.nranks, rem = nblocks_global_ % cfg_.nranks;
  return rank * base + std::min(rank, rem);
}

int HydroSim::RankOfGid(int gid) const {
  const int base = nblocks_global_ / cfg_.nranks, rem = nblocks_global_ % cfg_.nranks;
  const int cut = rem * (base + 1);
  return (gid < cut) ? gid / (base + 1) : rem + (gid - cut) / base;
}

// forest.cpp:66-141 + tree.cpp:296-318: tree decomposition then block-in-tree bounds
BlockInfo HydroSim::MakeBlockInfo(int gid, int level, const int *loc) const {
  bool any_sym = false;
  int mcp2 = 1 << 30;
  for (int d = 0; d < 3; ++d) {
    if (cfg_.nx[d] == 1) { any_sym = true; continue; }
    mcp2 = std::min(mcp2, MaximumPowerOf2Divisor(nb_[d]));
  }
  if (any_sym) mcp2 = 1;
  BlockInfo B{};
  B.gid = gid;
  B.level = level;
  for (int d = 0; d < 3; ++d) {
    B.loc[d] = loc[d];
    if (cfg_.nx[d] == 1) {
      B.bxmin[d] = SymmetrizedLogicalToActualPosition(IndexToSymmetrizedCoordinate(0, 0, 1),
                                                      cfg_.xmin[d], cfg_.xmax[d]);
      B.bxmax[d] = SymmetrizedLogicalToActualPosition(IndexToSymmetrizedCoordinate(0, 2, 1),
                                                      cfg_.xmin[d], cfg_.xmax[d]);
    } else {
      // a block of refinement level l: its tree is cut mcp2 * 2^l times (tree.cpp:296-318)
      const int per_tree = mcp2 << B.level;
      const int ntree = nb_[d] / mcp2, it = B.loc[d] / per_tree, lx = B.loc[d] % per_tree;
      const double tmin = SymmetrizedLogicalToActualPosition(
          IndexToSymmetrizedCoordinate(it, 0, ntree), cfg_.xmin[d], cfg_.xmax[d]);
      const double tmax = SymmetrizedLogicalToActualPosition(
          IndexToSymmetrizedCoordinate(it, 2, ntree), cfg_.xmin[d], cfg_.xmax[d]);
      B.bxmin[d] = SymmetrizedLogicalToActualPosition(IndexToSymmetrizedCoordinate(lx, 0, per_tree),
                                                      tmin, tmax);
      B.bxmax[d] = SymmetrizedLogicalToActualPosition(IndexToSymmetrizedCoordinate(lx, 2, per_tree),
                                                      tmin, tmax);
    }
    // uniform_coordinates.hpp:32-36
    B.dx[d] = (B.bxmax[d] - B.bxmin[d]) / cfg_.mb[d];
    const int istart = (cfg_.nx[d] == 1) ? 0 : cfg_.ng;
    B.xmin[d] = B.bxmin[d] - istart * B.dx[d];
  }
  return B;
}

// ---- static mesh refinement -----------------------------------------------------------------
namespace {
struct LeafKey {
  int level, x, y, z;
  bool operator<(const LeafKey &o) const {
    return std::tie(level, x, y, z) < std::tie(o.level, o.x, o.y, o.z);
  }
};
}  // namespace

// The leaf set Parthenon's forest ends up with for <parthenon/static_refinement> regions
// (mesh.cpp:1076-1165: index range of the requested blocks per region; forest.AddMeshBlock refines
// down to them and keeps neighbouring leaves -- faces, edges, corners, periodic wrap -- within one
// level of each other): the minimal properly nested tree containing the requested blocks.
std::vector<HydroSim::Leaf> HydroSim::BuildLeaves() const {
  int rlo[4][3] = {}, rhi[4][3] = {};
  for (int r = 0; r < cfg_.nref; ++r)
    for (int d = 0; d < ndim_; ++d) {
      const int lxtot = nb_[d] * (1 << cfg_.ref_level[r]);
      int lxmin, lxmax;
      for (lxmin = 0; lxmin < lxtot; lxmin++) {
        const double x = SymmetrizedLogicalToActualPosition(
            IndexToSymmetrizedCoordinate(lxmin + 1, 0, lxtot), cfg_.xmin[d], cfg_.xmax[d]);
        if (x > cfg_.ref_xmin[r][d]) break;
      }
      for (lxmax = lxmin; lxmax < lxtot; lxmax++) {
        const double x = SymmetrizedLogicalToActualPosition(
            IndexToSymmetrizedCoordinate(lxmax + 1, 0, lxtot), cfg_.xmin[d], cfg_.xmax[d]);
        if (x >= cfg_.ref_xmax[r][d]) break;
      }
      if (lxmin % 2 == 1) lxmin--;
      if (lxmax % 2 == 0) lxmax++;
      rlo[r][d] = lxmin;
      rhi[r][d] = lxmax;
    }
  std::map<LeafKey, int> leaves;  // value unused
  for (int z = 0; z < nb_[2]; ++z)
    for (int y = 0; y < nb_[1]; ++y)
      for (int x = 0; x < nb_[0]; ++x) leaves[{0, x, y, z}] = 0;
  auto cover = [&](int level, const int *lx, LeafKey *out) {
    for (int l = level; l >= 0; --l) {
      LeafKey k{l, lx[0] >> (level - l), lx[1] >> (level - l), lx[2] >> (level - l)};
      if (leaves.count(k)) {
        *out = k;
        return true;
      }
    }
    return false;
  };
  for (;;) {
    std::vector<LeafKey> marked;
    for (const auto &kv : leaves) {
      const LeafKey &q = kv.first;
      const int lx[3] = {q.x, q.y, q.z};
      bool refine = false;
      for (int r = 0; r < cfg_.nref && !refine; ++r) {
        const int dl = cfg_.ref_level[r] - q.level;
        if (dl <= 0) continue;
        bool hit = true;
        for (int d = 0; d < ndim_; ++d) {
          const int lo = lx[d] << dl, hi = ((lx[d] + 1) << dl) - 1;
          if (hi < rlo[r][d] || lo > rhi[r][d]) hit = false;
        }
        refine = hit;
      }
      if (refine) marked.push_back(q);
      if (q.level >= 2) {  // proper nesting around this leaf
        for (int o3 = (ndim_ > 2 ? -1 : 0); o3 <= (ndim_ > 2 ? 1 : 0); ++o3)
          for (int o2 = (ndim_ > 1 ? -1 : 0); o2 <= (ndim_ > 1 ? 1 : 0); ++o2)
            for (int o1 = -1; o1 <= 1; ++o1) {
              const int off[3] = {o1, o2, o3};
              int nl[3];
              bool inside = true;
              for (int d = 0; d < 3; ++d) {
                const int tot = nb_[d] << q.level;
                nl[d] = lx[d] + off[d];
                if (nl[d] < 0 || nl[d] >= tot) {
                  if (d < ndim_ && cfg_.bc[2 * d] != APK_BC_PERIODIC) inside = false;
                  nl[d] = (nl[d] % tot + tot) % tot;
                }
              }
              LeafKey c{};
              if (inside && cover(q.level, nl, &c) && c.level < q.level - 1) marked.push_back(c);
            }
      }
    }
    if (marked.empty()) break;
    for (const LeafKey &q : marked) {
      if (!leaves.erase(q)) continue;  // marked twice
      for (int c = 0; c < (1 << ndim_); ++c)
        leaves[{q.level + 1, 2 * q.x + (c & 1), ndim_ > 1 ? 2 * q.y + ((c >> 1) & 1) : 0,
                ndim_ > 2 ? 2 * q.z + ((c >> 2) & 1) : 0}] = 0;
    }
  }
  std::vector<Leaf> out;
  for (const auto &kv : leaves) out.push_back({kv.first.level, {kv.first.x, kv.first.y, kv.first.z}});
  return out;
}

void HydroSim::BuildMultilevelNeighbors() {
  // every rank builds the topology of ALL leaves (ids = global block ids): items of the exchange
  // plan exist wherever either side of a relation is local
  // (hashed: an adaptive run rebuilds these tables after every remesh)
  struct PackedIndex {
    std::unordered_map<std::uint64_t, int> m;
    static std::uint64_t key(const LeafKey &k) {
      return ((std::uint64_t)k.level << 60) | ((std::uint64_t)k.x << 40) | ((std::uint64_t)k.y << 20) |
             (std::uint64_t)k.z;
    }
    int &operator[](const LeafKey &k) { return m[key(k)]; }
    std::unordered_map<std::uint64_t, int>::const_iterator find(const LeafKey &k) const {
      return m.find(key(k));
    }
    std::unordered_map<std::uint64_t, int>::const_iterator end() const { return m.end(); }
  } index;
  index.m.reserve(leaves_.size() * 2);
  for (size_t g = 0; g < leaves_.size(); ++g)
    index[{leaves_[g].level, leaves_[g].lx[0], leaves_[g].lx[1], leaves_[g].lx[2]}] = (int)g;
  gnbr_.assign(leaves_.size(), NbrInfo{});
  for (size_t b = 0; b < leaves_.size(); ++b) {
    struct { int level; int loc[3]; } B = {leaves_[b].level,
                                           {leaves_[b].lx[0], leaves_[b].lx[1], leaves_[b].lx[2]}};
    NbrInfo &N = gnbr_[b];
    N.has_coarser = false;
    for (int idx = 0; idx < 27; ++idx) {
      N.type[idx] = 0;
      N.nchild[idx] = 0;
      const int off[3] = {idx % 3 - 1, (idx / 3) % 3 - 1, idx / 9 - 1};
      if (idx == 13) continue;
      bool valid = true;
      int nl[3] = {0, 0, 0};
      for (int d = 0; d < 3; ++d) {
        if (d >= ndim_) {
          if (off[d] != 0) valid = false;
          continue;
        }
        const int tot = nb_[d] << B.level;
        nl[d] = B.loc[d] + off[d];
        if (nl[d] < 0 || nl[d] >= tot) {
          if (cfg_.bc[2 * d] != APK_BC_PERIODIC) valid = false;  // physical boundary: no neighbour
          nl[d] = (nl[d] % tot + tot) % tot;
        }
      }
      if (!valid) continue;
      auto same = index.find({B.level, nl[0], nl[1], nl[2]});
      if (same != index.end()) {
        N.type[idx] = 1;
        N.idx[idx][0] = same->second;
        N.nchild[idx] = 1;
        continue;
      }
      auto coarser = (B.level > 0) ? index.find({B.level - 1, nl[0] >> 1, nl[1] >> 1, nl[2] >> 1})
                                   : index.end();
      if (coarser != index.end()) {
        N.type[idx] = 2;
        N.idx[idx][0] = coarser->second;
        N.nchild[idx] = 1;
        N.has_coarser = true;
        continue;
      }
      N.type[idx] = 3;  // finer: the children that touch this block
      for (int c = 0; c < 8; ++c) {
        const int cb[3] = {c & 1, (c >> 1) & 1, (c >> 2) & 1};
        bool ok = true;
        int cl[3] = {0, 0, 0};
        for (int d = 0; d < 3; ++d) {
          if (d >= ndim_) {
            if (cb[d]) ok = false;
            continue;
          }
          if (off[d] > 0 && cb[d] != 0) ok = false;
          if (off[d] < 0 && cb[d] != 1) ok = false;
          cl[d] = 2 * nl[d] + cb[d];
        }
        if (!ok) continue;
        auto f = index.find({B.level + 1, cl[0], cl[1], cl[2]});
        if (f == index.end()) throw std::runtime_error("apk_host: mesh is not properly nested");
        N.idx[idx][N.nchild[idx]++] = f->second;
      }
    }
  }
  nbrinfo_.assign(blocks_.size(), NbrInfo{});
  for (size_t b = 0; b < blocks_.size(); ++b) nbrinfo_[b] = gnbr_[gid_lo_ + b];
}

void HydroSim::BuildNeighbors() {
  h_blocks_.assign(blocks_.size(), apk_block_t{});
  bc_face_mask_ = 0;
  if (multilevel_) {
    BuildMultilevelNeighbors();
    for (size_t b = 0; b < blocks_.size(); ++b) {
      apk_block_t &hb = h_blocks_[b];
      for (int d = 0; d < 3; ++d) {
        hb.dx[d] = blocks_[b].dx[d];
        hb.xmin[d] = blocks_[b].xmin[d];
      }
      // tree.cpp:320-333 GetBlockBCs at the block's own level
      for (int f = 0; f < 6; ++f) {
        const int d = f / 2;
        const int last = (nb_[d] << blocks_[b].level) - 1;
        const bool on_edge = (f % 2 == 0) ? (blocks_[b].loc[d] == 0) : (blocks_[b].loc[d] == last);
        hb.bc[f] = APK_BC_BLOCK;
        if (d < ndim_ && on_edge && cfg_.bc[f] != APK_BC_PERIODIC) {
          hb.bc[f] = cfg_.bc[f];
          bc_face_mask_ |= 1u << f;
        }
      }
      // the same-level table drives apk_copy_ghosts_local; other levels go through the plan
      for (int idx = 0; idx < 27; ++idx)
        hb.nbr[idx] = (nbrinfo_[b].type[idx] == 1 && IsLocal(nbrinfo_[b].idx[idx][0]))
                          ? nbrinfo_[b].idx[idx][0] - gid_lo_
                          : APK_NBR_NONE;  // (same-level neighbours on other ranks: ml_same_remote_)
    }
    return;
  }
  for (size_t b = 0; b < blocks_.size(); ++b) {
    apk_block_t &hb = h_blocks_[b];
    const BlockInfo &B = blocks_[b];
    for (int d = 0; d < 3; ++d) {
      hb.dx[d] = B.dx[d];
      hb.xmin[d] = B.xmin[d];
    }
    // tree.cpp:320-333 GetBlockBCs
    for (int f = 0; f < 6; ++f) {
      const int d = f / 2;
      const bool on_edge = (f % 2 == 0) ? (B.loc[d] == 0) : (B.loc[d] == nb_[d] - 1);
      hb.bc[f] = APK_BC_BLOCK;
      if (d < ndim_ && on_edge && cfg_.bc[f] != APK_BC_PERIODIC) {
        hb.bc[f] = cfg_.bc[f];
        bc_face_mask_ |= 1u << f;
      }
    }
    for (int o3 = -1; o3 <= 1; ++o3)
      for (int o2 = -1; o2 <= 1; ++o2)
        for (int o1 = -1; o1 <= 1; ++o1) {
          const int idx = (o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1);
          hb.nbr[idx] = APK_NBR_NONE;
          const int off[3] = {o1, o2, o3};
          if (!o1 && !o2 && !o3) continue;
          bool valid = true;
          int nl[3];
          for (int d = 0; d < 3; ++d) {
            if (off[d] != 0 && cfg_.nx[d] == 1) valid = false;
            nl[d] = B.loc[d] + off[d];
            if (nl[d] < 0 || nl[d] >= nb_[d]) {
              const int face = 2 * d + (nl[d] < 0 ? 0 : 1);
              if (cfg_.bc[face] != APK_BC_PERIODIC) valid = false;
              nl[d] = (nl[d] + nb_[d]) % nb_[d];
            }
          }
          if (!valid) continue;
          const int ngid = gid_of_loc_[nl[0] + nb_[0] * (nl[1] + nb_[1] * nl[2])];
          if (ngid >= gid_lo_ && ngid < gid_hi_) hb.nbr[idx] = ngid - gid_lo_;
          else hb.nbr[idx] = APK_NBR_REMOTE(0);  // slot resolved by the halo plan
        }
  }
}

// Per-mesh device arrays (see Arena in hydro_sim.hpp).
void HydroSim::ArenaBegin(double growth) {
  const size_t last = arena_[arena_cur_].want;
  arena_cur_ ^= 1;
  Arena &A = arena_[arena_cur_];
  for (void *p : A.extra) cudaFree(p);
  A.extra.clear();
  const size_t need = (size_t)((double)std::max(last, A.want) * std::max(growth, 1.0) * 1.05) + (8u << 20);
  if (A.cap < need) {
    cudaFree(A.base);
    A.base = nullptr;
    A.cap = 2 * need;  // (a growing mesh regrows an arena rarely: cudaMalloc of GBs costs tens of ms)
    cuda_check(cudaMalloc(reinterpret_cast<void **>(&A.base), A.cap), "cudaMalloc mesh arena");
  }
  A.used = A.want = 0;
}

void *HydroSim::SlabAlloc(size_t bytes, const char *what) {
  void *p = nullptr;
  if (!cfg_.adaptive) {
    cuda_check(cudaMalloc(&p, bytes), what);
    return p;
  }
  Arena &A = arena_[arena_cur_];
  const size_t off = (A.used + 255) & ~(size_t)255;
  A.want += bytes + 256;
  if (off + bytes <= A.cap) {
    A.used = off + bytes;
    return A.base + off;
  }
  cuda_check(cudaMalloc(&p, bytes), what);
  A.extra.push_back(p);
  return p;
}

void HydroSim::SlabFree(void *p) {
  if (!p || cfg_.adaptive) return;  // (arena memory is reclaimed when its arena has its next turn)
  cudaFree(p);
}

// everything UploadBlocks allocates for one mesh, except the cons slab (a remesh still reads it)
void HydroSim::ReleaseMeshArrays() {
  auto slab = [&](double *&p) { SlabFree(p); p = nullptr; };
  slab(d_prim_[0]);
  slab(d_prim_[1]);
  slab(d_u1_);
  for (int d = 0; d < 3; ++d) slab(d_flux_[d]);
  slab(d_my0_);
  slab(d_yjm2_);
  slab(d_coarse_);
  auto fr = [](auto *&p) { cudaFree(p); p = nullptr; };
  auto pooled = [&](auto *&p) { SlabFree(p); p = nullptr; };
  fr(d_my0_ptrs_);
  fr(d_yjm2_ptrs_);
  pooled(d_blocks_[0]);
  pooled(d_blocks_[1]);
  pooled(d_prim_out_[0]);
  pooled(d_prim_out_[1]);
  pooled(d_cblocks_);
  fr(d_undo_);
  fr(d_fofc_work_);
  fr(d_hist_);
  fr(d_stage_);
  fr(d_int_items_);
  for (MlList *l : {&ml_restrict_child_, &ml_child_to_ghost_, &ml_restrict_own_, &ml_coarse_from_nbr_,
                    &ml_prolong_, &ml_fc_restrict_[0], &ml_fc_restrict_[1], &ml_fc_restrict_[2],
                    &ml_fc_copy_[0], &ml_fc_copy_[1], &ml_fc_copy_[2], &ml_same_remote_}) {
    SlabFree(l->d);
    l->d = nullptr;
    l->n = 0;
  }
}

void HydroSim::UploadBlocks() {
  const size_t per = BlockFieldElems();
  const size_t nbl = blocks_.size();
  const size_t bytes = per * nbl * sizeof(double);
  // (a remesh re-enters with recycled slabs: prim, u1, the fluxes and the coarse buffers are
  // rewritten in full before they are read -- FillDerived over the entire index range, InitU1,
  // CalculateFluxes, the exchange -- so only cons is cleared again)
  const bool first = (d_scalars_ == nullptr);
  auto slab = [&](const char *what, bool clear = true) {
    double *p = static_cast<double *>(SlabAlloc(bytes, what));
    if (clear || first) cuda_check(cudaMemsetAsync(p, 0, bytes, S(stream_)), "memset");
    return p;
  };
  d_cons_ = slab("cudaMalloc cons");
  d_prim_[0] = slab("cudaMalloc prim", false);
  d_u1_ = slab("cudaMalloc u1", false);
  if (fused_) d_prim_[1] = slab("cudaMalloc prim scratch", false);
  if (!fused_ || cfg_.fofc) {
    for (int d = 0; d < ndim_; ++d) d_flux_[d] = slab("cudaMalloc flux", false);
  }
  if (cfg_.diffint == 2) {  // RKL2 registers (hydro_driver.cpp:211-212)
    d_my0_ = slab("cudaMalloc MY0");
    d_yjm2_ = slab("cudaMalloc Yjm2");
    std::vector<double *> pm(nbl), py(nbl);
    for (size_t b = 0; b < nbl; ++b) {
      pm[b] = d_my0_ + b * per;
      py[b] = d_yjm2_ + b * per;
    }
    cuda_check(cudaMalloc(&d_my0_ptrs_, nbl * sizeof(double *)), "cudaMalloc MY0 table");
    cuda_check(cudaMalloc(&d_yjm2_ptrs_, nbl * sizeof(double *)), "cudaMalloc Yjm2 table");
    cuda_check(cudaMemcpy(d_my0_ptrs_, pm.data(), nbl * sizeof(double *), cudaMemcpyHostToDevice), "upload");
    cuda_check(cudaMemcpy(d_yjm2_ptrs_, py.data(), nbl * sizeof(double *), cudaMemcpyHostToDevice), "upload");
  }
  if (fused_ && cfg_.fofc && cfg_.integrator != (int)Integrator::vl2 && integrator_.nstages > 1) {
    // stages after the first overwrite their input cons: keep a copy to roll back to
    cuda_check(cudaMalloc(&d_undo_, bytes), "cudaMalloc fofc undo");
  }
  for (int which = 0; which < (fused_ ? 2 : 1); ++which) {
    std::vector<apk_block_t> hb = h_blocks_;
    std::vector<double *> pout(nbl);
    for (size_t b = 0; b < nbl; ++b) {
      hb[b].cons = d_cons_ + b * per;
      hb[b].prim = d_prim_[which] + b * per;
      hb[b].cons_u1 = d_u1_ + b * per;
      for (int d = 0; d < 3; ++d) hb[b].flux[d] = d_flux_[d] ? d_flux_[d] + b * per : nullptr;
      pout[b] = fused_ ? d_prim_[1 - which] + b * per : nullptr;
    }
    d_blocks_[which] = static_cast<apk_block_t *>(SlabAlloc(nbl * sizeof(apk_block_t), "cudaMalloc blocks"));
    cuda_check(cudaMemcpy(d_blocks_[which], hb.data(), nbl * sizeof(apk_block_t),
                          cudaMemcpyHostToDevice), "upload blocks");
    // A/B: TMA-staged prim tiles for the PPM stage (APK_B200_FUSED_V=3; GLM-MHD PPM+HLLD 3-D, FMA mode)
    const char *ver = std::getenv("APK_B200_FUSED_V");
    if (fused_ && ver && ver[0] == '3' && ndim_ == 3 && cfg_.fluid == APK_FLUID_GLMMHD &&
        cfg_.recon == APK_RECON_PPM && cfg_.riemann == APK_RIEMANN_HLLD && cfg_.fp_mode == APK_FP_FAST)
      abi_check(apk_tma_maps_create(&desc_, hb.data(), &d_tma_[which]), "apk_tma_maps_create");
    d_prim_out_[which] = static_cast<double **>(SlabAlloc(nbl * sizeof(double *), "cudaMalloc prim_out"));
    cuda_check(cudaMemcpy(d_prim_out_[which], pout.data(), nbl * sizeof(double *),
                          cudaMemcpyHostToDevice), "upload prim_out");
  }
  if (multilevel_) {
    const size_t cbytes = CoarseElems() * nbl * sizeof(double);
    d_coarse_ = static_cast<double *>(SlabAlloc(cbytes, "cudaMalloc coarse"));
    if (first) cuda_check(cudaMemsetAsync(d_coarse_, 0, cbytes, S(stream_)), "memset");
    if (cfg_.nranks == 1) BuildMultilevelPlan();  // several ranks: after the peer mapping (InitComm)
    if (bc_face_mask_) {
      // physical BCs on the coarse buffers (ApplyBoundaryConditionsOnCoarseOrFine(coarse = true)):
      // the same kernel on a second block table whose arrays are the coarse buffers
      desc_coarse_ = desc_;
      for (int d = 0; d < ndim_; ++d) desc_coarse_.nx[d] = cfg_.mb[d] / 2;
      std::vector<apk_block_t> cb = h_blocks_;
      for (size_t b = 0; b < nbl; ++b) {
        cb[b].cons = d_coarse_ + b * CoarseElems();
        cb[b].prim = cb[b].cons_u1 = nullptr;
        for (int d = 0; d < 3; ++d) {
          cb[b].flux[d] = nullptr;
          cb[b].dx[d] = (d < ndim_) ? 2.0 * h_blocks_[b].dx[d] : h_blocks_[b].dx[d];
        }
        // (every array pointer of a descriptor must be set: the coarse buffer stands in)
        cb[b].prim = cb[b].cons_u1 = cb[b].cons;
      }
      d_cblocks_ = static_cast<apk_block_t *>(SlabAlloc(nbl * sizeof(apk_block_t), "cudaMalloc coarse blocks"));
      cuda_check(cudaMemcpy(d_cblocks_, cb.data(), nbl * sizeof(apk_block_t), cudaMemcpyHostToDevice),
                 "upload coarse blocks");
    }
  }
  // the slabs are zeroed on the simulation stream; everything that follows (blocking copies of the
  // problem generator, kernels on any stream) must see them cleared
  cuda_check(cudaStreamSynchronize(S(stream_)), "sync after memset");
  if (!d_scalars_) {
    cuda_check(cudaMalloc(&d_scalars_, 64 * sizeof(double)), "cudaMalloc scalars");
    cuda_check(cudaMemset(d_scalars_, 0, 64 * sizeof(double)), "memset scalars");
    const double big = DBL_MAX;
    cuda_check(cudaMemcpy(d_scalars_ + 8, &big, sizeof(double), cudaMemcpyHostToDevice), "scalars");
    cuda_check(cudaHostAlloc(reinterpret_cast<void **>(&h_scalars_), 8 * sizeof(double),
                             cudaHostAllocDefault), "pinned scalars");
    cuda_check(cudaMalloc(&d_status_, sizeof(int)), "cudaMalloc status");
    cuda_check(cudaMemset(d_status_, 0, sizeof(int)), "memset");
  }
  if (cfg_.fofc) {
    cuda_check(cudaMalloc(&d_fofc_work_, apk_fofc_workspace_bytes(&desc_)), "cudaMalloc fofc");
  }
}

// Remote halo plan: every (local block, offset) whose neighbour lives on another rank
// contributes one slab to that peer's send buffer and one to its receive buffer.  Both sides
// order slabs by (receiver gid, receiver-side offset index) so offsets agree without talking.
void HydroSim::BuildHaloPlan() {
  if (cfg_.nranks == 1 || multilevel_) return;  // refined meshes: peer pulls, BuildMultilevelPlan
  const int ng = cfg_.ng;
  const int s[3] = {cfg_.ng, ndim_ >= 2 ? cfg_.ng : 0, ndim_ >= 3 ? cfg_.ng : 0};
  struct Entry {
    long long key;
    apk_halo_item_t item;
    long long cells;
  };
  std::map<int, std::vector<Entry>> sends, recvs;
  const int base = nblocks_global_ / cfg_.nranks, rem = nblocks_global_ % cfg_.nranks;
  auto rank_of = [&](int gid) {
    const int cut = rem * (base + 1);
    return (gid < cut) ? gid / (base + 1) : rem + (gid - cut) / base;
  };
  for (size_t b = 0; b < blocks_.size(); ++b) {
    const BlockInfo &B = blocks_[b];
    for (int o3 = -1; o3 <= 1; ++o3)
      for (int o2 = -1; o2 <= 1; ++o2)
        for (int o1 = -1; o1 <= 1; ++o1) {
          const int idx = (o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1);
          if (h_blocks_[b].nbr[idx] > APK_NBR_NONE || h_blocks_[b].nbr[idx] == APK_NBR_NONE)
            continue;
          const int off[3] = {o1, o2, o3};
          int nl[3];
          for (int d = 0; d < 3; ++d) nl[d] = ((B.loc[d] + off[d]) % nb_[d] + nb_[d]) % nb_[d];
          const int ngid = gid_of_loc_[nl[0] + nb_[0] * (nl[1] + nb_[1] * nl[2])];
          const int peer = rank_of(ngid);
          // receive: my ghost slab at offset `off` (bnd_info.cpp:136-230)
          Entry r{};
          r.item.block = (int)b;
          r.cells = 1;
          // send: my interior slab adjacent to the neighbour
          Entry sd{};
          sd.item.block = (int)b;
          sd.cells = 1;
          for (int d = 0; d < 3; ++d) {
            const int e = s[d] + cfg_.mb[d] - 1;
            if (off[d] == 0) {
              r.item.lo[d] = sd.item.lo[d] = s[d];
              r.item.ext[d] = sd.item.ext[d] = (cfg_.nx[d] == 1) ? 1 : cfg_.mb[d];
            } else if (off[d] > 0) {
              r.item.lo[d] = e + 1;
              sd.item.lo[d] = e - ng + 1;
              r.item.ext[d] = sd.item.ext[d] = ng;
            } else {
              r.item.lo[d] = s[d] - ng;
              sd.item.lo[d] = s[d];
              r.item.ext[d] = sd.item.ext[d] = ng;
            }
            r.cells *= r.item.ext[d];
            sd.cells *= sd.item.ext[d];
          }
          r.key = (long long)B.gid * 27 + idx;
          const int ridx = (-o3 + 1) * 9 + (-o2 + 1) * 3 + (-o1 + 1);
          sd.key = (long long)ngid * 27 + ridx;
          recvs[peer].push_back(r);
          sends[peer].push_back(sd);
        }
  }
  send_items_.clear();
  recv_items_.clear();
  peers_.clear();
  long long soff = 0, roff = 0;
  // factor 2 leaves room for prim when the fused path exchanges cons+prim
  const long long per_cell = (long long)nvar_ * 2;
  for (auto &kv : sends) {
    Peer p{};
    p.rank = kv.first;
    auto &sv = kv.second;
    auto &rv = recvs[kv.first];
    auto by_key = [](const Entry &a, const Entry &b) { return a.key < b.key; };
    std::sort(sv.begin(), sv.end(), by_key);
    std::sort(rv.begin(), rv.end(), by_key);
    p.send_off = soff;
    for (auto &e : sv) {
      e.item.offset = soff;
      soff += e.cells * per_cell;
      send_items_.push_back(e.item);
    }
    p.send_cnt = soff - p.send_off;
    p.recv_off = roff;
    for (auto &e : rv) {
      e.item.offset = roff;
      roff += e.cells * per_cell;
      recv_items_.push_back(e.item);
    }
    p.recv_cnt = roff - p.recv_off;
    peers_.push_back(p);
  }
  send_total_ = soff;
  recv_total_ = roff;
}

int HydroSim::CommUniqueId(void *out, int nbytes) {
  if (nbytes < (int)sizeof(ncclUniqueId)) return -1;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return -2;
  std::memcpy(out, &id, sizeof(id));
  return (int)sizeof(id);
}

int HydroSim::InitComm(const void *uid, int nbytes) {
  if (cfg_.nranks == 1) return 0;
  if (nbytes < (int)sizeof(ncclUniqueId)) return -1;
  ncclUniqueId id;
  std::memcpy(&id, uid, sizeof(id));
  ncclComm_t c;
  if (ncclCommInitRank(&c, cfg_.nranks, id, cfg_.rank) != ncclSuccess) return -2;
  comm_ = reinterpret_cast<ncclComm *>(c);
  if (multilevel_) {
    // refined mesh over several ranks: every cross-rank relation of the multilevel exchange
    // (same level, fine -> coarse ghosts, coarse -> coarse buffers, flux correction) is a box copy
    // that PULLS from the owner's memory, mapped with CUDA IPC; stage barriers order the phases
    cuda_check(cudaMalloc(&d_barrier_, 2 * sizeof(int)), "barrier word");
    cuda_check(cudaMemset(d_barrier_, 0, 2 * sizeof(int)), "barrier word");
    std::vector<int> need(cfg_.nranks, 1);
    need[cfg_.rank] = 0;
    void *bufs[2] = {d_cons_, d_coarse_};
    ml_peer_.assign((size_t)cfg_.nranks * 2, nullptr);
    int all_ok = 0;
    abi_check(apk_p2p_map_peers(comm_, cfg_.rank, cfg_.nranks, bufs, 2, need.data(), ml_peer_.data(),
                                &all_ok, stream_), "apk_p2p_map_peers");
    if (!all_ok)
      throw std::runtime_error("apk_host: refined meshes on several ranks need CUDA peer access "
                               "(IPC) between the GPUs");
    BuildMultilevelPlan();
    return 0;
  }
  // the exchange itself lives behind the C-ABI (apk_halo_*): hand it the slab lists
  std::vector<apk_halo_peer_t> hp;
  for (const Peer &p : peers_) hp.push_back({p.rank, p.send_off, p.send_cnt, p.recv_off, p.recv_cnt});
  abi_check(apk_halo_plan_create(send_items_.data(), (int)send_items_.size(), recv_items_.data(),
                                 (int)recv_items_.size(), hp.data(), (int)hp.size(), send_total_,
                                 recv_total_, comm_, &halo_plan_), "apk_halo_plan_create");
  SetupPeerPush();
  return 0;
}

bool HydroSim::PushTarget(int b, int idx, int *rank, int *local_index) const {
  const int nb = h_blocks_[b].nbr[idx];
  if (nb == APK_NBR_NONE || idx == 13) return false;
  if (nb >= 0) {
    *rank = cfg_.rank;
    *local_index = nb;
    return true;
  }
  const int nr = cfg_.nranks;
  const int base = nblocks_global_ / nr, rem = nblocks_global_ % nr;
  const int off[3] = {idx % 3 - 1, (idx / 3) % 3 - 1, idx / 9 - 1};
  int nl[3];
  for (int d = 0; d < 3; ++d) nl[d] = ((blocks_[b].loc[d] + off[d]) % nb_[d] + nb_[d]) % nb_[d];
  const int ngid = gid_of_loc_[nl[0] + nb_[0] * (nl[1] + nb_[1] * nl[2])];
  const int cut = rem * (base + 1);
  const int r = (ngid < cut) ? ngid / (base + 1) : rem + (ngid - cut) / base;
  *rank = r;
  *local_index = ngid - (r * base + std::min(r, rem));
  return true;
}

// Map the field buffers of every neighbour rank (CUDA IPC over NVLink/NVSwitch peer access) and
// build, per prim ping-pong parity, the [nblocks][27] table of ghost-push destinations used by
// apk_stage_fused_p2p.  Any failure (no peer access, ranks sharing a process, APK_B200_P2P=0) on
// any rank keeps every rank on the packed NCCL send/recv path.
void HydroSim::SetupPeerPush() {
  p2p_ = false;
  if (!fused_ || cfg_.plan_only || cfg_.nranks == 1) return;
  const int nr = cfg_.nranks;
  const char *env = std::getenv("APK_B200_P2P");
  cuda_check(cudaMalloc(&d_barrier_, 2 * sizeof(int)), "barrier word");
  cuda_check(cudaMemset(d_barrier_, 0, 2 * sizeof(int)), "barrier word");
  // ranks whose cons / prim buffers this rank stores into
  std::vector<int> need(nr, 0);
  for (size_t b = 0; b < blocks_.size(); ++b)
    for (int idx = 0; idx < 27; ++idx) {
      const int nb = h_blocks_[b].nbr[idx];
      if (nb >= 0 || nb == APK_NBR_NONE) continue;
      int r = 0, l = 0;
      if (PushTarget((int)b, idx, &r, &l)) need[r] = 1;
    }
  void *bufs[3] = {d_cons_, d_prim_[0], d_prim_[1]};
  peer_mapped_.assign((size_t)nr * 3, nullptr);
  int all_ok = 0;
  if (env && env[0] == '0') {
    // every rank reads the same environment: nobody maps, everybody keeps the packed path
    all_ok = 0;
  } else {
    abi_check(apk_p2p_map_peers(comm_, cfg_.rank, nr, bufs, 3, need.data(), peer_mapped_.data(),
                                &all_ok, stream_), "apk_p2p_map_peers");
  }
  if (!all_ok) {
    peer_mapped_.clear();
    return;
  }
  std::vector<std::array<double *, 3>> peer(nr, std::array<double *, 3>{nullptr, nullptr, nullptr});
  for (int r = 0; r < nr; ++r)
    for (int f = 0; f < 3; ++f) peer[r][f] = static_cast<double *>(peer_mapped_[(size_t)r * 3 + f]);
  const size_t per = BlockFieldElems(), nbl = blocks_.size();
  for (int which = 0; which < 2; ++which) {
    std::vector<apk_push_dst_t> tab(nbl * 27, apk_push_dst_t{nullptr, nullptr});
    for (size_t b = 0; b < nbl; ++b)
      for (int idx = 0; idx < 27; ++idx) {
        const int nb = h_blocks_[b].nbr[idx];
        if (nb == APK_NBR_NONE) continue;
        apk_push_dst_t &e = tab[b * 27 + idx];
        if (nb >= 0) {
          e.prim = d_prim_[1 - which] + nb * per;
          e.cons = d_cons_ + nb * per;
        } else {
          int r = 0, l = 0;
          PushTarget((int)b, idx, &r, &l);
          e.prim = peer[r][1 + (1 - which)] + (size_t)l * per;
          e.cons = peer[r][0] + (size_t)l * per;
        }
      }
    cuda_check(cudaMalloc(&d_push_[which], tab.size() * sizeof(apk_push_dst_t)), "push table");
    cuda_check(cudaMemcpy(d_push_[which], tab.data(), tab.size() * sizeof(apk_push_dst_t),
                          cudaMemcpyHostToDevice), "push table");
  }
  p2p_ = true;
}

// cross-rank barrier in stream order: every rank's stage kernel (and its peer stores) has
// completed before any rank's next kernel on this stream starts
void HydroSim::StageBarrier() {
  abi_check(apk_stage_barrier(comm_, d_barrier_, stream_), "apk_stage_barrier");
}

double *HydroSim::DevicePtr(int field, int block) {
  const size_t per = BlockFieldElems();
  switch (field) {
  case 0: return d_cons_ + block * per;
  case 1: return d_prim_[cur_] + block * per;
  case 2: return d_u1_ + block * per;
  case 3: case 4: case 5:
    return d_flux_[field - 3] ? d_flux_[field - 3] + block * per : nullptr;
  }
  return nullptr;
}

int HydroSim::CopyFromHost(int field, int block, const double *src) {
  double *p = DevicePtr(field, block);
  if (!p) return -1;
  cuda_check(cudaMemcpyAsync(p, src, BlockFieldElems() * sizeof(double), cudaMemcpyHostToDevice,
                             S(stream_)), "H2D");
  Synchronize();
  return 0;
}

int HydroSim::CopyToHost(int field, int block, double *dst) {
  double *p = DevicePtr(field, block);
  if (!p) return -1;
  cuda_check(cudaMemcpyAsync(dst, p, BlockFieldElems() * sizeof(double), cudaMemcpyDeviceToHost,
                             S(stream_)), "D2H");
  Synchronize();
  return 0;
}

void HydroSim::History(double *out8) {
  if (!d_hist_) cuda_check(cudaMalloc(&d_hist_, apk_history_workspace_bytes(&desc_)), "history");
  abi_check(apk_history_sums(&desc_, d_blocks_[cur_], d_hist_, out8, stream_), "apk_history_sums");
}

void HydroSim::EnableHistoryOutput(const std::string &fname, double dt, const std::string &data_format) {
  hst_name_ = fname;
  hst_dt_ = dt;
  hst_fmt_ = " " + (data_format.empty() ? std::string("%12.5e") : data_format);  // outputs.cpp:267-268
  hst_next_time_ = tm.time;
  hst_file_number_ = 0;
  hst_last_cycle_ = -1;
}

// history.cpp:172-209
void HydroSim::WriteHistoryRow() {
  double sums[8];
  History(sums);
  if (cfg_.nranks > 1) {  // MPI_Reduce(SUM) of history.cpp:146-168
    cuda_check(cudaMemcpyAsync(d_scalars_ + 16, sums, 8 * sizeof(double), cudaMemcpyHostToDevice,
                               S(stream_)), "history H2D");
    if (ncclAllReduce(d_scalars_ + 16, d_scalars_ + 16, 8, ncclDouble, ncclSum,
                      reinterpret_cast<ncclComm_t>(comm_), S(stream_)) != ncclSuccess)
      throw std::runtime_error("apk_host: ncclAllReduce (history) failed");
    cuda_check(cudaMemcpyAsync(sums, d_scalars_ + 16, 8 * sizeof(double), cudaMemcpyDeviceToHost,
                               S(stream_)), "history D2H");
    Synchronize();
  }
  hst_last_cycle_ = tm.ncycle;
  if (cfg_.rank != 0) {
    hst_file_number_++;
    return;
  }
  FILE *f = std::fopen(hst_name_.c_str(), hst_file_number_ == 0 ? "w" : "a");
  if (!f) throw std::runtime_error("Output file '" + hst_name_ + "' could not be opened");
  static const char *labels[8] = {"mass", "1-mom", "2-mom", "3-mom", "KE", "tot-E", "ME", "relDivB"};
  const int nq = (cfg_.fluid == APK_FLUID_GLMMHD) ? 8 : 6;
  if (hst_file_number_ == 0) {
    int iout = 1;
    std::fprintf(f, "#  History data\n");
    std::fprintf(f, "# [%d]=time    ", iout++);
    std::fprintf(f, " [%d]=dt      ", iout++);
    std::fprintf(f, " [%d]=cycle   ", iout++);
    std::fprintf(f, " [%d]=nbtotal ", iout++);
    for (int q = 0; q < nq; ++q) std::fprintf(f, " [%d]=%-8s", iout++, labels[q]);
    std::fprintf(f, "\n");
  }
  std::fprintf(f, hst_fmt_.c_str(), tm.time);
  std::fprintf(f, hst_fmt_.c_str(), tm.dt);
  std::fprintf(f, " %12d", tm.ncycle);
  std::fprintf(f, " %12d", nblocks_global_);
  for (int q = 0; q < nq; ++q) std::fprintf(f, hst_fmt_.c_str(), sums[q]);
  std::fprintf(f, "\n");
  std::fclose(f);
  hst_file_number_++;
}

// outputs.cpp:445-485 for one `hst` output block
void HydroSim::MakeOutputs(bool final_signal) {
  if (hst_name_.empty() || hst_dt_ < 0.0) return;
  const bool due = (tm.ncycle == 0) || (tm.time >= hst_next_time_) ||
                   (tm.tlim > 0.0 && tm.time >= tm.tlim) || final_signal;
  if (!due) return;
  WriteHistoryRow();
  if (hst_dt_ > 0.0) hst_next_time_ += hst_dt_;  // history.cpp:213-216
}

void HydroSim::FinalizeOutputs() {
  if (hst_last_cycle_ == tm.ncycle && tm.ncycle != 0) return;  // (already written by an earlier call)
  MakeOutputs(true);
}

size_t HydroSim::InteriorElems() const {
  return blocks_.size() * (size_t)nvar_ * cfg_.mb[0] * cfg_.mb[1] * cfg_.mb[2];
}

void HydroSim::EnsureStage() {
  if (d_stage_) return;
  const size_t per = (size_t)nvar_ * cfg_.mb[0] * cfg_.mb[1] * cfg_.mb[2];
  std::vector<apk_halo_item_t> items(blocks_.size());
  for (size_t b = 0; b < blocks_.size(); ++b) {
    apk_halo_item_t &it = items[b];
    it.block = (int)b;
    for (int d = 0; d < 3; ++d) {
      it.lo[d] = (d < ndim_) ? cfg_.ng : 0;
      it.ext[d] = cfg_.mb[d];
    }
    it.offset = (long long)(2 * b * per);  // single-field copies use offset / 2 (apk_halo_item_t)
  }
  cuda_check(cudaMalloc(&d_stage_, InteriorElems() * sizeof(double)), "interior staging");
  cuda_check(cudaMalloc(&d_int_items_, items.size() * sizeof(apk_halo_item_t)), "interior items");
  cuda_check(cudaMemcpy(d_int_items_, items.data(), items.size() * sizeof(apk_halo_item_t),
                        cudaMemcpyHostToDevice), "interior items");
}

int HydroSim::LoadStateInterior(const double *h_src) {
  EnsureStage();
  cuda_check(cudaMemcpyAsync(d_stage_, h_src, InteriorElems() * sizeof(double),
                             cudaMemcpyHostToDevice, S(stream_)), "H2D interior");
  abi_check(apk_halo_unpack(&desc_, d_blocks_[cur_], d_int_items_, (int)blocks_.size(), 0, d_stage_,
                            stream_), "scatter interior");
  // Mesh::Initialize sequence: exchange ghosts, BCs, FillDerived
  SendBoundBufs(0);
  ReceiveBoundBufs();
  SetBounds(0);
  ProlongateBounds();
  ApplyBoundaryConditions(0);
  FillDerived();
  return 0;
}

int HydroSim::StoreStateInterior(double *h_dst) {
  EnsureStage();
  abi_check(apk_halo_pack(&desc_, d_blocks_[cur_], d_int_items_, (int)blocks_.size(), 0, d_stage_,
                          stream_), "gather interior");
  cuda_check(cudaMemcpyAsync(h_dst, d_stage_, InteriorElems() * sizeof(double),
                             cudaMemcpyDeviceToHost, S(stream_)), "D2H interior");
  return 0;
}

int HydroSim::UploadAll(int field, const double *src) {
  double *p = DevicePtr(field, 0);
  if (!p) return -1;
  cuda_check(cudaMemcpyAsync(p, src, BlockFieldElems() * blocks_.size() * sizeof(double),
                             cudaMemcpyHostToDevice, S(stream_)), "H2D all");
  return 0;
}

int HydroSim::DownloadAll(int field, double *dst) {
  double *p = DevicePtr(field, 0);
  if (!p) return -1;
  cuda_check(cudaMemcpyAsync(dst, p, BlockFieldElems() * blocks_.size() * sizeof(double),
                             cudaMemcpyDeviceToHost, S(stream_)), "D2H all");
  return 0;
}

bool HydroSim::TaskTiming(const std::string &name, double *ms_total, long long *count) const {
  auto it = timing_.find(name);
  if (it == timing_.end()) return false;
  *ms_total = it->second.ms;
  *count = it->second.count;
  return true;
}

// run a task list with a CUDA event pair around every task (events on the launching stream)
TaskStatus HydroSim::ExecuteTimed(TaskList &tl) {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs(tl.tasks.size());
  for (size_t t = 0; t < tl.tasks.size(); ++t) {
    cuda_check(cudaEventCreate(&evs[t].first), "event");
    cuda_check(cudaEventCreate(&evs[t].second), "event");
    cuda_check(cudaEventRecord(evs[t].first, S(stream_)), "record");
    TaskStatus st;
    while ((st = tl.tasks[t].second()) == TaskStatus::incomplete) {
    }
    if (st == TaskStatus::fail) return st;
    cuda_check(cudaEventRecord(evs[t].second, S(stream_)), "record");
  }
  Synchronize();
  for (size_t t = 0; t < tl.tasks.size(); ++t) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, evs[t].first, evs[t].second);
    auto &T = timing_[tl.tasks[t].first];
    T.ms += ms;
    T.count += 1;
    cudaEventDestroy(evs[t].first);
    cudaEventDestroy(evs[t].second);
  }
  return TaskStatus::complete;
}

// ================================================================================================
// Tasks (hydro_driver.cpp:474-613)
// ================================================================================================
TaskStatus HydroSim::InitU1() {  // hydro_driver.cpp:474-495
  abi_check(apk_save_u1(&desc_, d_blocks_[cur_], stream_), "apk_save_u1");
  return TaskStatus::complete;
}

TaskStatus HydroSim::CalculateFluxes(int stage) {
  // flux_first_stage: VL2 predictor uses donor cell (hydro.cpp:459-464)
  const int recon = (stage == 1 && cfg_.integrator == (int)Integrator::vl2) ? APK_RECON_DC : -1;
  abi_check(apk_calc_fluxes(&desc_, d_blocks_[cur_], c_h, recon, stream_), "apk_calc_fluxes");
  if (cfg_.diffint == 1) CalcDiffFluxes();  // hydro.cpp:1202-1205 (diffusion/integrator = unsplit)
  return TaskStatus::complete;
}

TaskStatus HydroSim::CalcDiffFluxes() {  // diffusion/diffusion.cpp:19-53
  abi_check(apk_calc_diff_fluxes(&desc_, d_blocks_[cur_], &diff_, stream_), "apk_calc_diff_fluxes");
  return TaskStatus::complete;
}

// AddSTSTasks (hydro_driver.cpp:170-344): one RKL2 super time step of length tau (Meyer, Balsara &
// Aslam 2014), same task order: reset fluxes -> diffusive fluxes of base -> [flux correction] ->
// register update -> ghost exchange -> FillDerived, s_rkl times
TaskStatus HydroSim::SuperTimeStep(Real tau) {
  const Real mindt_diff = dt_diff_;
  // eq (21) using half hyperbolic timestep due to Strang split (hydro_driver.cpp:176-181)
  int s_rkl = static_cast<int>(0.5 * (std::sqrt(9.0 + 16.0 * tau / mindt_diff) - 1.0)) + 1;
  if (s_rkl % 2 == 0) s_rkl += 1;
  last_s_rkl = s_rkl;
  auto exchange_and_fill = [this] {
    SendBoundBufs(0);
    ReceiveBoundBufs();
    SetBounds(0);
    if (multilevel_) ProlongateBounds();
    ApplyBoundaryConditions(0);
    FillDerived();
  };
  abi_check(apk_reset_fluxes(&desc_, d_blocks_[cur_], stream_), "apk_reset_fluxes");
  CalcDiffFluxes();
  if (multilevel_) FluxCorrection();
  abi_check(apk_rkl2_step_first(&desc_, d_blocks_[cur_], d_my0_ptrs_, d_yjm2_ptrs_, s_rkl, tau,
                                stream_), "apk_rkl2_step_first");
  exchange_and_fill();
  // Meyer+2012 eq. (16) (hydro_driver.cpp:277-344)
  Real b_j = 1. / 3., b_jm1 = 1. / 3., b_jm2 = 1. / 3.;
  const Real w1 = 4. / (static_cast<Real>(s_rkl) * static_cast<Real>(s_rkl) +
                        static_cast<Real>(s_rkl) - 2.);
  for (int jj = 2; jj <= s_rkl; jj++) {
    const Real j = static_cast<Real>(jj);
    b_j = (j * j + j - 2.0) / (2 * j * (j + 1.0));
    const Real mu_j = (2.0 * j - 1.0) / j * b_j / b_jm1;
    const Real nu_j = -(j - 1.0) / j * b_j / b_jm2;
    const Real mu_tilde_j = mu_j * w1;
    const Real gamma_tilde_j = -(1.0 - b_jm1) * mu_tilde_j;  // -a_jm1*mu_tilde_j
    abi_check(apk_reset_fluxes(&desc_, d_blocks_[cur_], stream_), "apk_reset_fluxes");
    CalcDiffFluxes();
    if (multilevel_) FluxCorrection();
    abi_check(apk_rkl2_step_other(&desc_, d_blocks_[cur_], d_my0_ptrs_, d_yjm2_ptrs_, mu_j, nu_j,
                                  mu_tilde_j, gamma_tilde_j, tau, stream_), "apk_rkl2_step_other");
    exchange_and_fill();
    b_jm2 = b_jm1;
    b_jm1 = b_j;
  }
  return TaskStatus::complete;
}

// hydro.cpp:935-946: min of EstimateConductionTimestep (conduction.cpp:44-190),
// EstimateViscosityTimestep (viscosity.cpp:34-92) and EstimateResistivityTimestep
// (resistivity.cpp:33-84).  Fixed coefficients: the estimate depends on dx only.
Real HydroSim::EstimateDiffusionTimestep() {
  const Real fac = (ndim_ == 1) ? 0.5 : (ndim_ == 2 ? 0.25 : 1.0 / 6.0);
  const Real cfl_diff = (cfg_.cfl_diff > 0.0) ? cfg_.cfl_diff : cfg_.cfl;
  constexpr Real tiny = 1.0e-20;  // TINY_NUMBER
  Real dt_diff = std::numeric_limits<Real>::max();
  const Real fixed[3] = {cfg_.conduction == APK_CONDUCTION_ISOTROPIC ? cfg_.thermal_diff_coeff : -1.0,
                         cfg_.viscosity ? cfg_.mom_diff_coeff : -1.0,
                         cfg_.resistivity ? cfg_.ohm_diff_coeff : -1.0};
  if (cfg_.conduction == APK_CONDUCTION_ANISOTROPIC) {
    abi_check(apk_min_dt_conduction(&desc_, d_blocks_[cur_], &diff_, d_scalars_ + 2, stream_),
              "apk_min_dt_conduction");
    if (cfg_.nranks > 1)
      abi_check(apk_allreduce_scalar(comm_, d_scalars_ + 2, 0, stream_), "apk_allreduce_scalar");
    cuda_check(cudaMemcpyAsync(h_scalars_ + 2, d_scalars_ + 2, sizeof(double), cudaMemcpyDeviceToHost,
                               S(stream_)), "dt_cond D2H");
    Synchronize();
    dt_diff = std::min(dt_diff, cfl_diff * fac * h_scalars_[2]);
  }
  for (int c = 0; c < 3; ++c) {
    if (fixed[c] < 0.0) continue;
    Real m = std::numeric_limits<Real>::max();
    for (const auto &B : blocks_) {
      m = std::fmin(m, B.dx[0] * B.dx[0] / (fixed[c] + tiny));
      if (ndim_ >= 2) m = std::fmin(m, B.dx[1] * B.dx[1] / (fixed[c] + tiny));
      if (ndim_ >= 3) m = std::fmin(m, B.dx[2] * B.dx[2] / (fixed[c] + tiny));
    }
    dt_diff = std::min(dt_diff, cfl_diff * fac * m);
  }
  return dt_diff;
}

TaskStatus HydroSim::FirstOrderFluxCorrect(Real gam0, Real gam1, Real beta_dt) {
  long long n = 0;
  abi_check(apk_first_order_flux_correct(&desc_, d_blocks_[cur_], gam0, gam1, beta_dt, c_h,
                                         d_fofc_work_, &n, stream_), "apk_first_order_flux_correct");
  fofc_corrections += n;
  return TaskStatus::complete;
}

TaskStatus HydroSim::UpdateWithFluxDivergence(Real gam0, Real gam1, Real beta_dt) {
  abi_check(apk_flux_div_update(&desc_, d_blocks_[cur_], gam0, gam1, beta_dt, stream_),
            "apk_flux_div_update");
  return TaskStatus::complete;
}

TaskStatus HydroSim::AddUnsplitSources(Real beta_dt) {  // hydro.cpp:227-246
  if (cfg_.fluid == APK_FLUID_GLMMHD) {
    abi_check(apk_dedner_source(&desc_, d_blocks_[cur_], c_h, mindx, cfg_.glmmhd_alpha, beta_dt,
                                cfg_.dedner_extended, stream_), "apk_dedner_source");
  }
  return TaskStatus::complete;
}

TaskStatus HydroSim::FusedStage(int stage, Real gam0, Real gam1, Real beta_dt, bool last_stage) {
  const int recon = (stage == 1 && cfg_.integrator == (int)Integrator::vl2) ? APK_RECON_DC : -1;
  // VL2: u^{n+1} = u^n + dt L(w^{n+1/2}) never reads the predictor's cons (gam0 == 0), so stage 1
  // produces only the half-step primitives and u^n stays in cons, where stage 2 picks it up as
  // u1 (no u1 <- u0 copy, no predictor cons store).  Other integrators: u1 saved by stage 1.
  const bool vl2 = cfg_.integrator == (int)Integrator::vl2;
  int stage_flags = vl2 ? (stage == 1 ? APK_FUSED_PRIM_ONLY : APK_FUSED_U1_FROM_CONS)
                        : (stage == 1 ? APK_FUSED_SAVE_U1 : 0);
  unsigned long long *d_fofc_count = nullptr;
  if (cfg_.fofc) {
    // optimistic FOFC: count the cells that fail the first attempt and keep the stage's input
    // cons restorable -- VL2 stage 1 leaves cons alone, stage 1 of the other integrators and the
    // VL2 corrector save it into cons_u1, later stages of rk2 / rk3 into d_undo_
    stage_flags |= APK_FUSED_FOFC_CHECK;
    if (vl2 && stage == 2) stage_flags |= APK_FUSED_SAVE_U1;
    if (!vl2 && stage > 1)
      cuda_check(cudaMemcpyAsync(d_undo_, d_cons_, BlockFieldElems() * blocks_.size() * sizeof(double),
                                 cudaMemcpyDeviceToDevice, S(stream_)), "fofc undo copy");
    d_fofc_count = reinterpret_cast<unsigned long long *>(d_scalars_ + 16);
    cuda_check(cudaMemsetAsync(d_fofc_count, 0, sizeof(unsigned long long), S(stream_)), "fofc count");
  }
  double *dtmin = nullptr;
  if (last_stage) {
    dtmin = d_scalars_;
    // device-to-device from a resident DBL_MAX constant: fully asynchronous (a pageable H2D copy
    // stalls the launch queue, which shows at 64 meshblocks per GPU)
    cuda_check(cudaMemcpyAsync(d_scalars_, d_scalars_ + 8, sizeof(double), cudaMemcpyDeviceToDevice,
                               S(stream_)), "init dtmin");
  }
  apk_stage_args_t a{};
  a.gam0 = gam0;
  a.gam1 = gam1;
  a.beta_dt = beta_dt;
  a.c_h = c_h;
  a.mindx = mindx;
  a.alpha = cfg_.glmmhd_alpha;
  a.recon_override = recon;
  a.flags = stage_flags | APK_FUSED_PUSH_PRIM | (last_stage ? APK_FUSED_PUSH_CONS : 0);
  a.d_dtmin = dtmin;
  a.d_status = d_status_;
  a.d_push = p2p_ ? d_push_[cur_] : nullptr;
  a.d_fofc_count = d_fofc_count;
  a.d_tma_maps = d_tma_[cur_];
  abi_check(apk_stage_fused_ex(&desc_, d_blocks_[cur_], d_prim_out_[cur_], &a, stream_),
            "apk_stage_fused");
  cur_ ^= 1;  // prim_out becomes "prim"
  return TaskStatus::complete;
}

// FirstOrderFluxCorrect for a fused stage (hydro.cpp:1223-1342).  The stage kernel has counted the
// cells whose trial update fails the first attempt's test.  Zero (the usual case): FOFC would not
// have changed any flux and the fused result stands.  Otherwise the stage is rolled back and redone
// with the reference's task sequence (calc_flux -> first_order_flux_correct -> update -> source ->
// exchange -> fill_derived); the read-back blocks the host like the reference's own reduction.
TaskStatus HydroSim::FusedFofcCheck(int stage, Real gam0, Real gam1, Real beta_dt, bool last_stage) {
  unsigned long long *d_count = reinterpret_cast<unsigned long long *>(d_scalars_ + 16);
  if (cfg_.nranks > 1)
    abi_check(apk_allreduce_scalar(comm_, d_count, 1, stream_), "apk_allreduce_scalar");
  unsigned long long *h_count = reinterpret_cast<unsigned long long *>(h_scalars_ + 2);
  cuda_check(cudaMemcpyAsync(h_count, d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                             S(stream_)), "fofc count D2H");
  Synchronize();
  if (*h_count == 0) return TaskStatus::complete;
  fofc_redone_stages += 1;
  const bool vl2 = cfg_.integrator == (int)Integrator::vl2;
  const size_t bytes = BlockFieldElems() * blocks_.size() * sizeof(double);
  cur_ ^= 1;  // back to the stage's input primitives (the un-fused tasks work in place)
  // restore the stage's input cons
  if (vl2 && stage == 2)
    cuda_check(cudaMemcpyAsync(d_cons_, d_u1_, bytes, cudaMemcpyDeviceToDevice, S(stream_)), "undo");
  else if (!vl2 && stage == 1)
    cuda_check(cudaMemcpyAsync(d_cons_, d_u1_, bytes, cudaMemcpyDeviceToDevice, S(stream_)), "undo");
  else if (!vl2)
    cuda_check(cudaMemcpyAsync(d_cons_, d_undo_, bytes, cudaMemcpyDeviceToDevice, S(stream_)), "undo");
  if (stage == 1) InitU1();
  CalculateFluxes(stage);
  FirstOrderFluxCorrect(gam0, gam1, beta_dt);
  UpdateWithFluxDivergence(gam0, gam1, beta_dt);
  AddUnsplitSources(beta_dt);
  SendBoundBufs(0);
  ReceiveBoundBufs();
  SetBounds(0);
  ApplyBoundaryConditions(0);
  FillDerived();
  if (vl2 && stage == 1) {
    // back to the fused convention of the VL2 corrector: u^n lives in cons
    cuda_check(cudaMemcpyAsync(d_cons_, d_u1_, bytes, cudaMemcpyDeviceToDevice, S(stream_)),
               "restore u^n");
  }
  if (last_stage) last_stage_unfused_ = true;
  return TaskStatus::complete;
}

TaskStatus HydroSim::StartReceiveBoundBufs() {
  // NCCL receives are posted together with the sends inside one group (SendBoundBufs); the
  // task exists to keep the reference's task graph shape (hydro_driver.cpp:505-506).
  recv_posted_ = true;
  return TaskStatus::complete;
}

TaskStatus HydroSim::SendBoundBufs(int with_prim, bool local_done, bool need_barrier) {
  nccl_inflight_ = false;
  if (multilevel_ && cfg_.nranks > 1) {
    // refined mesh over several ranks: every rank has finished its update (barrier), then the
    // same-level ghost slabs of remote neighbours are pulled from the owners' memory
    if (!comm_) throw std::runtime_error("apk_host: multi-rank run without InitComm");
    StageBarrier();
    abi_check(apk_copy_ghosts_local(&desc_, d_blocks_[cur_], with_prim, stream_),
              "apk_copy_ghosts_local");
    abi_check(apk_box_copy(static_cast<const apk_box_item_t *>(ml_same_remote_.d), ml_same_remote_.n,
                           nvar_, stream_), "apk_box_copy (remote same level)");
    return TaskStatus::complete;
  }
  if (cfg_.nranks > 1 && local_done && p2p_) {
    // remote neighbours were served by peer stores of the stage kernel as well: no pack, no
    // send/recv, no unpack -- only the stage boundary
    if (need_barrier) StageBarrier();
    return TaskStatus::complete;
  }
  if (cfg_.nranks > 1) {
    nccl_inflight_ = true;
    if (!comm_) throw std::runtime_error("apk_host: multi-rank run without InitComm");
    // one pack launch for all peers, then one grouped send/recv per peer on the plan's COMM stream
    // so that the NVLink transfer overlaps the same-device ghost fill below
    abi_check(apk_halo_start(halo_plan_, &desc_, d_blocks_[cur_], with_prim, stream_),
              "apk_halo_start");
  }
  // same-device neighbours: direct ghost fill, no buffers (fused stages have already pushed
  // their boundary cells into the neighbours' ghost zones, APK_FUSED_PUSH_*)
  if (!local_done)
    abi_check(apk_copy_ghosts_local(&desc_, d_blocks_[cur_], with_prim, stream_),
              "apk_copy_ghosts_local");
  return TaskStatus::complete;
}

TaskStatus HydroSim::ReceiveBoundBufs() {
  // The reference polls MPI_Test and returns `incomplete` (boundary_communication.cpp:200-235);
  // apk_halo_test offers the same contract (event query).  The transfer is stream-ordered, so the
  // unpack can simply be enqueued behind it: by default the task completes immediately (no host
  // polling, no fence); APK_B200_POLL_HALO=1 exercises the polling contract instead.
  static const char *poll = std::getenv("APK_B200_POLL_HALO");
  if (poll && poll[0] == '1' && cfg_.nranks > 1 && nccl_inflight_) {
    const int st = apk_halo_test(halo_plan_);
    if (st == 1) return TaskStatus::incomplete;
    if (st != 0) return TaskStatus::fail;
  }
  return TaskStatus::complete;
}

TaskStatus HydroSim::SetBounds(int with_prim) {
  if (cfg_.nranks > 1 && nccl_inflight_)
    abi_check(apk_halo_finish(halo_plan_, &desc_, d_blocks_[cur_], with_prim, stream_),
              "apk_halo_finish");
  recv_posted_ = false;
  return TaskStatus::complete;
}

// ---- multilevel exchange plan ------------------------------------------------------------------
// What AddBoundaryExchangeTasks does on a multilevel mesh (boundary_communication.cpp:405-445; index
// ranges bnd_info.cpp:105-230, :354-445), regrouped into six launches over item lists:
//   1 same-level ghost copies (apk_copy_ghosts_local, the nbr[] table);
//   2 every fine block next to a coarser one restricts the cells along the shared face / edge /
//     corner into ITS coarse buffer (apk_restrict_average) ...
//   3 ... and those coarse cells are copied into the coarser block's ghost cells (apk_box_copy);
//   4 blocks with a coarser neighbour restrict their own cells -- interior and the ghost cells
//     filled by 1 and 3 -- into their coarse buffer (what ProResInfo::GetSet schedules);
//   5 the coarser neighbours' interior cells go into the coarse buffer, nghost deep (apk_box_copy);
//   6 ghost cells over coarser neighbours are prolongated from the coarse buffer
//     (apk_prolongate_minmod_multid) over the nghost/2 coarse layers that have fine ghost cells.
void HydroSim::BuildMultilevelPlan() {
  const int ng = cfg_.ng;
  const size_t per = BlockFieldElems(), cper = CoarseElems();
  // arrays of a block by GLOBAL id: local slabs, or the IPC-mapped slabs of the owning rank
  ml_desc_.clear();
  auto desc = [&](int list, int exec, int src, int dst, const int *lo, const int *ext) {
    MlDesc m{list, exec, src, dst, {lo[0], lo[1], lo[2]}, {ext[0], ext[1], ext[2]}};
    ml_desc_.push_back(m);
  };
  auto desc_amr = [&](int list, int g, const int *lo, const int *hi) {
    const int ext[3] = {hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1};
    desc(list, g, g, g, lo, ext);
  };
  auto cons_ptr = [&](int g) -> double * {
    if (IsLocal(g)) return d_cons_ + (size_t)(g - gid_lo_) * per;
    if (cfg_.plan_only) return nullptr;
    const int r = RankOfGid(g);
    return static_cast<double *>(ml_peer_[(size_t)r * 2 + 0]) + (size_t)(g - GidLo(r)) * per;
  };
  auto coarse_ptr = [&](int g) -> double * {
    if (IsLocal(g)) return d_coarse_ + (size_t)(g - gid_lo_) * cper;
    if (cfg_.plan_only) return nullptr;
    const int r = RankOfGid(g);
    return static_cast<double *>(ml_peer_[(size_t)r * 2 + 1]) + (size_t)(g - GidLo(r)) * cper;
  };
  auto flux_ptr = [&](int d, int g) { return d_flux_[d] + (size_t)(g - gid_lo_) * per; };  // local only
  const int cis = ng;
  auto cie = [&](int d) { return ng + cfg_.mb[d] / 2 - 1; };
  const int is[3] = {cfg_.ng, ndim_ >= 2 ? cfg_.ng : 0, ndim_ >= 3 ? cfg_.ng : 0};
  auto ie = [&](int d) { return is[d] + cfg_.mb[d] - 1; };
  // index range of the region of neighbour offset o in direction d (fine or coarse index space)
  auto range = [&](int d, int o, bool coarse, int depth, int *lo, int *hi) {
    if (d >= ndim_) { *lo = *hi = 0; return; }
    const int s0 = coarse ? cis : is[d], t0 = coarse ? cie(d) : ie(d);
    if (o == 0) { *lo = s0; *hi = t0; }
    else if (o > 0) { *lo = t0 + 1; *hi = t0 + depth; }
    else { *lo = s0 - depth; *hi = s0 - 1; }
  };
  // restriction / prolongation item on a LOCAL block g (its arrays and coordinates)
  auto amr_item = [&](double *fine, int g, const int *lo, const int *hi) {
    const BlockInfo &L = blocks_[g - gid_lo_];
    apk_amr_item_t it{};
    it.fine = fine;
    it.coarse = coarse_ptr(g);
    for (int d = 0; d < 3; ++d) {
      it.nf[d] = n_[d];
      it.nc[d] = nc_[d];
      it.lo[d] = lo[d];
      it.hi[d] = hi[d];
      it.xmin[d] = L.bxmin[d];
      it.dx[d] = L.dx[d];
    }
    return it;
  };
  std::vector<apk_amr_item_t> restrict_child, restrict_own, prolong, fc_restrict[3];
  std::vector<apk_box_item_t> child_to_ghost, coarse_from_nbr, fc_copy[3], same_remote;
  for (int b = 0; b < nblocks_global_; ++b) {  // GLOBAL ids: an item exists where its executor is local
    const Leaf &B = leaves_[b];
    const NbrInfo &N = gnbr_[b];
    const bool b_local = IsLocal(b);
    for (int idx = 0; idx < 27; ++idx) {
      const int off[3] = {idx % 3 - 1, (idx / 3) % 3 - 1, idx / 9 - 1};
      if (N.type[idx] == 3) {
        for (int c = 0; c < N.nchild[idx]; ++c) {
          const int f = N.idx[idx][c];
          int clo[3], chi[3];
          apk_box_item_t bx{};
          for (int d = 0; d < 3; ++d) {
            if (d >= ndim_) { clo[d] = chi[d] = 0; }
            else if (off[d] > 0) { clo[d] = cis; chi[d] = cis + ng - 1; }
            else if (off[d] < 0) { clo[d] = cie(d) - ng + 1; chi[d] = cie(d); }
            else { clo[d] = cis; chi[d] = cie(d); }
            bx.ns[d] = nc_[d];
            bx.nd[d] = n_[d];
            bx.s0[d] = clo[d];
            bx.ext[d] = chi[d] - clo[d] + 1;
            if (d >= ndim_) bx.d0[d] = 0;
            else if (off[d] > 0) bx.d0[d] = ie(d) + 1;
            else if (off[d] < 0) bx.d0[d] = is[d] - ng;
            else bx.d0[d] = is[d] + (leaves_[f].lx[d] & 1) * (cfg_.mb[d] / 2);
          }
          // the fine block's owner restricts into the fine block's coarse buffer ...
          if (IsLocal(f)) {
            restrict_child.push_back(amr_item(cons_ptr(f), f, clo, chi));
            desc_amr(0, f, clo, chi);
          }
          // ... and the coarse block's owner pulls the result into its ghost cells
          if (b_local) {
            bx.src = coarse_ptr(f);
            bx.dst = cons_ptr(b);
            child_to_ghost.push_back(bx);
            desc(1, b, f, b, bx.s0, bx.ext);
          }
        }
      }
      if (!b_local) continue;
      if (N.type[idx] == 1 && !IsLocal(N.idx[idx][0])) {
        // same-level neighbour on another rank: pull its interior slab into my ghost slab
        // (bnd_info.cpp:136-230; local same-level neighbours go through the nbr[] table)
        apk_box_item_t bx{};
        for (int d = 0; d < 3; ++d) {
          int lo, hi;
          range(d, off[d], false, ng, &lo, &hi);
          bx.ns[d] = bx.nd[d] = n_[d];
          bx.d0[d] = lo;
          bx.ext[d] = hi - lo + 1;
          if (d >= ndim_) bx.s0[d] = 0;
          else if (off[d] > 0) bx.s0[d] = is[d];
          else if (off[d] < 0) bx.s0[d] = ie(d) - ng + 1;
          else bx.s0[d] = is[d];
        }
        bx.src = cons_ptr(N.idx[idx][0]);
        bx.dst = cons_ptr(b);
        same_remote.push_back(bx);
        desc(5, b, N.idx[idx][0], b, bx.d0, bx.ext);
      }
      if (!N.has_coarser) continue;
      int lo[3], hi[3];
      if (idx == 13 || N.type[idx] == 1 || N.type[idx] == 3) {
        for (int d = 0; d < 3; ++d) range(d, off[d], true, ng / 2, &lo[d], &hi[d]);
        restrict_own.push_back(amr_item(cons_ptr(b), b, lo, hi));
        desc_amr(2, b, lo, hi);
      } else if (N.type[idx] == 2) {
        const int cn = N.idx[idx][0];
        apk_box_item_t bx{};
        for (int d = 0; d < 3; ++d) {
          range(d, off[d], true, ng, &lo[d], &hi[d]);
          bx.ns[d] = n_[d];
          bx.nd[d] = nc_[d];
          bx.d0[d] = lo[d];
          bx.ext[d] = hi[d] - lo[d] + 1;
          if (d >= ndim_) { bx.s0[d] = 0; continue; }
          // global cell index of the first buffer cell on the coarser level (periodic wrap)
          const long tot = (long)(nb_[d] << (B.level - 1)) * cfg_.mb[d];
          long G = (long)B.lx[d] * (cfg_.mb[d] / 2) + (lo[d] - ng);
          G = ((G % tot) + tot) % tot;
          bx.s0[d] = is[d] + (int)(G - (long)leaves_[cn].lx[d] * cfg_.mb[d]);
          if (bx.s0[d] < is[d] || bx.s0[d] + bx.ext[d] - 1 > ie(d))
            throw std::runtime_error("apk_host: coarse neighbour does not cover a buffer region");
        }
        bx.src = cons_ptr(cn);
        bx.dst = coarse_ptr(b);
        coarse_from_nbr.push_back(bx);
        desc(3, b, cn, b, bx.d0, bx.ext);
        for (int d = 0; d < 3; ++d) range(d, off[d], true, ng / 2, &lo[d], &hi[d]);
        prolong.push_back(amr_item(cons_ptr(b), b, lo, hi));
        desc_amr(4, b, lo, hi);
      }
    }
    // flux correction: faces shared with finer blocks (bvals_in_one.hpp:69-80, bnd_info.cpp:71-103)
    for (int d = 0; d < ndim_; ++d)
      for (int side = -1; side <= 1; side += 2) {
        int off[3] = {0, 0, 0};
        off[d] = side;
        const int idx = (off[2] + 1) * 9 + (off[1] + 1) * 3 + (off[0] + 1);
        if (N.type[idx] != 3) continue;
        for (int c = 0; c < N.nchild[idx]; ++c) {
          const int f = N.idx[idx][c];
          int clo[3], chi[3];
          apk_box_item_t bx{};
          for (int dd = 0; dd < 3; ++dd) {
            if (dd >= ndim_) { clo[dd] = chi[dd] = 0; }
            else if (dd == d) { clo[dd] = chi[dd] = (side > 0) ? cis : cie(dd) + 1; }
            else { clo[dd] = cis; chi[dd] = cie(dd); }
            bx.ns[dd] = nc_[dd];
            bx.nd[dd] = n_[dd];
            bx.s0[dd] = clo[dd];
            bx.ext[dd] = chi[dd] - clo[dd] + 1;
            if (dd >= ndim_) bx.d0[dd] = 0;
            else if (dd == d) bx.d0[dd] = (side > 0) ? ie(dd) + 1 : is[dd];
            else bx.d0[dd] = is[dd] + (leaves_[f].lx[dd] & 1) * (cfg_.mb[dd] / 2);
          }
          // the child's coarse buffer is free between exchanges: scratch for the restricted faces
          if (IsLocal(f)) {
            fc_restrict[d].push_back(amr_item(flux_ptr(d, f), f, clo, chi));
            desc_amr(6 + d, f, clo, chi);
          }
          if (b_local) {
            bx.src = coarse_ptr(f);
            bx.dst = flux_ptr(d, b);
            fc_copy[d].push_back(bx);
            desc(9 + d, b, f, b, bx.d0, bx.ext);
          }
        }
      }
  }
  auto upload = [&](MlList &dst, const void *src, size_t n, size_t elem) {
    dst.n = (int)n;
    if (!n || cfg_.plan_only) return;
    dst.d = SlabAlloc(n * elem, "cudaMalloc multilevel plan");
    cuda_check(cudaMemcpy(dst.d, src, n * elem, cudaMemcpyHostToDevice), "upload multilevel plan");
  };
  upload(ml_restrict_child_, restrict_child.data(), restrict_child.size(), sizeof(apk_amr_item_t));
  upload(ml_child_to_ghost_, child_to_ghost.data(), child_to_ghost.size(), sizeof(apk_box_item_t));
  upload(ml_restrict_own_, restrict_own.data(), restrict_own.size(), sizeof(apk_amr_item_t));
  upload(ml_coarse_from_nbr_, coarse_from_nbr.data(), coarse_from_nbr.size(), sizeof(apk_box_item_t));
  upload(ml_prolong_, prolong.data(), prolong.size(), sizeof(apk_amr_item_t));
  upload(ml_same_remote_, same_remote.data(), same_remote.size(), sizeof(apk_box_item_t));
  for (int d = 0; d < 3; ++d) {
    upload(ml_fc_restrict_[d], fc_restrict[d].data(), fc_restrict[d].size(), sizeof(apk_amr_item_t));
    upload(ml_fc_copy_[d], fc_copy[d].data(), fc_copy[d].size(), sizeof(apk_box_item_t));
  }
}

TaskStatus HydroSim::ProlongateBounds() {
  if (!multilevel_) return TaskStatus::complete;
  const int fp = cfg_.fp_mode;
  abi_check(apk_restrict_average(ndim_, cfg_.ng, nvar_, 0,
                                 static_cast<const apk_amr_item_t *>(ml_restrict_child_.d),
                                 ml_restrict_child_.n, fp, stream_), "apk_restrict_average");
  // (several ranks: the fine blocks' coarse buffers are complete everywhere before they are pulled)
  if (cfg_.nranks > 1) StageBarrier();
  abi_check(apk_box_copy(static_cast<const apk_box_item_t *>(ml_child_to_ghost_.d),
                         ml_child_to_ghost_.n, nvar_, stream_), "apk_box_copy");
  abi_check(apk_restrict_average(ndim_, cfg_.ng, nvar_, 0,
                                 static_cast<const apk_amr_item_t *>(ml_restrict_own_.d),
                                 ml_restrict_own_.n, fp, stream_), "apk_restrict_average");
  abi_check(apk_box_copy(static_cast<const apk_box_item_t *>(ml_coarse_from_nbr_.d),
                         ml_coarse_from_nbr_.n, nvar_, stream_), "apk_box_copy");
  // cbound (boundary_communication.cpp:437-440): physical BCs on the coarse buffers
  if (bc_face_mask_ && d_cblocks_)
    abi_check(apk_apply_physical_bcs(&desc_coarse_, d_cblocks_, bc_face_mask_, 0, stream_),
              "apk_apply_physical_bcs (coarse buffers)");
  abi_check(apk_prolongate_minmod_multid(ndim_, cfg_.ng, nvar_,
                                         static_cast<const apk_amr_item_t *>(ml_prolong_.d),
                                         ml_prolong_.n, fp, stream_), "apk_prolongate_minmod_multid");
  // (several ranks: nobody modifies interior cells or coarse buffers while a peer still pulls them)
  if (cfg_.nranks > 1) StageBarrier();
  return TaskStatus::complete;
}

// hydro_driver.cpp:527-531 (AddFluxCorrectionTasks): the flux of a coarse face shared with finer
// blocks is replaced by the area-weighted mean of the fine faces
TaskStatus HydroSim::FluxCorrection() {
  if (!multilevel_) return TaskStatus::complete;
  for (int d = 0; d < ndim_; ++d) {
    abi_check(apk_restrict_average(ndim_, cfg_.ng, nvar_, d + 1,
                                   static_cast<const apk_amr_item_t *>(ml_fc_restrict_[d].d),
                                   ml_fc_restrict_[d].n, cfg_.fp_mode, stream_),
              "apk_restrict_average (flux)");
    if (cfg_.nranks > 1) StageBarrier();
    abi_check(apk_box_copy(static_cast<const apk_box_item_t *>(ml_fc_copy_[d].d), ml_fc_copy_[d].n,
                           nvar_, stream_), "apk_box_copy (flux)");
    // the next direction reuses the same scratch cells of the coarse buffers
    if (cfg_.nranks > 1 && d + 1 < ndim_) StageBarrier();
  }
  return TaskStatus::complete;
}

TaskStatus HydroSim::ApplyBoundaryConditions(int with_prim) {
  if (bc_face_mask_)
    abi_check(apk_apply_physical_bcs(&desc_, d_blocks_[cur_], bc_face_mask_, with_prim,
                                     stream_), "apk_apply_physical_bcs");
  return TaskStatus::complete;
}

TaskStatus HydroSim::FillDerived() {  // Update::FillDerived -> ConsToPrim (hydro.cpp:707,712)
  abi_check(apk_cons_to_prim(&desc_, d_blocks_[cur_], d_status_, stream_), "apk_cons_to_prim");
  return TaskStatus::complete;
}

TaskStatus HydroSim::EstimateTimestep(bool have_device_min) {
  // hydro.cpp:828-976 + update.hpp:269-278. have_device_min: the fused last stage already left
  // min dx/(|v|+c) in d_scalars_[0].
  Real min_dt = std::numeric_limits<Real>::max();
  Real dt_hyp_local = std::numeric_limits<Real>::max();
  if (cfg_.riemann != APK_RIEMANN_NONE) {  // calc_dt_hyp = false otherwise, hydro.cpp:354-361
    if (!have_device_min)
      abi_check(apk_min_dt_async(&desc_, d_blocks_[cur_], d_scalars_, stream_), "apk_min_dt_async");
    if (cfg_.nranks > 1) {
      // replaces MPI_Allreduce(MIN) of driver.cpp:237 / hydro.cpp:127
      abi_check(apk_allreduce_scalar(comm_, d_scalars_, 0, stream_), "apk_allreduce_scalar");
    }
    // pinned destinations: both copies are truly asynchronous, one synchronisation per cycle
    int *h_status = reinterpret_cast<int *>(h_scalars_ + 1);
    cuda_check(cudaMemcpyAsync(h_scalars_, d_scalars_, sizeof(double), cudaMemcpyDeviceToHost,
                               S(stream_)), "dt D2H");
    cuda_check(cudaMemcpyAsync(h_status, d_status_, sizeof(int), cudaMemcpyDeviceToHost, S(stream_)),
               "status D2H");
    Synchronize();
    const double m = h_scalars_[0];
    const int status = *h_status;
    if (status != 0) {
      // PARTHENON_REQUIRE in ConsToPrim (adiabatic_hydro.hpp:77-79,111-113)
      throw std::runtime_error(status == 1
                                   ? "Got negative density. Consider enabling first-order flux "
                                     "correction or setting a reasonble density floor."
                                   : "Got negative pressure. Consider enabling first-order flux "
                                     "correction or setting a reasonble pressure or temperature floor.");
    }
    const Real dt = cfg_.cfl * m;
    if (cfg_.fluid == APK_FLUID_GLMMHD && dt < dt_hyp) dt_hyp = dt;  // hydro.cpp:903-908
    dt_hyp_local = dt;
    min_dt = std::min(min_dt, dt);
  }
  if (cfg_.diffint != 0) {  // hydro.cpp:935-963
    const Real dt_diff = EstimateDiffusionTimestep();
    if (cfg_.diffint == 1) {
      min_dt = std::min(min_dt, dt_diff);
    } else {
      const Real r = cfg_.rkl2_max_dt_ratio;
      if (r > 0.0 && dt_hyp_local / dt_diff > r) min_dt = std::min(min_dt, r * dt_diff);
    }
    dt_diff_ = std::min(dt_diff, dt_diff_);
  }
  new_dt_ = min_dt;
  return TaskStatus::complete;
}

void HydroSim::PreStepMeshUserWorkInLoop() {  // hydro.cpp:102-143 (+ CalculateGlobalMinDx :65-95)
  if (cfg_.fluid != APK_FLUID_GLMMHD) return;
  Real m = std::numeric_limits<Real>::max();
  if (multilevel_) {
    mindx = global_mindx_;
    c_h = cfg_.cfl * mindx / dt_hyp;
    return;
  }
  // dx is identical on all ranks of a uniform mesh; no reduction needed
  for (const auto &B : blocks_) {
    m = std::fmin(m, B.dx[0]);
    if (n_[1] > 1) m = std::fmin(m, B.dx[1]);
    if (n_[2] > 1) m = std::fmin(m, B.dx[2]);
  }
  mindx = m;
  c_h = cfg_.cfl * mindx / dt_hyp;
}

void HydroSim::SetGlobalTimeStep() {  // driver.cpp:210-242,267-269 (default dt_* knobs)
  if (tm.dt < 0.1 * std::numeric_limits<Real>::max()) tm.dt *= 2.0;
  tm.dt = std::min(tm.dt, new_dt_);
  new_dt_ = std::numeric_limits<Real>::max();
  if (tm.time < tm.tlim && (tm.tlim - tm.time) < tm.dt) tm.dt = tm.tlim - tm.time;
}

void HydroSim::Initialize(double tlim) {
  tm = SimTime{};
  tm.tlim = tlim;
  cur_ = 0;
  // Mesh::Initialize: exchange ghosts, BCs, FillDerived
  auto fill = [this] {
    SendBoundBufs(0);
    ReceiveBoundBufs();
    SetBounds(0);
    ProlongateBounds();
    ApplyBoundaryConditions(0);
    FillDerived();
  };
  fill();
  if (cfg_.adaptive) {
    // mesh.cpp:746-840: tag, refine, re-run the problem generator on the new mesh until the number
    // of meshblocks stops changing
    if (pgen_id_.empty())
      throw std::runtime_error("apk_host: adaptive meshes are initialised through ProblemGenerator");
    for (;;) {
      TagRefinement();
      const int nb_before = nblocks_global_;
      LoadBalancingAndAdaptiveMeshRefinement();
      if (nblocks_global_ == nb_before) break;
      ProblemGenerator(pgen_id_, pgen_pars_);
      fill();
    }
  }
  dt_hyp = std::numeric_limits<Real>::max();
  dt_diff_ = std::numeric_limits<Real>::max();
  EstimateTimestep(false);
  SetGlobalTimeStep();
}

void HydroSim::MakeTaskCollection(TaskList &tl, int stage) {
  const Real gam0 = integrator_.gam0[stage - 1], gam1 = integrator_.gam1[stage - 1];
  const Real beta_dt = integrator_.beta[stage - 1] * integrator_.dt;
  const bool last = (stage == integrator_.nstages);
  // fused path: the u1 <- u0 copy is folded into the first fused stage (APK_FUSED_SAVE_U1)
  // Strang-split super time stepping, first half (hydro_driver.cpp:455-458)
  if (stage == 1 && cfg_.diffint == 2)
    tl.AddTask("sts_first_half", [this] { return SuperTimeStep(0.5 * tm.dt); });
  if (stage == 1 && !fused_) tl.AddTask("init_u1", [this] { return InitU1(); });
  tl.AddTask("start_bnd", [this] { return StartReceiveBoundBufs(); });
  if (fused_) {
    // calc_flux + update + source_unsplit + cons2prim(interior) [+ dt reduce] in one kernel;
    // ghosts of cons AND prim are then exchanged together (pointwise cons->prim commutes with
    // the copy), so the separate FillDerived region disappears.
    tl.AddTask(stage == 1 ? "fused_stage_first" : "fused_stage_other",
               [=] { return FusedStage(stage, gam0, gam1, beta_dt, last); });
    if (cfg_.fofc)
      tl.AddTask("first_order_flux_correct",
                 [=] { return FusedFofcCheck(stage, gam0, gam1, beta_dt, last); });
    // between stages only prim ghosts are consumed; the last stage also completes cons so the
    // state handed back after the cycle is identical to the reference's, ghosts included
    const int fields = last ? 1 : 2;
    // p2p mode: the dt all-reduce of the last stage doubles as the stage barrier unless physical
    // BCs (which read neighbour-filled ghost cells) run before it
    const bool barrier = !last || bc_face_mask_ != 0 || cfg_.riemann == APK_RIEMANN_NONE;
    tl.AddTask("send", [=] { return SendBoundBufs(fields, true, barrier); });
    tl.AddTask("recv", [this] { return ReceiveBoundBufs(); });
    tl.AddTask("set", [=] { return SetBounds(fields); });
    tl.AddTask("fbound", [=] { return ApplyBoundaryConditions(fields); });
  } else {
    tl.AddTask("calc_flux", [=] { return CalculateFluxes(stage); });
    if (cfg_.fofc)
      tl.AddTask("first_order_flux_correct",
                 [=] { return FirstOrderFluxCorrect(gam0, gam1, beta_dt); });
    if (multilevel_) tl.AddTask("flux_correction", [this] { return FluxCorrection(); });
    tl.AddTask("update", [=] { return UpdateWithFluxDivergence(gam0, gam1, beta_dt); });
    tl.AddTask("source_unsplit", [=] { return AddUnsplitSources(beta_dt); });
    tl.AddTask("send", [this] { return SendBoundBufs(0); });
    tl.AddTask("recv", [this] { return ReceiveBoundBufs(); });
    tl.AddTask("set", [this] { return SetBounds(0); });
    if (multilevel_) tl.AddTask("prolongate_bounds", [this] { return ProlongateBounds(); });
    tl.AddTask("fbound", [this] { return ApplyBoundaryConditions(0); });
    tl.AddTask("fill_derived", [this] { return FillDerived(); });
  }
  if (last) {
    // second half of the super time step (hydro_driver.cpp:578-583)
    if (cfg_.diffint == 2)
      tl.AddTask("sts_second_half", [this] { return SuperTimeStep(0.5 * tm.dt); });
    // hydro_driver.cpp:589-613: reset reduction Params, then EstimateTimestep
    tl.AddTask("reset_reduction_vars", [this] {
      mindx = std::numeric_limits<Real>::max();
      dt_hyp = std::numeric_limits<Real>::max();
      dt_diff_ = std::numeric_limits<Real>::max();
      return TaskStatus::complete;
    });
    tl.AddTask("new_dt", [this] {
      const bool have = fused_ && !last_stage_unfused_;
      last_stage_unfused_ = false;
      return EstimateTimestep(have);
    });
    // hydro_driver.cpp:662-670
    if (cfg_.adaptive) tl.AddTask("tag_refine", [this] { return TagRefinement(); });
  }
}

TaskStatus HydroSim::Step() {  // multistage.hpp:39-53
  integrator_.dt = tm.dt;
  for (int stage = 1; stage <= integrator_.nstages; ++stage) {
    TaskList tl;
    MakeTaskCollection(tl, stage);
    const TaskStatus st = timing_on_ ? ExecuteTimed(tl) : tl.Execute();
    if (st != TaskStatus::complete) return st;
  }
  // Integrators with an odd number of stages leave the final primitives in the second ping-pong
  // buffer: cur_ simply persists into the next cycle (every table is indexed by parity).  No
  // copy-back: in p2p mode a copy here would race with the neighbour rank's next stage-1 kernel,
  // which already peer-stores into the buffer being copied.
  return TaskStatus::complete;
}

int HydroSim::Execute(double tlim, int nlim) {  // driver.cpp:67-165
  const int n0 = tm.ncycle;
  if (tlim != tm.tlim) {
    tm.tlim = tlim;
    if (tm.time < tm.tlim && (tm.tlim - tm.time) < tm.dt) tm.dt = tm.tlim - tm.time;
  }
  tm.nlim = nlim;
  // driver.cpp:88-93: the initial dump
  if (tm.ncycle == 0 && hst_file_number_ == 0) MakeOutputs(false);
  while (tm.KeepGoing()) {
    PreStepMeshUserWorkInLoop();
    if (Step() != TaskStatus::complete) throw std::runtime_error("Step failed to complete all tasks.");
    tm.ncycle++;
    tm.time += tm.dt;
    cycles_done++;
    // driver.cpp:128-131: remesh, and new block time steps when the mesh changed
    if (cfg_.adaptive && LoadBalancingAndAdaptiveMeshRefinement()) EstimateTimestep(false);
    SetGlobalTimeStep();
    // driver.cpp:141-143: the last output of a run is left to the `final` signal
    if (tm.KeepGoing()) MakeOutputs(false);
  }
  Synchronize();
  return tm.ncycle - n0;
}

}  // namespace apk_host
