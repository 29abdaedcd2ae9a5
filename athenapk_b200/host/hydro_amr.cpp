// hydro_amr.cpp -- adaptive mesh refinement of the host driver mirror (BASELINE config 5 as the
// reference runs it: inputs/blast_3d_amr.in, parthenon/mesh/refinement = adaptive).  What is
// mirrored, on one rank:
//   * Refinement::Tag<MeshBlockData> + MeshRefinement::SetRefinement
//     (external/parthenon/src/amr_criteria/refinement_package.cpp:205-233,
//      mesh/mesh_refinement.cpp:72-110) with AthenaPK's pressure-gradient criterion
//     (src/refinement/gradient.cpp:18-61) evaluated on the device (apk_tag_pressure_gradient);
//   * Mesh::UpdateMeshBlockTree (mesh/mesh-amr_loadbalance.cpp:485-604) on a leaf set with
//     Tree::Refine / Tree::Derefine and their proper-nesting rules (mesh/forest/tree.cpp:95-272);
//   * Mesh::RedistributeAndRefineMeshBlocks (mesh-amr_loadbalance.cpp:626-995): blocks that stay
//     move as they are, new fine blocks get the octant of their parent (ghost cells included) in a
//     coarse buffer and are prolongated with ProlongateCellMinModMultiD, new coarse blocks are the
//     restriction of their 2^ndim children; then ghost cells and derived fields are rebuilt.
// Host C++ only: every device operation goes through the C-ABI of include/apk_b200.h
// (apk_box_copy, apk_prolongate_minmod_multid, apk_restrict_average, apk_tag_pressure_gradient).
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "hydro_sim.hpp"

namespace apk_host {

namespace {
void cuda_check(cudaError_t e, const char *what) {
  if (e != cudaSuccess)
    throw std::runtime_error(std::string("apk_host: CUDA failure in ") + what + ": " +
                             cudaGetErrorString(e));
}
void abi_check(int rc, const char *what) {
  if (rc != 0)
    throw std::runtime_error(std::string("apk_host: ") + what + " failed: " + apk_last_error());
}
cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }

struct Loc {
  int level;
  std::array<int, 3> lx;
  bool operator<(const Loc &o) const { return std::tie(level, lx) < std::tie(o.level, o.lx); }
  bool operator==(const Loc &o) const { return level == o.level && lx == o.lx; }
};

// the leaves of the forest plus the geometry needed to walk it
struct Forest {
  std::set<Loc> leaves;
  int ndim, nb[3];
  bool periodic[3];

  // location of the same-level neighbour at offset `off`; false when it lies outside the domain
  bool Neighbor(const Loc &l, const int *off, Loc *out) const {
    out->level = l.level;
    for (int d = 0; d < 3; ++d) {
      int v = l.lx[d] + off[d];
      if (d >= ndim) {
        if (off[d]) return false;
        out->lx[d] = 0;
        continue;
      }
      const int tot = nb[d] << l.level;
      if (v < 0 || v >= tot) {
        if (!periodic[d]) return false;
        v = (v + tot) % tot;
      }
      out->lx[d] = v;
    }
    return true;
  }
  Loc Child(const Loc &l, int c) const {
    Loc ch{l.level + 1, {0, 0, 0}};
    for (int d = 0; d < ndim; ++d) ch.lx[d] = 2 * l.lx[d] + ((c >> d) & 1);
    return ch;
  }
  static Loc Parent(const Loc &l) { return Loc{l.level - 1, {l.lx[0] >> 1, l.lx[1] >> 1, l.lx[2] >> 1}}; }
  // a location is an internal node when neither it nor one of its ancestors is a leaf
  bool Internal(const Loc &l) const {
    Loc a = l;
    for (;;) {
      if (leaves.count(a)) return false;
      if (a.level == 0) return true;
      a = Parent(a);
    }
  }

  // Tree::Refine(ref_loc, enforce_proper_nesting = true), tree.cpp:95-136
  int Refine(const Loc &l) {
    auto it = leaves.find(l);
    if (it == leaves.end()) return 0;  // can't refine a block that doesn't exist
    leaves.erase(it);
    const int nchild = 1 << ndim;
    for (int c = 0; c < nchild; ++c) leaves.insert(Child(l, c));
    int nadded = nchild - 1;
    if (l.level > 0) {
      // the 2^ndim blocks on level - 1 that could abut the new blocks on level + 1
      const Loc parent = Parent(l);
      for (int k = -(ndim > 2); k < 1; ++k)
        for (int j = -(ndim > 1); j < 1; ++j)
          for (int i = -1; i < 1; ++i) {
            const int off[3] = {i + (l.lx[0] & 1), ndim > 1 ? j + (l.lx[1] & 1) : 0,
                                ndim > 2 ? k + (l.lx[2] & 1) : 0};
            Loc n;
            if (Neighbor(parent, off, &n)) nadded += Refine(n);
          }
    }
    return nadded;
  }

  // Tree::Derefine(ref_loc, enforce_proper_nesting = true), tree.cpp:217-272
  int Derefine(const Loc &l) {
    const int nchild = 1 << ndim;
    for (int c = 0; c < nchild; ++c) {
      const Loc d = Child(l, c);
      if (!leaves.count(d)) return 0;
      for (int k = -(ndim > 2); k <= (ndim > 2); ++k)
        for (int j = -(ndim > 1); j <= (ndim > 1); ++j)
          for (int i = -1; i <= 1; ++i) {
            const int off[3] = {i, j, k};
            Loc n;
            if (Neighbor(d, off, &n) && Internal(n)) return 0;
          }
    }
    for (int c = 0; c < nchild; ++c) leaves.erase(Child(l, c));
    leaves.insert(l);
    return nchild - 1;
  }
};

double NowMs() {
  return std::chrono::duration<double, std::milli>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

std::uint64_t Interleave(const int *x) {
  std::uint64_t m = 0;
  for (int bit = 0; bit < 20; ++bit)
    for (int d = 0; d < 3; ++d) m |= (std::uint64_t)((x[d] >> bit) & 1) << (3 * bit + d);
  return m;
}
}  // namespace

// ---- Refinement::Tag on every block ------------------------------------------------------------
TaskStatus HydroSim::TagRefinement() {
  const double t_start = NowMs();
  const size_t nbl = blocks_.size();
  if (tags_cap_ < nbl) {
    cudaFree(d_tags_);
    cudaFree(d_maxeps_);
    tags_cap_ = nbl + nbl / 2 + 64;
    cuda_check(cudaMalloc(&d_tags_, tags_cap_ * sizeof(int)), "cudaMalloc tags");
    cuda_check(cudaMalloc(&d_maxeps_, tags_cap_ * sizeof(double)), "cudaMalloc maxeps");
  }
  // thresholds of the criterion (gradient.cpp:58-60,91-93; other.cpp:39-41)
  const double above = cfg_.refine_threshold;
  const double below = (cfg_.refine_type == APK_REFINE_PRESSURE_GRADIENT)     ? 0.25 * above
                       : (cfg_.refine_type == APK_REFINE_XYVELOCITY_GRADIENT) ? 0.5 * above
                                                                               : cfg_.refine_threshold2;
  abi_check(apk_tag_refinement(&desc_, d_blocks_[cur_], cfg_.refine_type, above, below, d_tags_,
                               d_maxeps_, stream_), "apk_tag_refinement");
  std::vector<int> tags(nbl);
  cuda_check(cudaMemcpyAsync(tags.data(), d_tags_, nbl * sizeof(int), cudaMemcpyDeviceToHost,
                             S(stream_)), "tags D2H");
  Synchronize();
  refine_flag_.resize(nbl, 0);
  deref_count_.resize(nbl, 0);
  const int max_level = cfg_.numlevel - 1;  // the root grid is level 0 here
  for (size_t b = 0; b < nbl; ++b) {
    // MeshRefinement::SetRefinement(CheckAllRefinement(rc, AmrTag::derefine))
    const int aret = std::max(-1, tags[b]);
    if (aret == 0) refine_flag_[b] = 0;
    if (aret >= 0) deref_count_[b] = 0;
    if (aret > 0) {
      refine_flag_[b] = (blocks_[b].level == max_level) ? 0 : 1;
    } else if (aret < 0) {
      if (blocks_[b].level == 0) {
        refine_flag_[b] = 0;
        deref_count_[b] = 0;
      } else {
        deref_count_[b]++;
        int ec = 0;  // finer neighbours
        for (int idx = 0; idx < 27; ++idx)
          if (nbrinfo_[b].type[idx] == 3) ec += nbrinfo_[b].nchild[idx];
        if (ec > 0) refine_flag_[b] = 0;
        else refine_flag_[b] = (deref_count_[b] >= cfg_.derefine_count) ? -1 : 0;
      }
    }
  }
  amr_ms[0] += NowMs() - t_start;
  return TaskStatus::complete;
}

// ---- Mesh::UpdateMeshBlockTree (:485-604) -------------------------------------------------------
// From the leaves (level, lx1, lx2, lx3) and their refinement flags to the new leaf list in the
// order of Parthenon's block list.  Pure host code (no device state): also exported through
// apk_sim_amr_update_tree for the CPU tests.
int UpdateMeshBlockTree(const SimConfig &cfg, const std::vector<std::array<int, 4>> &old_leaves,
                        const std::vector<int> &flags, std::vector<std::array<int, 4>> *new_leaves,
                        int *nnew_out, int *ndel_out) {
  const int ndim = (cfg.nx[2] > 1) ? 3 : ((cfg.nx[1] > 1) ? 2 : 1);
  const int nold = (int)old_leaves.size();
  const int nleaf = 1 << ndim;
  *nnew_out = *ndel_out = 0;
  int tnref = 0, tnderef = 0;
  for (int b = 0; b < nold; ++b) {
    tnref += (flags[b] == 1);
    tnderef += (flags[b] == -1);
  }
  if (tnref == 0 && tnderef < nleaf) return 0;

  Forest F;
  F.ndim = ndim;
  int nb[3];
  for (int d = 0; d < 3; ++d) {
    nb[d] = cfg.nx[d] / cfg.mb[d];
    F.nb[d] = nb[d];
    F.periodic[d] = cfg.bc[2 * d] == APK_BC_PERIODIC;
  }
  auto loc_of = [&](int b) {
    return Loc{old_leaves[b][0], {old_leaves[b][1], old_leaves[b][2], old_leaves[b][3]}};
  };
  std::map<Loc, int> old_index;
  for (int b = 0; b < nold; ++b) {
    F.leaves.insert(loc_of(b));
    old_index[loc_of(b)] = b;
  }
  // position in Parthenon's block list: trees in Morton order of the tree location
  // (forest.cpp:108-144), the leaves of a tree in Morton order (logical_location.hpp:150-154)
  int mcp2 = 1 << 30;
  bool any_sym = false;
  for (int d = 0; d < 3; ++d) {
    if (cfg.nx[d] == 1) { any_sym = true; continue; }
    mcp2 = std::min(mcp2, nb[d] & (~(nb[d] - 1)));
  }
  if (any_sym) mcp2 = 1;
  const int finest = cfg.numlevel - 1;
  auto order_key = [&](const Loc &l) {
    int t[3] = {0, 0, 0}, in[3] = {0, 0, 0};
    for (int d = 0; d < ndim; ++d) {
      const int per_tree = mcp2 << l.level;
      t[d] = l.lx[d] / per_tree;
      in[d] = (l.lx[d] % per_tree) << (finest - l.level);
    }
    return std::make_tuple(Interleave(t), Interleave(in), l.level);
  };
  // newly derefined blocks: parents whose 2^ndim children are all flagged, in block-list order of
  // the first child (:560-585)
  std::vector<Loc> clderef;
  if (tnderef >= nleaf) {
    std::vector<std::pair<std::tuple<std::uint64_t, std::uint64_t, int>, Loc>> cand;
    for (int b = 0; b < nold; ++b) {
      if (flags[b] != -1) continue;
      const Loc l = loc_of(b);
      bool lower_left = l.level > 0;
      for (int d = 0; d < ndim; ++d) lower_left = lower_left && !(l.lx[d] & 1);
      if (!lower_left) continue;
      const Loc parent = Forest::Parent(l);
      bool all = true;
      for (int c = 1; c < nleaf && all; ++c) {
        auto it = old_index.find(F.Child(parent, c));
        all = it != old_index.end() && flags[it->second] == -1;
      }
      if (all) cand.push_back({order_key(l), parent});
    }
    std::sort(cand.begin(), cand.end(),
              [](const auto &a, const auto &b) { return a.first < b.first; });
    for (auto &c : cand) clderef.push_back(c.second);
    // "sort the lists by level": std::sort(&clderef[0], &clderef[ctnd - 1], level descending)
    // leaves the LAST element where it is (:587-592)
    if (clderef.size() > 1)
      std::stable_sort(clderef.begin(), clderef.end() - 1,
                       [](const Loc &a, const Loc &b) { return a.level > b.level; });
  }
  // Step 1: refinement, Step 2: derefinement (:596-603)
  int nnew = 0, ndel = 0;
  for (int b = 0; b < nold; ++b)
    if (flags[b] == 1) nnew += F.Refine(loc_of(b));
  for (const Loc &l : clderef) ndel += F.Derefine(l);
  *nnew_out = nnew;
  *ndel_out = ndel;
  if (nnew == 0 && ndel == 0) return 0;
  std::vector<std::pair<std::tuple<std::uint64_t, std::uint64_t, int>, Loc>> keyed;
  keyed.reserve(F.leaves.size());
  for (const Loc &l : F.leaves) keyed.push_back({order_key(l), l});
  std::sort(keyed.begin(), keyed.end(),
            [](const auto &a, const auto &b) { return a.first < b.first; });
  new_leaves->clear();
  for (const auto &kl : keyed)
    new_leaves->push_back({kl.second.level, kl.second.lx[0], kl.second.lx[1], kl.second.lx[2]});
  return 1;
}

// ---- Mesh::LoadBalancingAndAdaptiveMeshRefinement ---------------------------------------------
bool HydroSim::LoadBalancingAndAdaptiveMeshRefinement() {
  if (!cfg_.adaptive) return false;
  const int nold = (int)blocks_.size();
  const int nleaf = 1 << ndim_;
  refine_flag_.resize(nold, 0);
  deref_count_.resize(nold, 0);
  std::vector<std::array<int, 4>> old_leaves(nold), new_leaves;
  std::map<Loc, int> old_index;
  for (int b = 0; b < nold; ++b) {
    old_leaves[b] = {blocks_[b].level, blocks_[b].loc[0], blocks_[b].loc[1], blocks_[b].loc[2]};
    old_index[Loc{blocks_[b].level, {blocks_[b].loc[0], blocks_[b].loc[1], blocks_[b].loc[2]}}] = b;
  }
  int nnew = 0, ndel = 0;
  double t0 = NowMs();
  const int changed = UpdateMeshBlockTree(cfg_, old_leaves, refine_flag_, &new_leaves, &nnew, &ndel);
  nbnew += nnew;
  nbdel += ndel;
  amr_ms[1] += NowMs() - t0;
  if (!changed) return false;
  amr_remeshes++;
  t0 = NowMs();
  Forest F;  // (Child() only)
  F.ndim = ndim_;

  // ---- RedistributeAndRefineMeshBlocks (:626-995), one rank
  const std::vector<BlockInfo> old_blocks = blocks_;
  const std::vector<int> old_flag = refine_flag_, old_count = deref_count_;
  double *old_cons = d_cons_;
  d_cons_ = nullptr;
  Synchronize();
  ReleaseMeshArrays();
  ArenaBegin((double)new_leaves.size() / (double)nold);
  leaves_.clear();
  for (const auto &l : new_leaves) leaves_.push_back(Leaf{l[0], {l[1], l[2], l[3]}});
  BuildMesh();
  double t1 = NowMs();
  amr_ms[5] += t1 - t0;
  BuildNeighbors();
  amr_ms[6] += NowMs() - t1;
  t1 = NowMs();
  UploadBlocks();  // new (zeroed) slabs, block tables, the multilevel exchange plan
  amr_ms[7] += NowMs() - t1;
  amr_ms[2] += NowMs() - t0;
  t0 = NowMs();

  const int nnewb = (int)blocks_.size();
  const size_t per = BlockFieldElems(), cper = CoarseElems();
  refine_flag_.assign(nnewb, 0);
  deref_count_.assign(nnewb, 0);
  std::vector<apk_box_item_t> move, to_scratch, from_scratch;
  std::vector<apk_amr_item_t> prolong, restr;
  int cis[3], cie[3], is[3];
  for (int d = 0; d < 3; ++d) {
    cis[d] = (d < ndim_) ? cfg_.ng : 0;
    cie[d] = (d < ndim_) ? cfg_.ng + cfg_.mb[d] / 2 - 1 : 0;
    is[d] = (d < ndim_) ? cfg_.ng : 0;
  }
  // scratch coarse buffers: one per new fine block and one per removed fine block
  size_t nscratch = 0;
  for (int n = 0; n < nnewb; ++n) {
    const Loc l{blocks_[n].level, {blocks_[n].loc[0], blocks_[n].loc[1], blocks_[n].loc[2]}};
    if (old_index.count(l)) continue;
    nscratch += (l.level > 0 && old_index.count(Forest::Parent(l))) ? 1 : nleaf;
  }
  double *scratch =
      nscratch ? static_cast<double *>(SlabAlloc(nscratch * cper * sizeof(double), "remesh scratch"))
               : nullptr;
  size_t used = 0;
  auto amr_item = [&](double *fine, double *coarse, const BlockInfo &B) {
    apk_amr_item_t it{};
    it.fine = fine;
    it.coarse = coarse;
    for (int d = 0; d < 3; ++d) {
      it.nf[d] = n_[d];
      it.nc[d] = nc_[d];
      it.lo[d] = cis[d];
      it.hi[d] = cie[d];
      it.xmin[d] = B.bxmin[d];
      it.dx[d] = B.dx[d];
    }
    return it;
  };
  for (int n = 0; n < nnewb; ++n) {
    const BlockInfo &B = blocks_[n];
    const Loc l{B.level, {B.loc[0], B.loc[1], B.loc[2]}};
    double *dst = d_cons_ + (size_t)n * per;
    auto same = old_index.find(l);
    if (same != old_index.end()) {
      // same level: the block moves as it is, with its derefinement counter
      apk_box_item_t it{};
      it.src = old_cons + (size_t)same->second * per;
      it.dst = dst;
      for (int d = 0; d < 3; ++d) {
        it.ns[d] = it.nd[d] = it.ext[d] = n_[d];
        it.s0[d] = it.d0[d] = 0;
      }
      move.push_back(it);
      refine_flag_[n] = old_flag[same->second];
      deref_count_[n] = old_count[same->second];
      continue;
    }
    auto parent = (l.level > 0) ? old_index.find(Forest::Parent(l)) : old_index.end();
    if (parent != old_index.end()) {
      // coarse to fine (TryRecvCoarseToFine, :82-146): the ENTIRE coarse buffer is the octant of
      // the parent's array (ghost cells included); the interior is then prolongated
      double *cb = scratch + (used++) * cper;
      apk_box_item_t it{};
      it.src = old_cons + (size_t)parent->second * per;
      it.dst = cb;
      for (int d = 0; d < 3; ++d) {
        it.ns[d] = n_[d];
        it.nd[d] = it.ext[d] = nc_[d];
        it.s0[d] = (d < ndim_ && (l.lx[d] & 1)) ? cfg_.mb[d] / 2 : 0;
        it.d0[d] = 0;
      }
      to_scratch.push_back(it);
      prolong.push_back(amr_item(dst, cb, B));
    } else {
      // fine to coarse (:710-748 restriction of the children's interiors into their coarse buffers,
      // TryRecvFineToCoarse :163-228 copies them into the octants)
      for (int c = 0; c < nleaf; ++c) {
        auto child = old_index.find(F.Child(l, c));
        if (child == old_index.end()) throw std::runtime_error("apk_host: remesh lost a meshblock");
        double *cb = scratch + (used++) * cper;
        restr.push_back(amr_item(old_cons + (size_t)child->second * per, cb, old_blocks[child->second]));
        apk_box_item_t it{};
        it.src = cb;
        it.dst = dst;
        for (int d = 0; d < 3; ++d) {
          it.ns[d] = nc_[d];
          it.nd[d] = n_[d];
          it.s0[d] = cis[d];
          it.ext[d] = (d < ndim_) ? cfg_.mb[d] / 2 : 1;
          it.d0[d] = is[d] + ((d < ndim_ && ((c >> d) & 1)) ? cfg_.mb[d] / 2 : 0);
        }
        from_scratch.push_back(it);
      }
    }
  }
  // one upload for all five item lists
  const size_t nbox = move.size() + to_scratch.size() + from_scratch.size();
  const size_t namr = prolong.size() + restr.size();
  const size_t bytes = nbox * sizeof(apk_box_item_t) + namr * sizeof(apk_amr_item_t);
  std::vector<char> host(bytes);
  char *hp = host.data();
  auto put = [&](const void *src, size_t n) {
    const size_t off = hp - host.data();
    if (n) std::memcpy(hp, src, n);
    hp += n;
    return off;
  };
  const size_t o_move = put(move.data(), move.size() * sizeof(apk_box_item_t));
  const size_t o_to = put(to_scratch.data(), to_scratch.size() * sizeof(apk_box_item_t));
  const size_t o_from = put(from_scratch.data(), from_scratch.size() * sizeof(apk_box_item_t));
  const size_t o_pro = put(prolong.data(), prolong.size() * sizeof(apk_amr_item_t));
  const size_t o_res = put(restr.data(), restr.size() * sizeof(apk_amr_item_t));
  char *d_items = static_cast<char *>(SlabAlloc(bytes, "remesh items"));
  cuda_check(cudaMemcpyAsync(d_items, host.data(), bytes, cudaMemcpyHostToDevice, S(stream_)),
             "remesh items H2D");
  auto box = [&](size_t off, size_t n) {
    if (n)
      abi_check(apk_box_copy(reinterpret_cast<const apk_box_item_t *>(d_items + off), (int)n, nvar_,
                             stream_), "apk_box_copy (remesh)");
  };
  box(o_move, move.size());
  if (!restr.empty())
    abi_check(apk_restrict_average(ndim_, cfg_.ng, nvar_, 0,
                                   reinterpret_cast<const apk_amr_item_t *>(d_items + o_res),
                                   (int)restr.size(), cfg_.fp_mode, stream_),
              "apk_restrict_average (remesh)");
  box(o_from, from_scratch.size());
  box(o_to, to_scratch.size());
  if (!prolong.empty())
    abi_check(apk_prolongate_minmod_multid(ndim_, cfg_.ng, nvar_,
                                           reinterpret_cast<const apk_amr_item_t *>(d_items + o_pro),
                                           (int)prolong.size(), cfg_.fp_mode, stream_),
              "apk_prolongate_minmod_multid (remesh)");
  Synchronize();  // the host item buffer and the old slabs are released next
  SlabFree(d_items);
  SlabFree(scratch);
  SlabFree(old_cons);
  amr_ms[3] += NowMs() - t0;
  t0 = NowMs();
  // :973-990 ghost cells with real data, derived fields
  SendBoundBufs(0);
  ReceiveBoundBufs();
  SetBounds(0);
  ProlongateBounds();
  ApplyBoundaryConditions(0);
  FillDerived();
  amr_ms[4] += NowMs() - t0;
  return true;
}

}  // namespace apk_host
