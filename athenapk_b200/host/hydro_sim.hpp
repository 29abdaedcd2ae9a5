// hydro_sim.hpp -- host-side C++ mirror of the slice of AthenaPK/Parthenon that drives the hot
// path: HydroDriver::MakeTaskCollection's stage region (src/hydro/hydro_driver.cpp:474-613),
// the LowStorageIntegrator coefficients (time_integration/low_storage_integrator.cpp:53-132),
// EvolutionDriver's dt logic (driver/driver.cpp:67-165,210-242) and PreStepMeshUserWorkInLoop
// (src/hydro/hydro.cpp:102-143).  Same task names, argument meaning and TaskStatus contract;
// every device operation goes through the C-ABI of include/apk_b200.h -- this file contains no
// CUDA kernels and is what a Parthenon-side shim replaces one-for-one.
#ifndef APK_HOST_HYDRO_SIM_HPP_
#define APK_HOST_HYDRO_SIM_HPP_

#include <array>
#include <cstdint>
#include <functional>
#include <limits>
#include <map>
#include <string>
#include <vector>

#include "../../include/apk_b200.h"

struct ncclComm;

namespace apk_host {

using Real = double;

// parthenon basic_types.hpp:62
enum class TaskStatus { complete, incomplete, iterate, fail };
enum class Integrator { rk1 = 0, rk2 = 1, vl2 = 2, rk3 = 3 };

// what the input deck provides (docs/input.md; hydro.cpp:267-744)
struct SimConfig {
  int fluid = APK_FLUID_EULER;
  int recon = APK_RECON_PLM;
  int riemann = APK_RIEMANN_HLLE;
  int integrator = (int)Integrator::vl2;
  int nx[3] = {1, 1, 1};  // parthenon/mesh/nx*
  int mb[3] = {1, 1, 1};  // parthenon/meshblock/nx*
  int ng = 2;             // parthenon/mesh/nghost
  int bc[6] = {0, 0, 0, 0, 0, 0};  // ix1,ox1,... APK_BC_PERIODIC/OUTFLOW/REFLECT
  double xmin[3] = {-0.5, -0.5, -0.5}, xmax[3] = {0.5, 0.5, 0.5};
  double gamma = 5.0 / 3.0, cfl = 0.3;
  double dfloor = -1.0, pfloor = -1.0, efloor = -1.0;
  double vceil = std::numeric_limits<double>::infinity();
  double eceil = std::numeric_limits<double>::infinity();
  int fofc = 0;             // hydro/first_order_flux_correct
  int dedner_extended = 0;  // hydro/glmmhd_source = dedner_extended
  double glmmhd_alpha = 0.1;
  int nscalars = 0;
  int fp_mode = APK_FP_FAST;
  int fused = 1;  // use apk_stage_fused when legal (no FOFC / extended source)
  int rank = 0, nranks = 1;
  int plan_only = 0;  // mesh partition + halo plan only, no CUDA (CPU-side tests of the N>1 logic)
  // static mesh refinement: the <parthenon/static_refinementN> blocks of a deck (mesh.cpp:1100-1165)
  int nref = 0;
  int ref_level[4] = {0, 0, 0, 0};
  double ref_xmin[4][3] = {}, ref_xmax[4][3] = {};
  // <diffusion> block (hydro.cpp:539-702), fixed coefficients
  int diffint = 0;      // diffusion/integrator: 0 none, 1 unsplit, 2 rkl2
  int conduction = 0;   // APK_CONDUCTION_*
  int viscosity = 0;    // 1: isotropic
  int resistivity = 0;  // 1: ohmic
  double thermal_diff_coeff = 0.0, mom_diff_coeff = 0.0, ohm_diff_coeff = 0.0;
  double cfl_diff = -1.0;           // diffusion/cfl (default: the hyperbolic cfl, hydro.cpp:697)
  double rkl2_max_dt_ratio = -1.0;  // hydro.cpp:688
  double conduction_sat_phi = 0.3;  // hydro.cpp:563-564
  // adaptive mesh refinement: parthenon/mesh/refinement = adaptive, numlevel, derefine_count
  // (mesh_refinement.cpp:56); <refinement> type = pressure_gradient, threshold_pressure_gradient
  int adaptive = 0, numlevel = 1, derefine_count = 10;
  double refine_threshold = 0.0;
  // <refinement> type (hydro.cpp:788-816): APK_REFINE_PRESSURE_GRADIENT (refine_threshold =
  // threshold_pressure_gradient), _XYVELOCITY_GRADIENT (threshold_xyvelocity_gradient), _MAXDENSITY
  // (refine_threshold = maxdensity_refine_above, refine_threshold2 = maxdensity_deref_below)
  int refine_type = APK_REFINE_PRESSURE_GRADIENT;
  double refine_threshold2 = 0.0;
};

// time_integration/low_storage_integrator.cpp:53-132
struct LowStorageIntegrator {
  int nstages = 0;
  std::vector<Real> gam0, gam1, beta;
  Real dt = 0.0;
  explicit LowStorageIntegrator(Integrator which);
};

// parthenon SimTime (basic_types.hpp:200-245)
struct SimTime {
  Real time = 0.0, tlim = std::numeric_limits<Real>::max(), dt = std::numeric_limits<Real>::max();
  int ncycle = 0, nlim = -1;
  bool KeepGoing() const { return ((time < tlim) && (nlim < 0 || ncycle < nlim)); }
};

struct BlockInfo {
  int gid;        // position in the Morton-ordered global block list
  int level = 0;  // refinement level (0 = root grid)
  int loc[3];     // logical location (lx1,lx2,lx3) at that level
  double bxmin[3], bxmax[3];
  double dx[3], xmin[3];  // UniformCartesian: dx_, xmin_ (uniform_coordinates.hpp:32-36)
};

struct TaskList {  // minimal stand-in for parthenon::TaskList (tasks/tasks.hpp): in-order
  std::vector<std::pair<std::string, std::function<TaskStatus()>>> tasks;
  void AddTask(const std::string &name, std::function<TaskStatus()> f) {
    tasks.emplace_back(name, std::move(f));
  }
  TaskStatus Execute();  // re-polls `incomplete` tasks like Parthenon's task loop
};

// Mesh::UpdateMeshBlockTree (mesh-amr_loadbalance.cpp:485-604) on plain leaf lists
// {level, lx1, lx2, lx3}: returns 1 when the mesh changes, with the new leaves in the order of
// Parthenon's block list and the numbers of blocks created / destroyed (hydro_amr.cpp)
int UpdateMeshBlockTree(const SimConfig &cfg, const std::vector<std::array<int, 4>> &old_leaves,
                        const std::vector<int> &flags, std::vector<std::array<int, 4>> *new_leaves,
                        int *nnew, int *ndel);

class HydroSim {
 public:
  explicit HydroSim(const SimConfig &cfg);
  ~HydroSim();
  HydroSim(const HydroSim &) = delete;

  // ---- mesh ----
  int NumLocalBlocks() const { return (int)blocks_.size(); }
  int NumGlobalBlocks() const { return nblocks_global_; }
  const BlockInfo &Block(int b) const { return blocks_[b]; }
  const SimConfig &Config() const { return cfg_; }
  int nvar() const { return nvar_; }
  const int *dims() const { return n_; }  // n1,n2,n3 incl. ghosts
  size_t BlockFieldElems() const { return (size_t)nvar_ * n_[0] * n_[1] * n_[2]; }
  const apk_pack_desc_t &desc() const { return desc_; }
  const apk_block_t *d_blocks() const { return d_blocks_[cur_]; }

  // ---- data movement (host <-> device), field: 0 cons, 1 prim, 2 u1, 3..5 flux x1..x3 ----
  int CopyFromHost(int field, int block, const double *src);
  // whole-pack transfers (all local blocks, one contiguous copy; async on the sim stream)
  int UploadAll(int field, const double *src);
  int DownloadAll(int field, double *dst);
  // per-task device timing (CUDA events on the launching stream)
  void EnableTaskTiming(bool on) { timing_on_ = on; }
  bool TaskTiming(const std::string &name, double *ms_total, long long *count) const;
  void ResetTaskTiming() { timing_.clear(); }
  int CopyToHost(int field, int block, double *dst);
  double *DevicePtr(int field, int block);

  // ---- problem generators (input side; src/pgen/*.cpp) ----
  int ProblemGenerator(const std::string &id, const std::vector<double> &pars);
  double LinearWavePeriod() const { return lw_.tlim_period; }
  // linear_wave.cpp:183-295 / linear_wave_mhd.cpp:183-330 (local blocks only): out[0] = RMS-L1,
  // out[1..] = per-variable L1 (5 values, or 8 for the magnetised wave); out has 9 entries
  void LinearWaveErrors(double *out9);

  // ---- driver (driver.cpp / multistage.hpp / hydro_driver.cpp) ----
  void Initialize(double tlim);  // Mesh::Initialize + InitializeBlockTimeSteps + SetGlobalTimeStep
  TaskStatus Step();             // MultiStageDriver::Step: all stages of one cycle
  int Execute(double tlim, int nlim);  // EvolutionDriver::Execute main loop; returns #cycles
  void MakeTaskCollection(TaskList &tl, int stage);

  // individual tasks (names follow the reference)
  TaskStatus InitU1();
  TaskStatus CalculateFluxes(int stage);
  TaskStatus FirstOrderFluxCorrect(Real gam0, Real gam1, Real beta_dt);
  TaskStatus UpdateWithFluxDivergence(Real gam0, Real gam1, Real beta_dt);
  TaskStatus AddUnsplitSources(Real beta_dt);
  TaskStatus FusedStage(int stage, Real gam0, Real gam1, Real beta_dt, bool last_stage);
  TaskStatus FusedFofcCheck(int stage, Real gam0, Real gam1, Real beta_dt, bool last_stage);
  TaskStatus StartReceiveBoundBufs();
  // with_prim: 0 cons, 1 cons+prim, 2 prim only; local_done: same-device ghosts already pushed
  TaskStatus SendBoundBufs(int with_prim, bool local_done = false, bool need_barrier = true);
  TaskStatus ReceiveBoundBufs();
  TaskStatus SetBounds(int with_prim);
  TaskStatus ApplyBoundaryConditions(int with_prim);
  // multilevel meshes: restriction into / prolongation from the coarse buffers (ProlongateBounds,
  // boundary_communication.cpp:361-393) and flux correction (hydro_driver.cpp:527-531)
  TaskStatus ProlongateBounds();
  TaskStatus FluxCorrection();
  bool Multilevel() const { return multilevel_; }
  // adaptive refinement (one rank): Refinement::Tag + MeshRefinement::SetRefinement on every block
  // (hydro_driver.cpp:662-670), and Mesh::LoadBalancingAndAdaptiveMeshRefinement followed by
  // EvolutionDriver::InitializeBlockTimeSteps when the mesh changed (driver.cpp:128-131);
  // returns true when blocks were created or destroyed
  TaskStatus TagRefinement();
  bool LoadBalancingAndAdaptiveMeshRefinement();
  int nbnew = 0, nbdel = 0;  // Mesh::nbnew, nbdel
  // host wall time spent so far in: 0 tagging, 1 tree update, 2 rebuilding mesh tables / device
  // arrays / exchange plan, 3 data move (copy, prolongate, restrict), 4 ghost cells + FillDerived
  // (5..7 split 2 into: release + mesh, neighbour tables, device arrays + exchange plan)
  double amr_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long amr_remeshes = 0;
  int RefineFlag(int b) const { return refine_flag_[b]; }
  TaskStatus FillDerived();
  // diffusion (src/hydro/diffusion/*, hydro_driver.cpp:36-344)
  TaskStatus CalcDiffFluxes();
  TaskStatus SuperTimeStep(Real tau);  // AddSTSTasks: one RKL2 super step of length tau
  Real EstimateDiffusionTimestep();    // min of the Estimate*Timestep functions (hydro.cpp:935-946)
  int last_s_rkl = 0;
  TaskStatus EstimateTimestep(bool have_device_min);
  void PreStepMeshUserWorkInLoop();
  void SetGlobalTimeStep();

  // multi-GPU: one NCCL communicator over the ranks of the box
  int InitComm(const void *nccl_unique_id, int nbytes);
  static int CommUniqueId(void *out, int nbytes);

  SimTime tm;
  Real c_h = 0.0, mindx = std::numeric_limits<Real>::max(),
       dt_hyp = std::numeric_limits<Real>::max();
  long long fofc_corrections = 0;
  long long fofc_redone_stages = 0;  // fused stages rolled back and redone with the FOFC tasks
  long long cycles_done = 0;
  bool FusedActive() const { return fused_; }
  bool P2PActive() const { return p2p_; }
  // Host <-> device exchange of the conserved state WITHOUT ghost cells (what a restart file or
  // an external producer holds): dense [nblocks][nvar][nx3][nx2][nx1].  Load = H2D + scatter into
  // the block interiors + ghost exchange + physical BCs + FillDerived; Store = gather + D2H.
  // Everything is enqueued on the simulation stream (no host synchronisation).
  // HydroHst volume sums of the local blocks (hydro.cpp:145-208): mass, mom1-3, KE, E, ME, relDivB
  void History(double *out8);
  // Parthenon's history output (`file_type = hst`, outputs/history.cpp:172-209; scheduling
  // outputs.cpp:445-485 and driver.cpp:95-165): Execute() appends a row at cycle 0, then whenever
  // time >= next_time while the run keeps going; FinalizeOutputs() is the `final` signal after the
  // loop.  Same header, columns ("mass", "1-mom", ... hydro.cpp:422-441) and number format.
  void EnableHistoryOutput(const std::string &fname, double dt, const std::string &data_format);
  void FinalizeOutputs();
  size_t InteriorElems() const;
  int LoadStateInterior(const double *h_src);
  int StoreStateInterior(double *h_dst);
  // halo plan introspection
  struct PeerInfo { int rank; long long send_off, send_cnt, recv_off, recv_cnt; };
  int NumPeers() const { return (int)peers_.size(); }
  PeerInfo GetPeer(int p) const {
    return {peers_[p].rank, peers_[p].send_off, peers_[p].send_cnt, peers_[p].recv_off, peers_[p].recv_cnt};
  }
  const std::vector<apk_halo_item_t> &HaloItems(bool recv) const { return recv ? recv_items_ : send_items_; }
  int Neighbor(int block, int o1, int o2, int o3) const {
    return h_blocks_[block].nbr[(o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1)];
  }
  void Synchronize();
  void *stream() const { return stream_; }

 private:
  void BuildMesh();
  void ReleaseMeshArrays();  // everything UploadBlocks allocates per mesh, except the cons slab
  // Per-mesh device arrays.  Static meshes: one cudaMalloc each.  Adaptive meshes: bump allocation
  // from two grow-only arenas used alternately (the old cons slab must outlive the construction of
  // the new mesh), so a remesh costs no cudaMalloc / cudaFree once the arenas have reached their
  // size; requests that do not fit fall back to cudaMalloc until the arena is regrown at its next turn
  struct Arena {
    char *base = nullptr;
    size_t cap = 0, used = 0, want = 0;
    std::vector<void *> extra;
  };
  Arena arena_[2];
  int arena_cur_ = 0;
  void ArenaBegin(double growth);  // switch arenas before the arrays of a new mesh are allocated
  void *SlabAlloc(size_t bytes, const char *what);
  void SlabFree(void *p);
  std::vector<int> refine_flag_, deref_count_;  // MeshRefinement::refine_flag_, deref_count_
  int *d_tags_ = nullptr;
  double *d_maxeps_ = nullptr;
  size_t tags_cap_ = 0;
  std::string hst_name_, hst_fmt_ = " %12.5e";
  double hst_dt_ = -1.0, hst_next_time_ = 0.0;
  int hst_file_number_ = 0, hst_last_cycle_ = -1;
  void MakeOutputs(bool final_signal);
  void WriteHistoryRow();
  std::string pgen_id_;
  std::vector<double> pgen_pars_;
  void BuildNeighbors();
  void BuildHaloPlan();
  void UploadBlocks();

  SimConfig cfg_;
  LowStorageIntegrator integrator_;
  int ndim_ = 1, nvar_ = 5, nhydro_ = 5;
  int nb_[3] = {1, 1, 1}, n_[3] = {1, 1, 1};
  int nblocks_global_ = 1;
  std::vector<int> gid_of_loc_;        // lexicographic loc -> gid (Morton order)
  std::vector<std::array<int, 3>> loc_of_gid_;
  int gid_lo_ = 0, gid_hi_ = 1;        // this rank owns gids [lo,hi)
  std::vector<BlockInfo> blocks_;
  std::vector<apk_block_t> h_blocks_;
  apk_pack_desc_t desc_{};
  unsigned bc_face_mask_ = 0;
  bool fused_ = false;
  int cur_ = 0;  // which prim buffer is the current "prim" (fused ping-pong)

  // device storage (one slab per field for all local blocks)
  double *d_cons_ = nullptr, *d_prim_[2] = {nullptr, nullptr}, *d_u1_ = nullptr,
         *d_flux_[3] = {nullptr, nullptr, nullptr};
  apk_block_t *d_blocks_[2] = {nullptr, nullptr};
  double **d_prim_out_[2] = {nullptr, nullptr};
  double *d_scalars_ = nullptr;  // [0] dt-min, [1] status (as double), [8] DBL_MAX constant
  double *h_scalars_ = nullptr;  // pinned host mirror of d_scalars_[0..1] (one async D2H per cycle)
  int *d_status_ = nullptr;
  void *d_fofc_work_ = nullptr;
  // ---- static mesh refinement ----
  bool multilevel_ = false;
  struct Leaf { int level; int lx[3]; };
  std::vector<Leaf> BuildLeaves() const;
  // neighbour kind per (block, offset): 0 none, 1 same level, 2 coarser, 3 finer
  struct NbrInfo { int type[27]; int idx[27][4]; int nchild[27]; bool has_coarser; };
  std::vector<NbrInfo> nbrinfo_;  // per LOCAL block; idx[][] are GLOBAL block ids
  std::vector<NbrInfo> gnbr_;     // per global block (every rank builds the whole topology)
  std::vector<Leaf> leaves_;      // global leaf list (the partition cuts it into contiguous ranges)
  Real global_mindx_ = 0.0;       // min dx over ALL leaves (CalculateGlobalMinDx + all-reduce)
  int RankOfGid(int gid) const;
  int GidLo(int rank) const;
  bool IsLocal(int gid) const { return gid >= gid_lo_ && gid < gid_hi_; }
  BlockInfo MakeBlockInfo(int gid, int level, const int *loc) const;
  void BuildMultilevelNeighbors();
  void BuildMultilevelPlan();
  apk_block_t *d_cblocks_ = nullptr;  // block table over the coarse buffers (physical BCs on them)
  apk_pack_desc_t desc_coarse_{};
  double *d_coarse_ = nullptr;  // coarse buffers [nblocks][nvar][nc3][nc2][nc1] (Variable::coarse_s)
  int nc_[3] = {1, 1, 1};
  size_t CoarseElems() const { return (size_t)nvar_ * nc_[0] * nc_[1] * nc_[2]; }
  // item lists of the multilevel exchange (device copies), one launch each
  struct MlList { void *d = nullptr; int n = 0; };
  MlList ml_restrict_child_, ml_child_to_ghost_, ml_restrict_own_, ml_coarse_from_nbr_, ml_prolong_;
  MlList ml_fc_restrict_[3], ml_fc_copy_[3];
  MlList ml_same_remote_;  // same-level neighbours on other ranks: ghost slabs pulled from peer memory
 public:
  // host-side description of every item of the plan (CPU tests of the multi-rank partition):
  // list id (0 restrict_child, 1 child_to_ghost, 2 restrict_own, 3 coarse_from_nbr, 4 prolong,
  // 5 same_remote, 6 + d flux restrict, 9 + d flux copy), executing block, source block,
  // destination block (global ids), region start and extent
  struct MlDesc { int list, exec, src, dst, lo[3], ext[3]; };
  std::vector<MlDesc> ml_desc_;
 private:
  std::vector<void *> ml_peer_;  // [rank][2]: IPC-mapped cons / coarse slabs of the other ranks
  void *d_tma_[2] = {nullptr, nullptr};  // per prim parity: CUtensorMap per block (A/B variant)
  double *d_undo_ = nullptr;         // copy of a stage's input cons (fused FOFC, rk2 / rk3)
  bool last_stage_unfused_ = false;  // the cycle's last stage was redone un-fused: no device dt min
  void *stream_ = nullptr;

  // halo plan
  struct Peer {
    int rank;
    long long send_off, send_cnt, recv_off, recv_cnt;  // in doubles (for with_prim = 0)
  };
  std::vector<Peer> peers_;
  std::vector<apk_halo_item_t> send_items_, recv_items_;
  long long send_total_ = 0, recv_total_ = 0;
  ncclComm *comm_ = nullptr;
  apk_halo_plan_t *halo_plan_ = nullptr;  // buffers, comm stream and events live behind the C-ABI
  bool recv_posted_ = false;
  // peer-to-peer ghost push (fused stages, nranks > 1): the stage kernel stores boundary cells
  // straight into the neighbour GPU's ghost zones through CUDA-IPC mapped pointers
  bool p2p_ = false;            // all ranks mapped their neighbours' field buffers
  bool nccl_inflight_ = false;  // the last SendBoundBufs used the packed NCCL path
  apk_push_dst_t *d_push_[2] = {nullptr, nullptr};  // [nblocks][27] per prim ping-pong parity
  int *d_barrier_ = nullptr;
  double *d_hist_ = nullptr;                // history-reduction workspace
  double *d_stage_ = nullptr;               // interior-only staging buffer
  apk_halo_item_t *d_int_items_ = nullptr;  // one slab per block: its interior
  void EnsureStage();
  std::vector<void *> peer_mapped_;
  void SetupPeerPush();
 public:
  // ghost-push destination of (block, neighbour offset index): owning rank and the block's local
  // index there; returns false when there is no neighbour (physical boundary / unused dimension)
  bool PushTarget(int block, int idx, int *rank, int *local_index) const;
 private:
  void StageBarrier();

  // linear wave pgen state (src/pgen/linear_wave.cpp:43-51)
  struct {
    bool set = false;
    int wave_flag = 0;
    double amp = 0, vflow = 0, sin_a2 = 0, cos_a2 = 1, sin_a3 = 0, cos_a3 = 1, lambda = 1, k_par = 1,
           d0 = 1, p0 = 1, u0 = 0, gm1 = 1, tlim_period = 1;
    double ev[5], rem[5][5];
    // magnetised wave (src/pgen/linear_wave_mhd.cpp:40-48)
    bool mhd = false;
    double bx0 = 0, by0 = 0, bz0 = 0, dby = 0, dbz = 0;
    double ev7[7], rem7[7][7];
  } lw_;
  void LinearWaveSetup(const std::vector<double> &pars, bool mhd_wave);
  void LinearWaveState(const BlockInfo &B, int k, int j, int i, double u[5]) const;
  void LinearWaveMhdState(const BlockInfo &B, int k, int j, int i, double u[8]) const;
  double LinearWaveMhdA(int c, double x1, double x2, double x3) const;
  Real new_dt_ = std::numeric_limits<Real>::max();
  Real dt_diff_ = std::numeric_limits<Real>::max();  // "dt_diff" Param (hydro.cpp:700,961-962)
  apk_diffusion_t diff_{};
  double *d_my0_ = nullptr, *d_yjm2_ = nullptr;      // RKL2 registers "MY0", "Yjm2"
  double **d_my0_ptrs_ = nullptr, **d_yjm2_ptrs_ = nullptr;
  struct Timing { double ms = 0.0; long long count = 0; };
  std::map<std::string, Timing> timing_;
  bool timing_on_ = false;
  TaskStatus ExecuteTimed(TaskList &tl);
};

}  // namespace apk_host
#endif
