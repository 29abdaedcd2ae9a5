/* apk_b200_sim.h -- flat C interface of the host-side driver mirror (athenapk_b200/host):
 * a HydroDriver-equivalent time loop (src/hydro/hydro_driver.cpp:347-673,
 * external/parthenon/src/driver/driver.cpp:67-242) built on the task-level C-ABI of
 * apk_b200.h.  Used by tests/bench through ctypes; a Parthenon integration does not need it
 * (it keeps its own driver and calls apk_b200.h directly). */
#ifndef APK_B200_SIM_H_
#define APK_B200_SIM_H_

#include "apk_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct apk_sim apk_sim_t;

/* input-deck values (docs/input.md): <parthenon/mesh>, <parthenon/meshblock>,
 * <parthenon/time>, <hydro> */
typedef struct apk_sim_config {
  int fluid, recon, riemann, integrator; /* integrator: 0 rk1, 1 rk2, 2 vl2, 3 rk3 */
  int nx[3], mb[3], ng, bc[6];
  double xmin[3], xmax[3];
  double gamma, cfl;
  double dfloor, pfloor, efloor, vceil, eceil;
  int fofc, dedner_extended;
  double glmmhd_alpha;
  int nscalars;
  int fp_mode; /* APK_FP_FAST / APK_FP_STRICT */
  int fused;   /* use the fused stage kernel when legal */
  int rank, nranks;
  int plan_only; /* build mesh partition + halo plan only: no CUDA calls (CPU-side tests) */
  /* static mesh refinement: <parthenon/static_refinementN> regions (level >= 1); nref = 0: uniform */
  int nref;
  int ref_level[4];
  double ref_xmin[4][3], ref_xmax[4][3];
  /* <diffusion> (hydro.cpp:539-702), fixed coefficients */
  int diffint;      /* diffusion/integrator: 0 none, 1 unsplit, 2 rkl2 */
  int conduction;   /* APK_CONDUCTION_* */
  int viscosity;    /* 1: isotropic */
  int resistivity;  /* 1: ohmic */
  double thermal_diff_coeff, mom_diff_coeff, ohm_diff_coeff;
  double cfl_diff;           /* diffusion/cfl; <= 0: parthenon/time/cfl */
  double rkl2_max_dt_ratio;  /* <= 0: no cap */
  double conduction_sat_phi; /* default 0.3 */
  /* adaptive mesh refinement (one rank): parthenon/mesh/refinement = adaptive with numlevel levels
   * in total and derefine_count (default 10, mesh_refinement.cpp:56); <refinement> type =
   * pressure_gradient with threshold_pressure_gradient (src/refinement/gradient.cpp:18-61) */
  int adaptive, numlevel, derefine_count;
  double refine_threshold;
  /* <refinement> type = APK_REFINE_* (apk_b200.h); refine_threshold is threshold_pressure_gradient /
   * threshold_xyvelocity_gradient / maxdensity_refine_above, refine_threshold2 maxdensity_deref_below */
  int refine_type;
  double refine_threshold2;
} apk_sim_config_t;

const char *apk_sim_last_error(void);
void apk_sim_default_config(apk_sim_config_t *c);
apk_sim_t *apk_sim_create(const apk_sim_config_t *c);
void apk_sim_destroy(apk_sim_t *s);
int apk_sim_nccl_unique_id(void *out, int nbytes);
int apk_sim_init_comm(apk_sim_t *s, const void *uid, int nbytes);
int apk_sim_pgen(apk_sim_t *s, const char *id, const double *pars, int npars);
int apk_sim_initialize(apk_sim_t *s, double tlim);
int apk_sim_execute(apk_sim_t *s, double tlim, int nlim); /* returns #cycles or <0 */
int apk_sim_step(apk_sim_t *s);
int apk_sim_num_local_blocks(const apk_sim_t *s);
int apk_sim_num_global_blocks(const apk_sim_t *s);
int apk_sim_nvar(const apk_sim_t *s);
void apk_sim_block_dims(const apk_sim_t *s, int *n3);
void apk_sim_block_info(const apk_sim_t *s, int b, int *gid, int *loc3, double *xmin3, double *dx3);
int apk_sim_block_level(const apk_sim_t *s, int b);
/* adaptive meshes (block counts and block order change whenever the mesh does): meshblocks created
 * / destroyed so far (Mesh::nbnew, nbdel; the refinement during initialisation counts) and blocks
 * per level; the two halves of the per-cycle AMR step for unit tests -- Refinement::Tag on every
 * block (amr_criteria/refinement_package.cpp:205-233 + mesh_refinement.cpp:72-110) and
 * Mesh::LoadBalancingAndAdaptiveMeshRefinement (mesh-amr_loadbalance.cpp:312-340; returns 1 when
 * the mesh changed: ghost cells and prim are rebuilt, the caller re-estimates dt) */
int apk_sim_amr_counts(const apk_sim_t *s, int *nbnew, int *nbdel, int *per_level8);
int apk_sim_amr_tag(apk_sim_t *s);
int apk_sim_amr_remesh(apk_sim_t *s);
int apk_sim_amr_flag(const apk_sim_t *s, int b);
/* host wall time (ms, summed) of the AMR step: 0 tagging, 1 tree update, 2 rebuilding mesh tables /
 * device arrays / exchange plan, 3 data move, 4 ghost cells + FillDerived, 5..7 = 2 split into
 * release + mesh, neighbour tables, device arrays + plan; returns #remeshes */
long long apk_sim_amr_timing(const apk_sim_t *s, double *ms8);
/* the tree update alone (no device state; CPU tests): leaves4 = nold x {level, lx1, lx2, lx3},
 * flags = refine_flag_ per leaf (-1 / 0 / 1); returns 0 when nothing changes, else the number of
 * new leaves written to out4 (capacity cap leaves) in the order of Parthenon's block list, or < 0 */
int apk_sim_amr_update_tree(const apk_sim_config_t *c, int nold, const int *leaves4,
                            const int *flags, int cap, int *out4, int *nnew, int *ndel);
int apk_sim_copy_to_host(apk_sim_t *s, int field, int block, double *dst);
int apk_sim_copy_from_host(apk_sim_t *s, int field, int block, const double *src);
/* whole-pack async copies on the sim stream (host buffer: [nblocks][nvar][n3][n2][n1]) */
int apk_sim_upload_all(apk_sim_t *s, int field, const double *src);
int apk_sim_download_all(apk_sim_t *s, int field, double *dst);
/* conserved state without ghost cells, dense [nblocks][nvar][nx3][nx2][nx1] (what a restart file
 * or an external producer holds).  load = H2D + scatter into the block interiors + ghost exchange
 * + physical BCs + FillDerived (the Mesh::Initialize sequence); store = gather + D2H.  Both are
 * enqueued on the simulation stream without host synchronisation (host buffers should be pinned). */
/* HydroHst volume sums of the local blocks (hydro.cpp:145-208): mass, mom1-3, KE, E, ME, relDivB */
int apk_sim_history(apk_sim_t *s, double *out8);
/* Parthenon's history file (`<parthenon/outputN> file_type = hst`, outputs/history.cpp:172-209):
 * apk_sim_execute appends a row at cycle 0 and whenever time >= next_time (every `dt`) while the
 * run keeps going, apk_sim_finalize_outputs writes the final row (the `final` signal of
 * driver.cpp:158-160).  Same header, columns and number format (data_format, default "%12.5e"). */
int apk_sim_enable_history(apk_sim_t *s, const char *fname, double dt, const char *data_format);
int apk_sim_finalize_outputs(apk_sim_t *s);
long long apk_sim_interior_elems(const apk_sim_t *s);
int apk_sim_load_interior(apk_sim_t *s, const double *src);
int apk_sim_store_interior(apk_sim_t *s, double *dst);
/* per-task device time (CUDA events on the launching stream), summed since enabling */
void apk_sim_enable_task_timing(apk_sim_t *s, int on);
int apk_sim_task_timing(const apk_sim_t *s, const char *name, double *ms_total, long long *count);
void *apk_sim_device_ptr(apk_sim_t *s, int field, int block);
double apk_sim_time(const apk_sim_t *s);
double apk_sim_dt(const apk_sim_t *s);
int apk_sim_ncycle(const apk_sim_t *s);
double apk_sim_c_h(const apk_sim_t *s);
long long apk_sim_fofc_corrections(const apk_sim_t *s);
/* RKL2 stages of the last super time step (diffusion/integrator = rkl2) */
int apk_sim_last_s_rkl(const apk_sim_t *s);
/* fused stages whose FOFC acceptance test failed somewhere and that were redone with the
 * flux-materialising tasks (0 for decks where FOFC never triggers) */
long long apk_sim_fofc_redone_stages(const apk_sim_t *s);
int apk_sim_fused_active(const apk_sim_t *s);
/* 1 when the remote halo exchange runs as peer stores inside the stage kernel (CUDA IPC over
 * NVLink), 0 when it uses the packed ncclSend/ncclRecv path */
int apk_sim_p2p_active(const apk_sim_t *s);
double apk_sim_linear_wave_period(const apk_sim_t *s);
/* out9[0] = RMS-L1, out9[1..] = per-variable L1 errors (5, or 8 for linear_wave_mhd) */
int apk_sim_linear_wave_errors(apk_sim_t *s, double *out9);
void *apk_sim_stream(apk_sim_t *s);
const void *apk_sim_pack_desc(const apk_sim_t *s);
const void *apk_sim_device_blocks(const apk_sim_t *s);
int apk_sim_synchronize(apk_sim_t *s);
int apk_sim_task(apk_sim_t *s, const char *name, int stage, double g0, double g1, double bdt);
void apk_sim_set_c_h(apk_sim_t *s, double c_h, double mindx);
/* halo plan introspection (valid in plan_only mode too): peers and per-peer slab lists.
 * counts/offsets are in doubles and sized for cons+prim (2*nvar values per cell). */
int apk_sim_num_peers(const apk_sim_t *s);
int apk_sim_peer(const apk_sim_t *s, int p, int *rank, long long *send_off, long long *send_cnt,
                 long long *recv_off, long long *recv_cnt);
int apk_sim_num_halo_items(const apk_sim_t *s, int recv);
int apk_sim_halo_item(const apk_sim_t *s, int recv, int idx, apk_halo_item_t *out);
/* ghost-push destination of (block, offset): owning rank and local block index there (the table
 * apk_stage_fused_p2p is fed with); -1 when the offset has no neighbour */
int apk_sim_push_target(const apk_sim_t *s, int block, int o1, int o2, int o3, int *rank,
                        int *local_index);
int apk_sim_neighbor(const apk_sim_t *s, int block, int o1, int o2, int o3);
/* refined meshes: items of the multilevel exchange / flux-correction plan this rank executes
 * (valid in plan_only mode too).  out10 = list id (0 restrict into the fine block's coarse buffer,
 * 1 fine -> coarse ghost cells, 2 restrict own cells, 3 coarser neighbour -> coarse buffer,
 * 4 prolongate, 5 same-level ghost slab pulled from another rank, 6+d / 9+d flux correction
 * restrict / replace), executing block, source block, destination block (global ids), region
 * start[3], extent[3] */
int apk_sim_ml_num_items(const apk_sim_t *s);
int apk_sim_ml_item(const apk_sim_t *s, int i, int *out10);

#ifdef __cplusplus
}
#endif
#endif
