/* apk_b200.h -- C-ABI of the B200-native AthenaPK hot path (libapk_b200.so).
 *
 * Drop-in boundary: every entry point below replaces one task/kernel that AthenaPK's
 * HydroDriver schedules per stage (hydro_driver.cpp:499-613 of the reference).  Plain
 * pointers and sizes only: no C++/torch/Kokkos types cross this boundary.  All work is
 * stream-ordered on the cudaStream_t passed in (as void*, 0 = default stream), no hidden
 * device synchronisation except where the reference itself blocks on a reduction
 * (apk_min_dt, apk_first_order_flux_correct).  Return value: 0 = ok, non-zero = error
 * (apk_last_error() gives the text); the Parthenon-side shim maps non-zero to
 * PARTHENON_FAIL.
 *
 * Ownership: the caller (Parthenon) owns every field array; the library never allocates,
 * frees or retains field memory.  apk_block_t arrays live in DEVICE memory.
 *
 * file:line cites are relative to the reference tree (parthenon-hpc-lab/athenapk).
 */
#ifndef APK_B200_H_
#define APK_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APK_ABI_VERSION 5  /* 5: apk_tag_refinement (all <refinement> criteria); 4: diffusion block */

/* src/main.hpp:19-50 (enum values are ours; the shim translates) */
enum { APK_FLUID_EULER = 0, APK_FLUID_GLMMHD = 1 };
enum { APK_RECON_DC = 0, APK_RECON_PLM = 1, APK_RECON_PPM = 2, APK_RECON_WENOZ = 3,
       APK_RECON_WENO3 = 4, APK_RECON_LIMO3 = 5 };
enum { APK_RIEMANN_NONE = 0, APK_RIEMANN_HLLE = 1, APK_RIEMANN_LLF = 2, APK_RIEMANN_HLLC = 3,
       APK_RIEMANN_HLLD = 4 };
enum { APK_BC_PERIODIC = 0, APK_BC_OUTFLOW = 1, APK_BC_REFLECT = 2, APK_BC_BLOCK = 3 };
/* floating-point contract of the kernels */
enum { APK_FP_FAST = 0,   /* FMA contraction on (what nvcc/Kokkos-CUDA does by default) */
       APK_FP_STRICT = 1  /* -fmad=false: bit-identical to the reference's CPU (serial) build */
};

/* neighbour table entry kinds for apk_block_t.nbr[] (27 entries, index (o3+1)*9+(o2+1)*3+(o1+1)) */
#define APK_NBR_NONE (-1)        /* no same-level neighbour: physical boundary or unused dim */
#define APK_NBR_REMOTE(slot) (-2 - (slot)) /* neighbour lives on another GPU: halo slot index */

/* One meshblock.  Field arrays are C-order [nvar][n3][n2][n1] with n* = nx* + 2*ng (1 in a
 * symmetry direction), exactly Parthenon's ParArrayND layout (variable_pack.hpp).
 * flux[d] has the same extents; entry (n,k,j,i) is the flux through the LOWER d-face of cell
 * (k,j,i) (hydro.cpp:1025-1208, metadata.cpp:174-205). */
typedef struct apk_block {
  double *cons;      /* "cons" of register base (u0)                       */
  double *prim;      /* "prim" of register base                            */
  double *cons_u1;   /* "cons" of register u1 (hydro_driver.cpp:474-495)   */
  double *flux[3];   /* bnd_flux::cons; may be NULL when only fused stages are used */
  double dx[3];      /* coords.Dxc<d>()  (uniform_coordinates.hpp:32-40)   */
  double xmin[3];    /* coords.Xf<d>(0): position of the first (ghost) face */
  int nbr[27];       /* same-level neighbour block index on THIS device, APK_NBR_NONE or
                        APK_NBR_REMOTE(slot)                               */
  int bc[6];         /* per-face flag ix1,ox1,ix2,ox2,ix3,ox3: APK_BC_* (APK_BC_BLOCK = interior
                        or periodic face handled by the neighbour table)   */
} apk_block_t;

/* One MeshBlockPack (mesh/meshblock_pack.hpp:35-91): what the kernels need to know */
typedef struct apk_pack_desc {
  int nblocks;
  int nvar;      /* nhydro + nscalars (5 / 9 + nscalars)                   */
  int nscalars;  /* hydro/nscalars, hydro.cpp:741-742                      */
  int nx[3];     /* interior cells per block                               */
  int ng;        /* parthenon/mesh/nghost                                  */
  int ndim;
  int fluid;     /* APK_FLUID_*                                            */
  int recon;     /* APK_RECON_*   (hydro/reconstruction)                   */
  int riemann;   /* APK_RIEMANN_* (hydro/riemann)                          */
  int fp_mode;   /* APK_FP_*                                               */
  double gamma;  /* hydro/gamma                                            */
  double dfloor, pfloor, efloor; /* <= 0 disables (hydro.cpp:507-523)      */
  double vceil, eceil;           /* +inf disables (hydro.cpp:525-537)      */
} apk_pack_desc_t;

const char *apk_last_error(void);
int apk_abi_version(void);
/* number of kernel launches issued by this library since load / since reset (bench evidence) */
long long apk_launch_count(void);
void apk_reset_launch_count(void);

/* ---- Hydro::CalculateFluxes<fluid,recon,riemann> (src/hydro/hydro.cpp:1025-1208, FluxFun_t
 * src/hydro/hydro.hpp:43-46) and CalculateFluxesTight (DC+LLF, hydro.cpp:980-1022).
 * Reads prim (ghosts valid), writes flux[0..ndim-1] over the reference's loop ranges.
 * recon_override >= 0 replaces desc->recon (VL2 predictor uses DC, hydro.cpp:459-464). */
int apk_calc_fluxes(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, double c_h,
                    int recon_override, void *stream);

/* ---- Hydro::FirstOrderFluxCorrect<fluid> (src/hydro/hydro.cpp:1223-1342).
 * Jacobi variant (all cells of one attempt test the fluxes of the previous attempt; the
 * reference's serial build is Gauss-Seidel in (b,k,j,i) order and its CUDA build is racy, see
 * DESIGN.md).  Blocks the host per attempt like the reference.  d_work: device scratch of
 * apk_fofc_workspace_bytes().  n_corrected (host) receives the total number of corrections. */
size_t apk_fofc_workspace_bytes(const apk_pack_desc_t *desc);
int apk_first_order_flux_correct(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                                 double gam0, double gam1, double beta_dt, double c_h,
                                 void *d_work, long long *n_corrected, void *stream);

/* ---- parthenon::Update::UpdateWithFluxDivergence<MeshData> (interface/update.cpp:115-141,
 * FluxDivHelper update.hpp:42-56): u0 = gam0*u0 + gam1*u1 + beta_dt*(-div F) on interior cells */
int apk_flux_div_update(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, double gam0,
                        double gam1, double beta_dt, void *stream);

/* ---- GLMMHD::DednerSource<extended> (src/hydro/glmmhd/dedner_source.cpp:17-75); the damping
 * coefficient exp(-alpha*c_h*beta_dt/mindx) is evaluated on the host like the reference */
int apk_dedner_source(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, double c_h,
                      double mindx, double alpha, double beta_dt, int extended, void *stream);

/* ---- EOS::ConservedToPrimitive (src/eos/adiabatic_{hydro,glmmhd}.cpp:33-56; cell kernels
 * adiabatic_hydro.hpp:51-142, adiabatic_glmmhd.hpp:61-167) over the ENTIRE index range.
 * d_status (device int, may be NULL): set to 1/2 on negative density/pressure without floor
 * (the reference aborts via PARTHENON_REQUIRE). */
int apk_cons_to_prim(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, int *d_status,
                     void *stream);

/* ---- Hydro::EstimateHyperbolicTimestep<fluid> (src/hydro/hydro.cpp:828-910): returns
 * cfl * min over interior cells/dirs of dx/(|v|+c) in *h_dt (blocking, like the reference's
 * parallel_reduce to a host value).  d_work: device scratch of >= 8 bytes. */
int apk_min_dt(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, double cfl,
               void *d_work, double *h_dt, void *stream);
/* non-blocking variant: leaves min dx/(|v|+c) (without cfl) in d_out[0] (device) */
int apk_min_dt_async(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, double *d_out,
                     void *stream);

/* ---- <diffusion> block (src/hydro/diffusion/, docs/input.md "diffusion"): fixed coefficients.
 * apk_diffusion_t carries what hydro.cpp:539-702 puts into the Hydro package Params. */
#define APK_CONDUCTION_NONE 0
#define APK_CONDUCTION_ISOTROPIC 1   /* diffusion/conduction = isotropic, conduction_coeff = fixed */
#define APK_CONDUCTION_ANISOTROPIC 2 /* ... = anisotropic (field aligned, saturated; needs glmmhd) */
typedef struct apk_diffusion {
  int conduction;  /* APK_CONDUCTION_* */
  int viscosity;   /* 1: diffusion/viscosity = isotropic, viscosity_coeff = fixed */
  int resistivity; /* 1: diffusion/resistivity = ohmic, resistivity_coeff = fixed */
  double thermal_diff_coeff, mom_diff_coeff, ohm_diff_coeff; /* *_diff_coeff_code */
  double conduction_sat_prefac; /* 5 * diffusion/conduction_sat_phi (hydro.cpp:604) */
} apk_diffusion_t;

/* ResetFluxes (src/hydro/hydro_driver.cpp:36-91): zero flux[d] over the faces the flux divergence
 * reads (the RKL2 sub-steps start from pure diffusive fluxes) */
int apk_reset_fluxes(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, void *stream);
/* CalcDiffFluxes (src/hydro/diffusion/diffusion.cpp:19-53): ADDS the conductive / viscous / Ohmic
 * fluxes computed from prim (ghosts valid) to flux[0..ndim-1] -- conduction.cpp:193-260 (isotropic)
 * and :265-550 (anisotropic + saturation), viscosity.cpp:98-291, resistivity.cpp:89-232.  With
 * diffusion/integrator = unsplit the reference calls it at the end of CalculateFluxes
 * (hydro.cpp:1202-1205); with rkl2 inside every sub-step (hydro_driver.cpp:246, :312). */
int apk_calc_diff_fluxes(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                         const apk_diffusion_t *diff, void *stream);
/* First RKL2 sub-step (hydro_driver.cpp:196-208 Y0 <- base into cons_u1, :255
 * Update::FluxDivergence into MY0, RKL2StepFirst :93-126): base = Y0 + mu~_1 tau MY0, Yjm2 = Y0.
 * d_my0 / d_yjm2: device arrays of per-block pointers to the "MY0" / "Yjm2" registers
 * (same layout as cons). */
int apk_rkl2_step_first(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                        double *const *d_my0, double *const *d_yjm2, int s_rkl, double tau,
                        void *stream);
/* RKL2StepOther (hydro_driver.cpp:128-166) for sub-step j: reads the diffusive fluxes of base */
int apk_rkl2_step_other(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                        double *const *d_my0, double *const *d_yjm2, double mu_j, double nu_j,
                        double mu_tilde_j, double gamma_tilde_j, double tau, void *stream);
/* EstimateConductionTimestep, general branch (conduction.cpp:91-186, anisotropic conduction): leaves
 * the minimum over interior cells (before cfl_diff * fac) in d_out[0]; DBL_MAX if no cell
 * constrains.  The fixed-coefficient estimates (conduction.cpp:64-89, viscosity.cpp:55-88,
 * resistivity.cpp:52-80) depend on dx only and are evaluated by the caller. */
int apk_min_dt_conduction(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                          const apk_diffusion_t *diff, double *d_out, void *stream);

/* ---- same-device part of parthenon::AddBoundaryExchangeTasks (bvals/comms/
 * boundary_communication.cpp:48-348, bnd_info.cpp:105-230) + physical boundary conditions
 * (bvals/boundary_conditions.cpp:36-59, boundary_conditions_generic.hpp:174-262,
 * src/bvals/boundary_conditions_apk.hpp:38-85): fills the ghost cells of `cons` straight
 * from the neighbour block's interior (no pack->buffer->unpack round trip), then applies
 * outflow/reflecting BCs in face order (= apk_copy_ghosts_local + apk_apply_physical_bcs).
 * with_prim (here and in the halo/BC entry points): 0 = cons only, 1 = cons and prim, 2 = prim
 * only (between fused stages nothing reads the ghost cells of cons).  with_prim != 0 additionally copies `prim` (valid
 * because cons->prim is pointwise) so no ghost cons->prim pass is needed. */
int apk_fill_ghosts_local(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                          unsigned bc_face_mask, int with_prim, void *stream);
/* physical BCs only (after apk_halo_unpack of remote neighbours); bit f of bc_face_mask set
 * = some block of the pack has an outflow/reflecting BC on face f (ix1,ox1,ix2,ox2,ix3,ox3) */
int apk_apply_physical_bcs(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                           unsigned bc_face_mask, int with_prim, void *stream);
/* neighbour copies only (no BCs) */
int apk_copy_ghosts_local(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                          int with_prim, void *stream);

/* ---- remote part of the exchange: pack all outgoing (block,neighbour) slabs of one peer
 * into ONE contiguous send buffer / unpack one contiguous receive buffer (replaces
 * SendBoundBufs/SetBounds per-buffer kernels + per-buffer MPI messages,
 * boundary_communication.cpp:95-140,273-334; utils/communication_buffer.hpp:209-236).
 * d_items: device array of apk_halo_item_t describing each slab. */
typedef struct apk_halo_item {
  int block;      /* local block index                                  */
  int lo[3];      /* first cell (i,j,k) of the slab in the block array   */
  int ext[3];     /* slab extents (i,j,k)                                */
  long long offset; /* offset (in doubles) of this slab in the buffer, sized for cons+prim
                       (2*nvar values per cell); single-field calls use offset/2 */
} apk_halo_item_t;
int apk_halo_pack(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                  const apk_halo_item_t *d_items, int nitems, int with_prim, double *d_buf,
                  void *stream);
int apk_halo_unpack(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                    const apk_halo_item_t *d_items, int nitems, int with_prim,
                    const double *d_buf, void *stream);

/* ---- the multi-rank exchange as three calls that map 1:1 onto the tasks of
 * parthenon::AddBoundaryExchangeTasks (bvals/comms/boundary_communication.cpp:405-445):
 *   apk_halo_start  = StartReceiveBoundBufs + SendBoundBufs (:48-140,:178-199): one pack launch for
 *                     the slabs of ALL peers, then one grouped ncclSend/ncclRecv pair per peer on the
 *                     plan's own communication stream (so the transfer overlaps whatever the caller
 *                     enqueues next on `stream`, e.g. the same-device ghost fill);
 *   apk_halo_test   = ReceiveBoundBufs (:201-235): 0 = complete, 1 = incomplete (cudaEventQuery of
 *                     the transfer) -- the polling contract of the reference's task (TaskStatus::
 *                     incomplete re-enqueues the task); a caller may skip it, apk_halo_finish orders
 *                     the unpack behind the transfer on the device anyway;
 *   apk_halo_finish = SetBounds (:273-348): stream-wait on the transfer, one unpack launch.
 * The plan owns the send/receive buffers, the device copies of the slab lists, the communication
 * stream and its events; it does not own the communicator.  Slab lists are host arrays; both sides
 * must order the slabs of a pair of ranks identically (HydroSim: by receiver gid and offset index,
 * what bnd_info.cpp:105-230 derives per buffer).  Offsets / counts are in doubles, sized for
 * cons+prim (2*nvar values per cell); single-field exchanges use half of every offset. */
typedef struct apk_halo_peer {
  int rank;                          /* NCCL rank of the peer                         */
  long long send_off, send_cnt;      /* segment of the send buffer for this peer      */
  long long recv_off, recv_cnt;      /* segment of the receive buffer from this peer  */
} apk_halo_peer_t;
typedef struct apk_halo_plan apk_halo_plan_t;
int apk_halo_plan_create(const apk_halo_item_t *h_send_items, int nsend,
                         const apk_halo_item_t *h_recv_items, int nrecv,
                         const apk_halo_peer_t *h_peers, int npeers, long long send_total,
                         long long recv_total, void *nccl_comm, apk_halo_plan_t **out);
void apk_halo_plan_destroy(apk_halo_plan_t *plan);
int apk_halo_start(apk_halo_plan_t *plan, const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                   int with_prim, void *stream);
int apk_halo_test(apk_halo_plan_t *plan);
int apk_halo_finish(apk_halo_plan_t *plan, const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                    int with_prim, void *stream);

/* ---- peer-to-peer set-up for apk_stage_fused_p2p: every rank publishes `nbufs` device
 * allocations (cudaIpcGetMemHandle), the handles are all-gathered over the NCCL communicator and the
 * buffers of the ranks flagged in h_need[nranks] are mapped (cudaIpcOpenMemHandle).
 * h_mapped[r * nbufs + f] receives rank r's buffer f as a pointer valid on THIS device (NULL for
 * ranks not needed / this rank).  *all_ok = 1 only when EVERY rank mapped everything it needs
 * (agreed with an all-reduce), otherwise nothing stays mapped and the caller uses apk_halo_*.
 * apk_stage_barrier: stream-ordered cross-rank barrier (4-byte all-reduce) that separates two fused
 * stages: all peer stores of stage s have landed before any rank starts stage s+1. */
int apk_p2p_map_peers(void *nccl_comm, int rank, int nranks, void *const *h_local_bufs, int nbufs,
                      const int *h_need, void **h_mapped, int *all_ok, void *stream);
int apk_p2p_unmap(void **h_mapped, int n);
int apk_stage_barrier(void *nccl_comm, int *d_word, void *stream);
/* min / sum all-reduce of one device value (replaces MPI_Allreduce of driver.cpp:237,
 * hydro.cpp:127): op 0 = min of a double, 1 = sum of an unsigned 64-bit count */
int apk_allreduce_scalar(void *nccl_comm, void *d_value, int op, void *stream);

/* ---- fused stage: calc_flux -> flux divergence update -> Dedner source -> cons->prim of the
 * interior cells (+ optional dt reduction on the last stage) without materialising fluxes
 * (K1-K3+K6+K7+K10(+K11) of SURVEY.md 2.3).  Legal only when nothing else consumes cons.flux
 * (no FOFC, no AMR flux correction, no unsplit diffusion; hydro_driver.cpp:517-531).
 * Reads prim_in (= block.prim), cons (u0), cons_u1; writes cons (interior) and prim_out.
 * prim_out[b] may not alias prim (neighbouring tiles still read it); d_prim_out is a device
 * array of nblocks pointers.  d_dtmin (device, may be NULL) accumulates min dx/(|v|+c). */
int apk_stage_fused(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                    double *const *d_prim_out, double gam0, double gam1, double beta_dt,
                    double c_h, double mindx, double alpha, int recon_override, int flags,
                    double *d_dtmin, void *stream);
#define APK_FUSED_SAVE_U1 1 /* flags: also store the incoming cons into cons_u1 (first stage:
                               replaces the u1 <- u0 DeepCopy of hydro_driver.cpp:474-495) */
#define APK_FUSED_PUSH_PRIM 2 /* flags: finished cells within ng of a block face are also written
                                 into the ghost cells of prim_out of the same-device neighbour
                                 blocks (nbr[] >= 0): replaces apk_copy_ghosts_local(prim)     */
#define APK_FUSED_PRIM_ONLY 8  /* flags: store prim_out only -- cons and cons_u1 stay untouched (the u1
                                 term is taken from cons).  For a predictor stage whose cons nobody
                                 reads again (VL2: gam0 == 0 in the corrector).                 */
#define APK_FUSED_U1_FROM_CONS 16 /* flags: read the u1 term from cons instead of cons_u1 (corrector
                                 after an APK_FUSED_PRIM_ONLY predictor; cons_u1 is not touched)  */
#define APK_FUSED_PUSH_CONS 4 /* same for cons (last stage: the caller-visible state after the
                                 cycle has valid cons ghosts like the reference's)            */

/* Same as apk_stage_fused with explicit ghost-push destinations: d_push is a device array
 * [nblocks][27] (neighbour offset index (o3+1)*9+(o2+1)*3+(o1+1)) holding the BLOCK BASE of the
 * neighbour's prim_out and cons arrays, or NULL to skip that neighbour.  The pointers may refer to
 * another GPU's memory (mapped with cudaIpcOpenMemHandle / peer access): the halo exchange with
 * remote neighbours then happens inside the stage kernel as peer stores over NVLink -- no pack /
 * ncclSend / ncclRecv / unpack (boundary_communication.cpp:48-348, communication_buffer.hpp
 * :209-236).  The caller must separate consecutive stages with a cross-rank barrier (any stream-
 * ordered collective, e.g. the dt all-reduce). */
typedef struct apk_push_dst {
  double *prim;
  double *cons;
} apk_push_dst_t;
int apk_stage_fused_p2p(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                        double *const *d_prim_out, const apk_push_dst_t *d_push, double gam0,
                        double gam1, double beta_dt, double c_h, double mindx, double alpha,
                        int recon_override, int flags, double *d_dtmin, void *stream);

/* Full-argument form of the fused stage (ABI 3); apk_stage_fused / apk_stage_fused_p2p are thin
 * wrappers that leave d_status / d_fofc_count NULL.
 *  d_status (device int, may be NULL): atomicMax'ed with 1 / 2 when a finished cell has negative
 *    density / pressure and no floor is active -- the PARTHENON_REQUIRE of ConsToPrim
 *    (adiabatic_hydro.hpp:77-79,111-113, adiabatic_glmmhd.hpp:97,137); the caller aborts on it.
 *  d_fofc_count (device uint64, may be NULL; flag APK_FUSED_FOFC_CHECK): the stage also evaluates
 *    the first attempt of FirstOrderFluxCorrect (hydro.cpp:1268-1306) on its own trial update --
 *    the number of cells with rho_new <= 0 or (E - KE - ME)_new <= 0 is added to *d_fofc_count.
 *    Zero means FOFC would not have touched any flux and the stage result is final (the common
 *    case: the whole Orszag-Tang golden run); non-zero means the caller must discard the stage
 *    (restore cons from the copy APK_FUSED_SAVE_U1 left in cons_u1, or from cons itself after an
 *    APK_FUSED_PRIM_ONLY stage) and redo it with the flux-materialising tasks
 *    apk_calc_fluxes -> apk_first_order_flux_correct -> apk_flux_div_update -> ... */
typedef struct apk_stage_args {
  double gam0, gam1, beta_dt;   /* low_storage_integrator.cpp:53-132 */
  double c_h, mindx, alpha;     /* GLM: hydro.cpp:102-143, dedner_source.cpp:34 */
  int recon_override;           /* >= 0 replaces desc->recon (VL2 predictor: DC) */
  int flags;                    /* APK_FUSED_* */
  double *d_dtmin;              /* may be NULL */
  int *d_status;                /* may be NULL */
  const apk_push_dst_t *d_push; /* may be NULL: same-device neighbours from nbr[] */
  unsigned long long *d_fofc_count; /* may be NULL */
  const void *d_tma_maps;       /* may be NULL.  Device array of one 128-byte CUtensorMap per block
                                   over its prim array (apk_tma_maps_create): selects the variant of
                                   the PPM/WENOZ stage that stages the halo'd prim tile of each plane
                                   with cp.async.bulk.tensor (TMA) instead of per-lane loads */
} apk_stage_args_t;
#define APK_FUSED_FOFC_CHECK 32
int apk_stage_fused_ex(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                       double *const *d_prim_out, const apk_stage_args_t *args, void *stream);

/* Tensor maps for the TMA variant of the fused stage: one 4-D map {n1, n2, n3, nvar} per block
 * over h_blocks[b].prim (HOST copy of the block structs; the prim pointers are device pointers),
 * box = (32+6) x (4+6) x 1 x nhydro, no swizzle, zero fill outside the array.  Returns a device
 * array (cudaMalloc; release with apk_tma_maps_destroy) in *d_maps.  Requires n1 even (16-byte
 * global strides).  Needs a driver that exports cuTensorMapEncodeTiled (CUDA >= 12). */
int apk_tma_maps_create(const apk_pack_desc_t *desc, const apk_block_t *h_blocks, void **d_maps);
void apk_tma_maps_destroy(void *d_maps);

/* copy cons -> cons_u1 for all blocks (u1 <- u0 at stage 1, hydro_driver.cpp:474-495) */
int apk_save_u1(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, void *stream);

/* ---- AMR building blocks (config 5).  One item = one (fine array, coarse buffer)
 * pair and an inclusive COARSE index range; arrays are C-order [nvar][n3][n2][n1]; interior start
 * index = ng in active dimensions on both levels (Parthenon's cellbounds / c_cellbounds). */
typedef struct apk_amr_item {
  double *fine;    /* fine block field (cons, or one direction of bnd_flux::cons)          */
  double *coarse;  /* coarse buffer, nx/2 + 2*ng cells per active dimension               */
  int nf[3];       /* fine extents (i,j,k), ghosts included                                */
  int nc[3];       /* coarse extents (i,j,k)                                               */
  int lo[3], hi[3]; /* inclusive coarse index range (i,j,k) to process                     */
  double xmin[3];  /* RegionSize::xmin of the fine block (first interior face)             */
  double dx[3];    /* fine cell sizes                                                      */
} apk_amr_item_t;
/* parthenon::refinement_ops::RestrictAverage (prolong_restrict/pr_ops.hpp:105-165): volume
 * weighted mean of the 2^ndim children (element 0), or area weighted mean of the 2^(ndim-1) fine
 * faces (element 1/2/3 = x1/x2/x3 face field: the restriction step of the flux correction,
 * bvals_in_one.hpp:69-80, bnd_info.cpp:71-103). */
int apk_restrict_average(int ndim, int ng, int nvar, int element, const apk_amr_item_t *d_items,
                         int nitems, int fp_mode, void *stream);
/* Hydro::refinement_ops::ProlongateCellMinModMultiD (src/hydro/prolongation/custom_ops.hpp
 * :49-186, registered for cons at hydro.cpp:780-781): multi-D minmod-limited prolongation that
 * writes the 2^ndim fine children of every coarse cell in the range. */
int apk_prolongate_minmod_multid(int ndim, int ng, int nvar, const apk_amr_item_t *d_items,
                                 int nitems, int fp_mode, void *stream);
/* Copy of 3-D boxes of all variables between two [nvar][n3][n2][n1] arrays of different extents:
 * the data movement of the multilevel exchange -- restricted fine cells (a child's coarse buffer)
 * into the coarse block's ghost cells, a coarser neighbour's interior cells into a block's coarse
 * buffer (the BoundaryInteriorSend / BoundaryExteriorRecv ranges of bnd_info.cpp:105-230 for
 * neighbours on other levels), restricted face fluxes into the coarse block's flux array (flux
 * correction, bvals_in_one.hpp:69-80).  One launch for all items. */
typedef struct apk_box_item {
  const double *src;
  double *dst;
  int ns[3], nd[3];  /* extents (i,j,k) of the source / destination arrays          */
  int s0[3], d0[3];  /* first cell (i,j,k) of the box in the source / destination   */
  int ext[3];        /* box extents (i,j,k)                                         */
} apk_box_item_t;
int apk_box_copy(const apk_box_item_t *d_items, int nitems, int nvar, void *stream);

/* refinement::gradient::PressureGradient (src/refinement/gradient.cpp:18-61): per block
 * max |grad p| / p over interior +- 1 -> d_maxeps[b] and the AmrTag d_tags[b] (-1 derefine,
 * 0 same, 1 refine). */
int apk_tag_pressure_gradient(const apk_pack_desc_t *desc, const apk_block_t *d_blocks,
                              double threshold, int *d_tags, double *d_maxeps, void *stream);
/* All `<refinement> type` criteria of src/refinement (hydro.cpp:788-816): per block a measure
 * (d_measure[b]) and the AmrTag (d_tags[b] = 1 when measure > refine_above, -1 when measure <
 * derefine_below, else 0):
 *   APK_REFINE_PRESSURE_GRADIENT    max |grad p| / p over interior +- 1 (gradient.cpp:18-61;
 *                                   derefine_below = 0.25 threshold)
 *   APK_REFINE_XYVELOCITY_GRADIENT  max sqrt((dv1/dx2)^2 + (dv2/dx1)^2), undivided differences
 *                                   (gradient.cpp:64-95; derefine_below = 0.5 threshold)
 *   APK_REFINE_MAXDENSITY           max density (other.cpp:18-44; maxdensity_refine_above /
 *                                   maxdensity_deref_below) */
enum { APK_REFINE_PRESSURE_GRADIENT = 0, APK_REFINE_XYVELOCITY_GRADIENT = 1, APK_REFINE_MAXDENSITY = 2 };
int apk_tag_refinement(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, int criterion,
                       double refine_above, double derefine_below, int *d_tags, double *d_measure,
                       void *stream);

/* ---- history reducers HydroHst<...> (src/hydro/hydro.cpp:145-208, registered :422-441): volume
 * sums over the interior cells of the pack: out8 = {mass, mom1, mom2, mom3, kinetic energy, total
 * energy, magnetic energy, relative div(B) error}; the last two are 0 for hydro / 1-D.  Blocking
 * (the result lands in host memory).  Deterministic: fixed-order two-level tree, so two calls on the
 * same state agree bit for bit (the reference's Kokkos reduction order is unspecified; compare with
 * a tolerance).  d_work: device scratch of apk_history_workspace_bytes(). */
size_t apk_history_workspace_bytes(const apk_pack_desc_t *desc);
int apk_history_sums(const apk_pack_desc_t *desc, const apk_block_t *d_blocks, void *d_work,
                     double *h_out8, void *stream);

/* FP64 pipe microbenchmark (roofline evidence): runs n_iter dependent-chain-free DFMAs per
 * thread, returns measured FLOP count in *flops; time it with CUDA events from the caller */
int apk_fp64_probe(int n_blocks, int n_threads, int n_iter, double *d_out, double *flops,
                   void *stream);

#ifdef __cplusplus
}
#endif
#endif /* APK_B200_H_ */
