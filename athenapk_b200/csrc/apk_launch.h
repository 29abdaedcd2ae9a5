// apk_launch.h -- internal: host-side launchers implemented once per floating-point mode.
// apk_tasks.cu / apk_fused.cu are compiled twice (APK_NS=apk_fast with FMA contraction,
// APK_NS=apk_strict with -fmad=false); apk_capi.cu dispatches on apk_pack_desc_t::fp_mode.
#ifndef APK_LAUNCH_H_
#define APK_LAUNCH_H_

#include <cuda_runtime.h>

#include "../../include/apk_b200.h"

#define APK_DECLARE_LAUNCHERS(NS)                                                               \
  namespace NS {                                                                                \
  int launch_calc_fluxes(const apk_pack_desc_t &d, const apk_block_t *blocks, double c_h,       \
                         int recon, cudaStream_t s);                                            \
  int launch_fofc_flag(const apk_pack_desc_t &d, const apk_block_t *blocks, double g0,          \
                       double g1, double bdt, int attempt, unsigned char *flags,                \
                       unsigned long long *counter, cudaStream_t s);                            \
  int launch_fofc_fix(const apk_pack_desc_t &d, const apk_block_t *blocks, double c_h,          \
                      const unsigned char *flags, cudaStream_t s);                              \
  int launch_flux_div_update(const apk_pack_desc_t &d, const apk_block_t *blocks, double g0,    \
                             double g1, double bdt, cudaStream_t s);                            \
  int launch_dedner(const apk_pack_desc_t &d, const apk_block_t *blocks, double coeff,          \
                    double bdt, int extended, cudaStream_t s);                                  \
  int launch_cons_to_prim(const apk_pack_desc_t &d, const apk_block_t *blocks, int *status,     \
                          cudaStream_t s);                                                      \
  int launch_min_dt(const apk_pack_desc_t &d, const apk_block_t *blocks, double *d_out,         \
                    cudaStream_t s);                                                            \
  int launch_stage_fused(const apk_pack_desc_t &d, const apk_block_t *blocks,                   \
                         double *const *prim_out, const apk_stage_args_t &a,                    \
                         double dedner_coeff, cudaStream_t s);                                  \
  int launch_restrict_average(int ndim, int ng, int nvar, int el, const apk_amr_item_t *items,  \
                              int nitems, cudaStream_t s);                                      \
  int launch_prolongate_minmod(int ndim, int ng, int nvar, const apk_amr_item_t *items,         \
                               int nitems, cudaStream_t s);                                     \
  int launch_refinement_tags(const apk_pack_desc_t &d, const apk_block_t *blocks,               \
                             int criterion, double refine_above, double derefine_below,         \
                             int *tags, double *measure, cudaStream_t s);                       \
  int launch_reset_fluxes(const apk_pack_desc_t &d, const apk_block_t *blocks, cudaStream_t s); \
  int launch_diff_fluxes(const apk_pack_desc_t &d, const apk_block_t *blocks,                   \
                         const apk_diffusion_t &df, cudaStream_t s);                            \
  int launch_rkl2_first(const apk_pack_desc_t &d, const apk_block_t *blocks,                    \
                        double *const *my0, double *const *yjm2, double mt, cudaStream_t s);    \
  int launch_rkl2_other(const apk_pack_desc_t &d, const apk_block_t *blocks,                    \
                        double *const *my0, double *const *yjm2, double mu_j, double nu_j,      \
                        double c0, double mtt, double gtt, cudaStream_t s);                     \
  int launch_min_dt_conduction(const apk_pack_desc_t &d, const apk_block_t *blocks,             \
                               const apk_diffusion_t &df, double *d_out, cudaStream_t s);       \
  size_t history_workspace_bytes(const apk_pack_desc_t &d);                                     \
  int launch_history(const apk_pack_desc_t &d, const apk_block_t *blocks, double *work,         \
                     cudaStream_t s);                                                           \
  }

APK_DECLARE_LAUNCHERS(apk_fast)
APK_DECLARE_LAUNCHERS(apk_strict)
// A/B variant of the fused stage with TMA-staged prim tiles (apk_fused.cu built with
// -DAPK_FUSED_TMA=1; GLM-MHD PPM+HLLD 3-D only)
namespace apk_fast_tma {
int launch_stage_fused(const apk_pack_desc_t &d, const apk_block_t *blocks, double *const *prim_out,
                       const apk_stage_args_t &a, double dedner_coeff, cudaStream_t s);
}
#define APK_TMA_TX 32
#define APK_TMA_TY 4

// mode-independent (pure data movement) launchers, apk_halo.cu
namespace apk_common {
int launch_fill_ghosts(const apk_pack_desc_t &d, const apk_block_t *blocks, int with_prim,
                       cudaStream_t s);
int launch_physical_bcs(const apk_pack_desc_t &d, const apk_block_t *blocks, unsigned face_mask,
                        int with_prim, cudaStream_t s);
int launch_halo_pack(const apk_pack_desc_t &d, const apk_block_t *blocks,
                     const apk_halo_item_t *items, int nitems, int with_prim, double *buf,
                     cudaStream_t s);
int launch_halo_unpack(const apk_pack_desc_t &d, const apk_block_t *blocks,
                       const apk_halo_item_t *items, int nitems, int with_prim,
                       const double *buf, cudaStream_t s);
int launch_save_u1(const apk_pack_desc_t &d, const apk_block_t *blocks, cudaStream_t s);
int launch_box_copy(const apk_box_item_t *items, int nitems, int nvar, cudaStream_t s);
int launch_fp64_probe(int nb, int nt, int niter, double *out, cudaStream_t s);
void count_launch(int n = 1);
int current_device();
int sm_count();  // of the current device
long long launches();
void reset_launches();
const char *last_error();
void set_error(const char *msg);
}  // namespace apk_common

// geometry shared by all kernels
struct ApkGeom {
  int nb, nvar, nhydro, nscalars, ndim, ng;
  int n1, n2, n3;  // array extents incl. ghosts
  int is, ie, js, je, ks, ke;
  long ncell;  // n1*n2*n3
};

static inline ApkGeom apk_make_geom(const apk_pack_desc_t &d) {
  ApkGeom g;
  g.nb = d.nblocks;
  g.nvar = d.nvar;
  g.nscalars = d.nscalars;
  g.nhydro = d.nvar - d.nscalars;
  g.ndim = d.ndim;
  g.ng = d.ng;
  g.n1 = d.nx[0] + 2 * d.ng;
  g.n2 = (d.ndim >= 2) ? d.nx[1] + 2 * d.ng : 1;
  g.n3 = (d.ndim >= 3) ? d.nx[2] + 2 * d.ng : 1;
  g.is = d.ng;
  g.ie = d.ng + d.nx[0] - 1;
  g.js = (d.ndim >= 2) ? d.ng : 0;
  g.je = (d.ndim >= 2) ? d.ng + d.nx[1] - 1 : 0;
  g.ks = (d.ndim >= 3) ? d.ng : 0;
  g.ke = (d.ndim >= 3) ? d.ng + d.nx[2] - 1 : 0;
  g.ncell = (long)g.n1 * g.n2 * g.n3;
  return g;
}

#endif  // APK_LAUNCH_H_
