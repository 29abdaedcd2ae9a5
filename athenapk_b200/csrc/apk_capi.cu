// apk_capi.cu -- the C-ABI of include/apk_b200.h: argument checks, FP-mode dispatch and the
// few host-side pieces the reference also keeps on the host (exp() of the Dedner coefficient,
// the blocking reductions of dt and of the FOFC correction count).
#include <cmath>
#include <cstdio>
#include <cstring>

#include <cuda.h>

#include <vector>

#include "apk_launch.h"

namespace {
inline cudaStream_t S(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int check(const apk_pack_desc_t *d, const void *blocks, const char *who) {
  if (!d || !blocks) {
    apk_common::set_error("null descriptor or block array");
    return 1;
  }
  const int nh = (d->fluid == APK_FLUID_GLMMHD) ? 9 : 5;
  if (d->nblocks <= 0 || d->ndim < 1 || d->ndim > 3 || d->nvar != nh + d->nscalars ||
      d->nscalars < 0 || d->ng < 1 || d->nx[0] < 1 ||
      (d->fluid != APK_FLUID_EULER && d->fluid != APK_FLUID_GLMMHD) ||
      (d->fp_mode != APK_FP_FAST && d->fp_mode != APK_FP_STRICT)) {
    apk_common::set_error(who);
    return 1;
  }
  return 0;
}

int cuda_ok(const char *who) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    static thread_local char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", who, cudaGetErrorString(e));
    apk_common::set_error(buf);
    return 2;
  }
  return 0;
}
}  // namespace

#define DISPATCH(d, call) ((d)->fp_mode == APK_FP_STRICT ? apk_strict::call : apk_fast::call)

extern "C" {

const char *apk_last_error(void) { return apk_common::last_error(); }
int apk_abi_version(void) { return APK_ABI_VERSION; }
long long apk_launch_count(void) { return apk_common::launches(); }
void apk_reset_launch_count(void) { apk_common::reset_launches(); }

int apk_calc_fluxes(const apk_pack_desc_t *d, const apk_block_t *blocks, double c_h,
                    int recon_override, void *stream) {
  if (check(d, blocks, "apk_calc_fluxes: bad descriptor")) return 1;
  int rc = DISPATCH(d, launch_calc_fluxes(*d, blocks, c_h, recon_override, S(stream)));
  return rc ? rc : cuda_ok("apk_calc_fluxes");
}

size_t apk_fofc_workspace_bytes(const apk_pack_desc_t *d) {
  const ApkGeom g = apk_make_geom(*d);
  return (size_t)g.ncell * g.nb + 16;
}

int apk_first_order_flux_correct(const apk_pack_desc_t *d, const apk_block_t *blocks,
                                 double gam0, double gam1, double beta_dt, double c_h,
                                 void *d_work, long long *n_corrected, void *stream) {
  if (check(d, blocks, "apk_first_order_flux_correct: bad descriptor")) return 1;
  if (!d_work) {
    apk_common::set_error("apk_first_order_flux_correct: null workspace");
    return 1;
  }
  const ApkGeom g = apk_make_geom(*d);
  // workspace: [counter (16 B)] [flags: nb*ncell bytes]
  unsigned long long *counter = reinterpret_cast<unsigned long long *>(d_work);
  unsigned char *flags = reinterpret_cast<unsigned char *>(d_work) + 16;
  long long total = 0;
  int attempts = 0;
  unsigned long long num_corrected = 0;
  do {  // hydro.cpp:1263-1339
    cudaMemsetAsync(d_work, 0, (size_t)g.ncell * g.nb + 16, S(stream));
    DISPATCH(d, launch_fofc_flag(*d, blocks, gam0, gam1, beta_dt, attempts, flags, counter, S(stream)));
    cudaMemcpyAsync(&num_corrected, counter, sizeof(num_corrected), cudaMemcpyDeviceToHost,
                    S(stream));
    cudaStreamSynchronize(S(stream));
    if (num_corrected > 0) DISPATCH(d, launch_fofc_fix(*d, blocks, c_h, flags, S(stream)));
    total += (long long)num_corrected;
    attempts += 1;
  } while (num_corrected > 0 && attempts < 4);
  if (n_corrected) *n_corrected = total;
  return cuda_ok("apk_first_order_flux_correct");
}

int apk_flux_div_update(const apk_pack_desc_t *d, const apk_block_t *blocks, double gam0,
                        double gam1, double beta_dt, void *stream) {
  if (check(d, blocks, "apk_flux_div_update: bad descriptor")) return 1;
  DISPATCH(d, launch_flux_div_update(*d, blocks, gam0, gam1, beta_dt, S(stream)));
  return cuda_ok("apk_flux_div_update");
}

int apk_dedner_source(const apk_pack_desc_t *d, const apk_block_t *blocks, double c_h,
                      double mindx, double alpha, double beta_dt, int extended, void *stream) {
  if (check(d, blocks, "apk_dedner_source: bad descriptor")) return 1;
  if (d->fluid != APK_FLUID_GLMMHD) return 0;
  // dedner_source.cpp:34 (host std::exp, like the reference)
  const double coeff = std::exp(-alpha * c_h * beta_dt / mindx);
  DISPATCH(d, launch_dedner(*d, blocks, coeff, beta_dt, extended, S(stream)));
  return cuda_ok("apk_dedner_source");
}

int apk_cons_to_prim(const apk_pack_desc_t *d, const apk_block_t *blocks, int *d_status,
                     void *stream) {
  if (check(d, blocks, "apk_cons_to_prim: bad descriptor")) return 1;
  DISPATCH(d, launch_cons_to_prim(*d, blocks, d_status, S(stream)));
  return cuda_ok("apk_cons_to_prim");
}

int apk_min_dt_async(const apk_pack_desc_t *d, const apk_block_t *blocks, double *d_out,
                     void *stream) {
  if (check(d, blocks, "apk_min_dt: bad descriptor")) return 1;
  DISPATCH(d, launch_min_dt(*d, blocks, d_out, S(stream)));
  return cuda_ok("apk_min_dt");
}

int apk_min_dt(const apk_pack_desc_t *d, const apk_block_t *blocks, double cfl, void *d_work,
               double *h_dt, void *stream) {
  int rc = apk_min_dt_async(d, blocks, reinterpret_cast<double *>(d_work), stream);
  if (rc) return rc;
  double m = 0.0;
  cudaMemcpyAsync(&m, d_work, sizeof(double), cudaMemcpyDeviceToHost, S(stream));
  cudaStreamSynchronize(S(stream));
  *h_dt = cfl * m;  // hydro.cpp:909
  return cuda_ok("apk_min_dt");
}

int apk_copy_ghosts_local(const apk_pack_desc_t *d, const apk_block_t *blocks, int with_prim,
                          void *stream) {
  if (check(d, blocks, "apk_copy_ghosts_local: bad descriptor")) return 1;
  apk_common::launch_fill_ghosts(*d, blocks, with_prim, S(stream));
  return cuda_ok("apk_copy_ghosts_local");
}

int apk_apply_physical_bcs(const apk_pack_desc_t *d, const apk_block_t *blocks,
                           unsigned bc_face_mask, int with_prim, void *stream) {
  if (check(d, blocks, "apk_apply_physical_bcs: bad descriptor")) return 1;
  apk_common::launch_physical_bcs(*d, blocks, bc_face_mask, with_prim, S(stream));
  return cuda_ok("apk_apply_physical_bcs");
}

int apk_fill_ghosts_local(const apk_pack_desc_t *d, const apk_block_t *blocks,
                          unsigned bc_face_mask, int with_prim, void *stream) {
  int rc = apk_copy_ghosts_local(d, blocks, with_prim, stream);
  if (rc) return rc;
  return apk_apply_physical_bcs(d, blocks, bc_face_mask, with_prim, stream);
}

int apk_halo_pack(const apk_pack_desc_t *d, const apk_block_t *blocks,
                  const apk_halo_item_t *items, int nitems, int with_prim, double *buf,
                  void *stream) {
  if (check(d, blocks, "apk_halo_pack: bad descriptor")) return 1;
  apk_common::launch_halo_pack(*d, blocks, items, nitems, with_prim, buf, S(stream));
  return cuda_ok("apk_halo_pack");
}

int apk_halo_unpack(const apk_pack_desc_t *d, const apk_block_t *blocks,
                    const apk_halo_item_t *items, int nitems, int with_prim, const double *buf,
                    void *stream) {
  if (check(d, blocks, "apk_halo_unpack: bad descriptor")) return 1;
  apk_common::launch_halo_unpack(*d, blocks, items, nitems, with_prim, buf, S(stream));
  return cuda_ok("apk_halo_unpack");
}

int apk_stage_fused_ex(const apk_pack_desc_t *d, const apk_block_t *blocks,
                       double *const *d_prim_out, const apk_stage_args_t *a, void *stream) {
  if (check(d, blocks, "apk_stage_fused: bad descriptor")) return 1;
  if (!d_prim_out || !a) {
    apk_common::set_error("apk_stage_fused: null prim_out / args");
    return 1;
  }
  if ((a->flags & APK_FUSED_FOFC_CHECK) && !a->d_fofc_count) {
    apk_common::set_error("apk_stage_fused: APK_FUSED_FOFC_CHECK needs d_fofc_count");
    return 1;
  }
  // dedner_source.cpp:34 (host std::exp, like the reference)
  const double coeff =
      (d->fluid == APK_FLUID_GLMMHD) ? std::exp(-a->alpha * a->c_h * a->beta_dt / a->mindx) : 1.0;
  int rc;
  const int recon = (a->recon_override < 0) ? d->recon : a->recon_override;
  if (a->d_tma_maps && d->fp_mode == APK_FP_FAST && d->ndim == 3 && d->fluid == APK_FLUID_GLMMHD &&
      d->riemann == APK_RIEMANN_HLLD && recon == APK_RECON_PPM && d->nscalars == 0)
    rc = apk_fast_tma::launch_stage_fused(*d, blocks, d_prim_out, *a, coeff, S(stream));
  else
    rc = DISPATCH(d, launch_stage_fused(*d, blocks, d_prim_out, *a, coeff, S(stream)));
  return rc ? rc : cuda_ok("apk_stage_fused");
}

int apk_stage_fused_p2p(const apk_pack_desc_t *d, const apk_block_t *blocks,
                        double *const *d_prim_out, const apk_push_dst_t *d_push, double gam0,
                        double gam1, double beta_dt, double c_h, double mindx, double alpha,
                        int recon_override, int flags, double *d_dtmin, void *stream) {
  apk_stage_args_t a;
  memset(&a, 0, sizeof(a));
  a.gam0 = gam0;
  a.gam1 = gam1;
  a.beta_dt = beta_dt;
  a.c_h = c_h;
  a.mindx = mindx;
  a.alpha = alpha;
  a.recon_override = recon_override;
  a.flags = flags;
  a.d_dtmin = d_dtmin;
  a.d_push = d_push;
  return apk_stage_fused_ex(d, blocks, d_prim_out, &a, stream);
}

int apk_stage_fused(const apk_pack_desc_t *d, const apk_block_t *blocks,
                    double *const *d_prim_out, double gam0, double gam1, double beta_dt,
                    double c_h, double mindx, double alpha, int recon_override, int flags,
                    double *d_dtmin, void *stream) {
  return apk_stage_fused_p2p(d, blocks, d_prim_out, nullptr, gam0, gam1, beta_dt, c_h, mindx, alpha,
                             recon_override, flags, d_dtmin, stream);
}

// ---- TMA tensor maps (driver entry point fetched at run time: no link dependency on libcuda)
int apk_tma_maps_create(const apk_pack_desc_t *d, const apk_block_t *h_blocks, void **d_maps) {
  if (!d || !h_blocks || !d_maps) {
    apk_common::set_error("apk_tma_maps_create: null argument");
    return 1;
  }
  const ApkGeom g = apk_make_geom(*d);
  if (g.ndim != 3 || (g.n1 % 2) != 0) {
    apk_common::set_error("apk_tma_maps_create: needs a 3-D pack with an even n1 (16 B strides)");
    return 1;
  }
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !fn) {
    apk_common::set_error("apk_tma_maps_create: cuTensorMapEncodeTiled not available");
    return 2;
  }
  const int nh = g.nhydro;
  std::vector<CUtensorMap> maps(d->nblocks);
  const cuuint64_t dims[4] = {(cuuint64_t)g.n1, (cuuint64_t)g.n2, (cuuint64_t)g.n3, (cuuint64_t)nh};
  const cuuint64_t strides[3] = {(cuuint64_t)g.n1 * 8, (cuuint64_t)g.n1 * g.n2 * 8,
                                 (cuuint64_t)g.ncell * 8};
  const cuuint32_t box[4] = {APK_TMA_TX + 6, APK_TMA_TY + 6, 1, (cuuint32_t)nh};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int b = 0; b < d->nblocks; ++b) {
    CUresult r = reinterpret_cast<encode_fn>(fn)(
        &maps[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, h_blocks[b].prim, dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      static thread_local char buf[128];
      snprintf(buf, sizeof(buf), "apk_tma_maps_create: cuTensorMapEncodeTiled failed (%d)", (int)r);
      apk_common::set_error(buf);
      return 2;
    }
  }
  void *dm = nullptr;
  if (cudaMalloc(&dm, maps.size() * sizeof(CUtensorMap)) != cudaSuccess ||
      cudaMemcpy(dm, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) !=
          cudaSuccess)
    return cuda_ok("apk_tma_maps_create") ? 2 : 2;
  *d_maps = dm;
  return 0;
}

void apk_tma_maps_destroy(void *d_maps) { cudaFree(d_maps); }

int apk_restrict_average(int ndim, int ng, int nvar, int element, const apk_amr_item_t *d_items,
                         int nitems, int fp_mode, void *stream) {
  if (ndim < 1 || ndim > 3 || ng < 0 || nvar < 1 || element < 0 || element > ndim ||
      (nitems > 0 && !d_items)) {
    apk_common::set_error("apk_restrict_average: bad arguments");
    return 1;
  }
  if (fp_mode == APK_FP_STRICT)
    apk_strict::launch_restrict_average(ndim, ng, nvar, element, d_items, nitems, S(stream));
  else
    apk_fast::launch_restrict_average(ndim, ng, nvar, element, d_items, nitems, S(stream));
  return cuda_ok("apk_restrict_average");
}

int apk_prolongate_minmod_multid(int ndim, int ng, int nvar, const apk_amr_item_t *d_items,
                                 int nitems, int fp_mode, void *stream) {
  if (ndim < 1 || ndim > 3 || ng < 1 || nvar < 1 || (nitems > 0 && !d_items)) {
    apk_common::set_error("apk_prolongate_minmod_multid: bad arguments");
    return 1;
  }
  if (fp_mode == APK_FP_STRICT)
    apk_strict::launch_prolongate_minmod(ndim, ng, nvar, d_items, nitems, S(stream));
  else
    apk_fast::launch_prolongate_minmod(ndim, ng, nvar, d_items, nitems, S(stream));
  return cuda_ok("apk_prolongate_minmod_multid");
}

int apk_tag_refinement(const apk_pack_desc_t *d, const apk_block_t *blocks, int criterion,
                       double refine_above, double derefine_below, int *d_tags, double *d_measure,
                       void *stream) {
  if (check(d, blocks, "apk_tag_refinement: bad descriptor")) return 1;
  if (!d_tags || !d_measure || criterion < 0 || criterion > 2) {
    apk_common::set_error("apk_tag_refinement: null output or unknown criterion");
    return 1;
  }
  if (criterion == APK_REFINE_XYVELOCITY_GRADIENT && d->ndim < 2) {
    apk_common::set_error("apk_tag_refinement: the velocity-gradient criterion needs a 2-D or 3-D mesh");
    return 1;
  }
  DISPATCH(d, launch_refinement_tags(*d, blocks, criterion, refine_above, derefine_below, d_tags,
                                     d_measure, S(stream)));
  return cuda_ok("apk_tag_refinement");
}

int apk_tag_pressure_gradient(const apk_pack_desc_t *d, const apk_block_t *blocks, double threshold,
                              int *d_tags, double *d_maxeps, void *stream) {
  return apk_tag_refinement(d, blocks, APK_REFINE_PRESSURE_GRADIENT, threshold, 0.25 * threshold,
                            d_tags, d_maxeps, stream);
}

size_t apk_history_workspace_bytes(const apk_pack_desc_t *d) {
  return d ? apk_fast::history_workspace_bytes(*d) : 0;
}

int apk_history_sums(const apk_pack_desc_t *d, const apk_block_t *blocks, void *d_work,
                     double *h_out8, void *stream) {
  if (check(d, blocks, "apk_history_sums: bad descriptor")) return 1;
  if (!d_work || !h_out8) {
    apk_common::set_error("apk_history_sums: null workspace / output");
    return 1;
  }
  double *work = static_cast<double *>(d_work);
  DISPATCH(d, launch_history(*d, blocks, work, S(stream)));
  const size_t off = (size_t)d->nblocks * 8 * 8;  // kHistSlices partials per block
  if (cudaMemcpyAsync(h_out8, work + off, 8 * sizeof(double), cudaMemcpyDeviceToHost, S(stream)) !=
          cudaSuccess ||
      cudaStreamSynchronize(S(stream)) != cudaSuccess)
    return cuda_ok("apk_history_sums") ? 1 : 1;
  return cuda_ok("apk_history_sums");
}

int apk_box_copy(const apk_box_item_t *d_items, int nitems, int nvar, void *stream) {
  if (nitems < 0 || nvar < 1 || (nitems > 0 && !d_items)) {
    apk_common::set_error("apk_box_copy: bad arguments");
    return 1;
  }
  apk_common::launch_box_copy(d_items, nitems, nvar, S(stream));
  return cuda_ok("apk_box_copy");
}

int apk_reset_fluxes(const apk_pack_desc_t *d, const apk_block_t *blocks, void *stream) {
  if (check(d, blocks, "apk_reset_fluxes: bad descriptor")) return 1;
  DISPATCH(d, launch_reset_fluxes(*d, blocks, S(stream)));
  return cuda_ok("apk_reset_fluxes");
}

int apk_calc_diff_fluxes(const apk_pack_desc_t *d, const apk_block_t *blocks,
                         const apk_diffusion_t *diff, void *stream) {
  if (check(d, blocks, "apk_calc_diff_fluxes: bad descriptor")) return 1;
  if (!diff) {
    apk_common::set_error("apk_calc_diff_fluxes: null diffusion parameters");
    return 1;
  }
  if (DISPATCH(d, launch_diff_fluxes(*d, blocks, *diff, S(stream)))) return 1;
  return cuda_ok("apk_calc_diff_fluxes");
}

int apk_rkl2_step_first(const apk_pack_desc_t *d, const apk_block_t *blocks, double *const *d_my0,
                        double *const *d_yjm2, int s_rkl, double tau, void *stream) {
  if (check(d, blocks, "apk_rkl2_step_first: bad descriptor")) return 1;
  // Meyer+2014 eq. (18), hydro_driver.cpp:103-105
  const double mu_tilde_1 = 4. / 3. /
                            (static_cast<double>(s_rkl) * static_cast<double>(s_rkl) +
                             static_cast<double>(s_rkl) - 2.);
  DISPATCH(d, launch_rkl2_first(*d, blocks, d_my0, d_yjm2, mu_tilde_1 * tau, S(stream)));
  return cuda_ok("apk_rkl2_step_first");
}

int apk_rkl2_step_other(const apk_pack_desc_t *d, const apk_block_t *blocks, double *const *d_my0,
                        double *const *d_yjm2, double mu_j, double nu_j, double mu_tilde_j,
                        double gamma_tilde_j, double tau, void *stream) {
  if (check(d, blocks, "apk_rkl2_step_other: bad descriptor")) return 1;
  DISPATCH(d, launch_rkl2_other(*d, blocks, d_my0, d_yjm2, mu_j, nu_j, 1.0 - mu_j - nu_j,
                                mu_tilde_j * tau, gamma_tilde_j * tau, S(stream)));
  return cuda_ok("apk_rkl2_step_other");
}

int apk_min_dt_conduction(const apk_pack_desc_t *d, const apk_block_t *blocks,
                          const apk_diffusion_t *diff, double *d_out, void *stream) {
  if (check(d, blocks, "apk_min_dt_conduction: bad descriptor")) return 1;
  if (!diff || !d_out) {
    apk_common::set_error("apk_min_dt_conduction: null argument");
    return 1;
  }
  DISPATCH(d, launch_min_dt_conduction(*d, blocks, *diff, d_out, S(stream)));
  return cuda_ok("apk_min_dt_conduction");
}

int apk_save_u1(const apk_pack_desc_t *d, const apk_block_t *blocks, void *stream) {
  if (check(d, blocks, "apk_save_u1: bad descriptor")) return 1;
  apk_common::launch_save_u1(*d, blocks, S(stream));
  return cuda_ok("apk_save_u1");
}

int apk_fp64_probe(int n_blocks, int n_threads, int n_iter, double *d_out, double *flops,
                   void *stream) {
  apk_common::launch_fp64_probe(n_blocks, n_threads, n_iter, d_out, S(stream));
  if (flops) *flops = 2.0 * 8.0 * (double)n_iter * (double)n_blocks * (double)n_threads;
  return cuda_ok("apk_fp64_probe");
}

}  // extern "C"
