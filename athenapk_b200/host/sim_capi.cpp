// sim_capi.cpp -- flat C interface of the host driver (HydroSim) for tests/bench bindings.
// Declared in include/apk_b200_sim.h.
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "../../include/apk_b200_sim.h"
#include "hydro_sim.hpp"

using apk_host::HydroSim;
using apk_host::SimConfig;

namespace {
thread_local std::string g_sim_err;
template <class F>
int guarded(F &&f) {
  try {
    return f();
  } catch (const std::exception &e) {
    g_sim_err = e.what();
    return -100;
  }
}
}  // namespace

extern "C" {

const char *apk_sim_last_error(void) { return g_sim_err.c_str(); }

void apk_sim_default_config(apk_sim_config_t *c) {
  SimConfig d;
  std::memset(c, 0, sizeof(*c));
  c->fluid = d.fluid; c->recon = d.recon; c->riemann = d.riemann; c->integrator = d.integrator;
  for (int i = 0; i < 3; ++i) { c->nx[i] = 1; c->mb[i] = 1; c->xmin[i] = -0.5; c->xmax[i] = 0.5; }
  c->ng = d.ng; c->gamma = d.gamma; c->cfl = d.cfl;
  c->dfloor = d.dfloor; c->pfloor = d.pfloor; c->efloor = d.efloor;
  c->vceil = d.vceil; c->eceil = d.eceil;
  c->glmmhd_alpha = d.glmmhd_alpha; c->fp_mode = d.fp_mode; c->fused = d.fused;
  c->rank = 0; c->nranks = 1;
  c->cfl_diff = d.cfl_diff; c->rkl2_max_dt_ratio = d.rkl2_max_dt_ratio;
  c->conduction_sat_phi = d.conduction_sat_phi;
  c->numlevel = d.numlevel; c->derefine_count = d.derefine_count;
}

apk_sim_t *apk_sim_create(const apk_sim_config_t *c) {
  HydroSim *s = nullptr;
  guarded([&] {
    SimConfig k;
    k.fluid = c->fluid; k.recon = c->recon; k.riemann = c->riemann; k.integrator = c->integrator;
    for (int i = 0; i < 3; ++i) { k.nx[i] = c->nx[i]; k.mb[i] = c->mb[i]; k.xmin[i] = c->xmin[i]; k.xmax[i] = c->xmax[i]; }
    for (int i = 0; i < 6; ++i) k.bc[i] = c->bc[i];
    k.ng = c->ng; k.gamma = c->gamma; k.cfl = c->cfl;
    k.dfloor = c->dfloor; k.pfloor = c->pfloor; k.efloor = c->efloor; k.vceil = c->vceil; k.eceil = c->eceil;
    k.fofc = c->fofc; k.dedner_extended = c->dedner_extended; k.glmmhd_alpha = c->glmmhd_alpha;
    k.nscalars = c->nscalars; k.fp_mode = c->fp_mode; k.fused = c->fused;
    k.rank = c->rank; k.nranks = c->nranks; k.plan_only = c->plan_only;
    k.nref = c->nref;
    for (int r = 0; r < 4; ++r) {
      k.ref_level[r] = c->ref_level[r];
      for (int i = 0; i < 3; ++i) { k.ref_xmin[r][i] = c->ref_xmin[r][i]; k.ref_xmax[r][i] = c->ref_xmax[r][i]; }
    }
    k.diffint = c->diffint; k.conduction = c->conduction; k.viscosity = c->viscosity;
    k.resistivity = c->resistivity; k.thermal_diff_coeff = c->thermal_diff_coeff;
    k.mom_diff_coeff = c->mom_diff_coeff; k.ohm_diff_coeff = c->ohm_diff_coeff;
    k.cfl_diff = c->cfl_diff; k.rkl2_max_dt_ratio = c->rkl2_max_dt_ratio;
    k.conduction_sat_phi = c->conduction_sat_phi;
    k.adaptive = c->adaptive; k.numlevel = c->numlevel; k.derefine_count = c->derefine_count;
    k.refine_threshold = c->refine_threshold;
    k.refine_type = c->refine_type; k.refine_threshold2 = c->refine_threshold2;
    s = new HydroSim(k);
    return 0;
  });
  return reinterpret_cast<apk_sim_t *>(s);
}

void apk_sim_destroy(apk_sim_t *s) { delete reinterpret_cast<HydroSim *>(s); }

#define SIM(s) (reinterpret_cast<HydroSim *>(s))

int apk_sim_ml_num_items(const apk_sim_t *s) {
  return (int)reinterpret_cast<const HydroSim *>(s)->ml_desc_.size();
}
int apk_sim_ml_item(const apk_sim_t *s, int i, int *out10) {
  const auto &v = reinterpret_cast<const HydroSim *>(s)->ml_desc_;
  if (i < 0 || i >= (int)v.size()) return -1;
  const auto &m = v[i];
  const int o[10] = {m.list, m.exec, m.src, m.dst, m.lo[0], m.lo[1], m.lo[2], m.ext[0], m.ext[1], m.ext[2]};
  for (int k = 0; k < 10; ++k) out10[k] = o[k];
  return 0;
}
int apk_sim_amr_counts(const apk_sim_t *s, int *nbnew, int *nbdel, int *per_level8) {
  const HydroSim *h = reinterpret_cast<const HydroSim *>(s);
  if (nbnew) *nbnew = h->nbnew;
  if (nbdel) *nbdel = h->nbdel;
  if (per_level8) {
    for (int l = 0; l < 8; ++l) per_level8[l] = 0;
    for (int b = 0; b < h->NumLocalBlocks(); ++b)
      if (h->Block(b).level < 8) per_level8[h->Block(b).level]++;
  }
  return 0;
}
int apk_sim_amr_update_tree(const apk_sim_config_t *c, int nold, const int *leaves4,
                            const int *flags, int cap, int *out4, int *nnew, int *ndel) {
  int n = -1;
  int rc = guarded([&] {
    SimConfig k;
    for (int i = 0; i < 3; ++i) { k.nx[i] = c->nx[i]; k.mb[i] = c->mb[i]; }
    for (int i = 0; i < 6; ++i) k.bc[i] = c->bc[i];
    k.numlevel = c->numlevel;
    std::vector<std::array<int, 4>> in(nold), out;
    for (int b = 0; b < nold; ++b) in[b] = {leaves4[4 * b], leaves4[4 * b + 1], leaves4[4 * b + 2], leaves4[4 * b + 3]};
    const int changed = apk_host::UpdateMeshBlockTree(k, in, std::vector<int>(flags, flags + nold), &out, nnew, ndel);
    if (!changed) { n = 0; return 0; }
    if ((int)out.size() > cap) return -1;
    for (size_t b = 0; b < out.size(); ++b)
      for (int q = 0; q < 4; ++q) out4[4 * b + q] = out[b][q];
    n = (int)out.size();
    return 0;
  });
  return rc ? rc : n;
}
long long apk_sim_amr_timing(const apk_sim_t *s, double *ms8) {
  const HydroSim *h = reinterpret_cast<const HydroSim *>(s);
  for (int q = 0; q < 8; ++q) ms8[q] = h->amr_ms[q];
  return h->amr_remeshes;
}
int apk_sim_amr_tag(apk_sim_t *s) {
  return guarded([&] { SIM(s)->TagRefinement(); return 0; });
}
int apk_sim_amr_remesh(apk_sim_t *s) {
  int changed = 0;
  int rc = guarded([&] { changed = SIM(s)->LoadBalancingAndAdaptiveMeshRefinement() ? 1 : 0; return 0; });
  return rc ? rc : changed;
}
int apk_sim_amr_flag(const apk_sim_t *s, int b) {
  return reinterpret_cast<const HydroSim *>(s)->RefineFlag(b);
}
int apk_sim_enable_history(apk_sim_t *s, const char *fname, double dt, const char *data_format) {
  return guarded([&] {
    SIM(s)->EnableHistoryOutput(fname, dt, data_format ? data_format : "");
    return 0;
  });
}
int apk_sim_finalize_outputs(apk_sim_t *s) {
  return guarded([&] { SIM(s)->FinalizeOutputs(); return 0; });
}
int apk_sim_last_s_rkl(const apk_sim_t *s) { return reinterpret_cast<const HydroSim *>(s)->last_s_rkl; }

int apk_sim_nccl_unique_id(void *out, int nbytes) { return HydroSim::CommUniqueId(out, nbytes); }
int apk_sim_init_comm(apk_sim_t *s, const void *uid, int nbytes) {
  return guarded([&] { return SIM(s)->InitComm(uid, nbytes); });
}
int apk_sim_pgen(apk_sim_t *s, const char *id, const double *pars, int npars) {
  return guarded([&] {
    return SIM(s)->ProblemGenerator(id, std::vector<double>(pars, pars + (pars ? npars : 0)));
  });
}
int apk_sim_initialize(apk_sim_t *s, double tlim) {
  return guarded([&] { SIM(s)->Initialize(tlim); return 0; });
}
int apk_sim_execute(apk_sim_t *s, double tlim, int nlim) {
  int n = -1;
  int rc = guarded([&] { n = SIM(s)->Execute(tlim, nlim); return 0; });
  return rc ? rc : n;
}
int apk_sim_step(apk_sim_t *s) {
  return guarded([&] {
    HydroSim *h = SIM(s);
    h->PreStepMeshUserWorkInLoop();
    if (h->Step() != apk_host::TaskStatus::complete) return -1;
    h->tm.ncycle++;
    h->tm.time += h->tm.dt;
    if (h->Config().adaptive && h->LoadBalancingAndAdaptiveMeshRefinement()) h->EstimateTimestep(false);
    h->SetGlobalTimeStep();
    return 0;
  });
}
int apk_sim_num_local_blocks(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->NumLocalBlocks(); }
int apk_sim_num_global_blocks(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->NumGlobalBlocks(); }
int apk_sim_block_level(const apk_sim_t *s, int b) { return SIM(const_cast<apk_sim_t *>(s))->Block(b).level; }
int apk_sim_nvar(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->nvar(); }
void apk_sim_block_dims(const apk_sim_t *s, int *n3) {
  const int *n = SIM(const_cast<apk_sim_t *>(s))->dims();
  n3[0] = n[0]; n3[1] = n[1]; n3[2] = n[2];
}
void apk_sim_block_info(const apk_sim_t *s, int b, int *gid, int *loc3, double *xmin3, double *dx3) {
  const auto &B = SIM(const_cast<apk_sim_t *>(s))->Block(b);
  *gid = B.gid;
  for (int d = 0; d < 3; ++d) { loc3[d] = B.loc[d]; xmin3[d] = B.xmin[d]; dx3[d] = B.dx[d]; }
}
int apk_sim_copy_to_host(apk_sim_t *s, int field, int block, double *dst) {
  return guarded([&] { return SIM(s)->CopyToHost(field, block, dst); });
}
int apk_sim_copy_from_host(apk_sim_t *s, int field, int block, const double *src) {
  return guarded([&] { return SIM(s)->CopyFromHost(field, block, src); });
}
int apk_sim_upload_all(apk_sim_t *s, int field, const double *src) {
  return guarded([&] { return SIM(s)->UploadAll(field, src); });
}
int apk_sim_history(apk_sim_t *s, double *out8) {
  return guarded([&] { SIM(s)->History(out8); return 0; });
}
long long apk_sim_interior_elems(const apk_sim_t *s) {
  return (long long)SIM(const_cast<apk_sim_t *>(s))->InteriorElems();
}
int apk_sim_load_interior(apk_sim_t *s, const double *src) {
  return guarded([&] { return SIM(s)->LoadStateInterior(src); });
}
int apk_sim_store_interior(apk_sim_t *s, double *dst) {
  return guarded([&] { return SIM(s)->StoreStateInterior(dst); });
}
int apk_sim_download_all(apk_sim_t *s, int field, double *dst) {
  return guarded([&] { return SIM(s)->DownloadAll(field, dst); });
}
void apk_sim_enable_task_timing(apk_sim_t *s, int on) {
  SIM(s)->EnableTaskTiming(on != 0);
  if (!on) SIM(s)->ResetTaskTiming();
}
int apk_sim_task_timing(const apk_sim_t *s, const char *name, double *ms_total, long long *count) {
  return SIM(const_cast<apk_sim_t *>(s))->TaskTiming(name, ms_total, count) ? 0 : -1;
}
void *apk_sim_device_ptr(apk_sim_t *s, int field, int block) { return SIM(s)->DevicePtr(field, block); }
double apk_sim_time(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->tm.time; }
double apk_sim_dt(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->tm.dt; }
int apk_sim_ncycle(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->tm.ncycle; }
double apk_sim_c_h(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->c_h; }
long long apk_sim_fofc_corrections(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->fofc_corrections; }
long long apk_sim_fofc_redone_stages(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->fofc_redone_stages; }
int apk_sim_fused_active(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->FusedActive() ? 1 : 0; }
int apk_sim_p2p_active(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->P2PActive() ? 1 : 0; }
double apk_sim_linear_wave_period(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->LinearWavePeriod(); }
int apk_sim_linear_wave_errors(apk_sim_t *s, double *out6) {
  return guarded([&] { SIM(s)->LinearWaveErrors(out6); return 0; });
}
void *apk_sim_stream(apk_sim_t *s) { return SIM(s)->stream(); }
const void *apk_sim_pack_desc(const apk_sim_t *s) { return &SIM(const_cast<apk_sim_t *>(s))->desc(); }
const void *apk_sim_device_blocks(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->d_blocks(); }
int apk_sim_synchronize(apk_sim_t *s) {
  return guarded([&] { SIM(s)->Synchronize(); return 0; });
}
/* run individual tasks of one stage (for task-level parity tests) */
int apk_sim_task(apk_sim_t *s, const char *name, int stage, double g0, double g1, double bdt) {
  return guarded([&] {
    HydroSim *h = SIM(s);
    const std::string n(name);
    using apk_host::TaskStatus;
    TaskStatus st = TaskStatus::fail;
    if (n == "init_u1") st = h->InitU1();
    else if (n == "calc_flux") st = h->CalculateFluxes(stage);
    else if (n == "first_order_flux_correct") st = h->FirstOrderFluxCorrect(g0, g1, bdt);
    else if (n == "update") st = h->UpdateWithFluxDivergence(g0, g1, bdt);
    else if (n == "source_unsplit") st = h->AddUnsplitSources(bdt);
    else if (n == "exchange") {
      h->SendBoundBufs(0); h->ReceiveBoundBufs(); h->SetBounds(0);
      if (h->Multilevel()) h->ProlongateBounds();
      st = h->ApplyBoundaryConditions(0);
    } else if (n == "fill_derived") st = h->FillDerived();
    else if (n == "pre_step") { h->PreStepMeshUserWorkInLoop(); st = TaskStatus::complete; }
    h->Synchronize();
    return st == TaskStatus::complete ? 0 : -1;
  });
}
void apk_sim_set_c_h(apk_sim_t *s, double c_h, double mindx) { SIM(s)->c_h = c_h; SIM(s)->mindx = mindx; }
int apk_sim_num_peers(const apk_sim_t *s) { return SIM(const_cast<apk_sim_t *>(s))->NumPeers(); }
int apk_sim_peer(const apk_sim_t *s, int p, int *rank, long long *send_off, long long *send_cnt,
                 long long *recv_off, long long *recv_cnt) {
  HydroSim *h = SIM(const_cast<apk_sim_t *>(s));
  if (p < 0 || p >= h->NumPeers()) return -1;
  const auto q = h->GetPeer(p);
  *rank = q.rank; *send_off = q.send_off; *send_cnt = q.send_cnt; *recv_off = q.recv_off; *recv_cnt = q.recv_cnt;
  return 0;
}
int apk_sim_num_halo_items(const apk_sim_t *s, int recv) {
  return (int)SIM(const_cast<apk_sim_t *>(s))->HaloItems(recv != 0).size();
}
int apk_sim_halo_item(const apk_sim_t *s, int recv, int idx, apk_halo_item_t *out) {
  const auto &v = SIM(const_cast<apk_sim_t *>(s))->HaloItems(recv != 0);
  if (idx < 0 || idx >= (int)v.size()) return -1;
  *out = v[idx];
  return 0;
}
int apk_sim_push_target(const apk_sim_t *s, int block, int o1, int o2, int o3, int *rank,
                        int *local_index) {
  const int idx = (o3 + 1) * 9 + (o2 + 1) * 3 + (o1 + 1);
  return SIM(const_cast<apk_sim_t *>(s))->PushTarget(block, idx, rank, local_index) ? 0 : -1;
}
int apk_sim_neighbor(const apk_sim_t *s, int block, int o1, int o2, int o3) {
  return SIM(const_cast<apk_sim_t *>(s))->Neighbor(block, o1, o2, o3);
}

}  // extern "C"
