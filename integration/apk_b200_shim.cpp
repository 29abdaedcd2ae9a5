// apk_b200_shim.cpp -- the AthenaPK-side translation unit a maintainer adds to src/hydro/ to route
// the hot path through libapk_b200.so.  It is written against the reference's own headers
// (Parthenon MeshData / MeshBlockData / StateDescriptor, AthenaPK's Hydro package Params) and is
// compile-checked in this repo (tests/test_integration_shim.py: g++ -fsyntax-only with the include
// paths of the reference's Kokkos-Serial build) -- it is NOT linked into libapk_b200.so and nothing
// of the reference is copied here: it only CALLS the reference's public interfaces.
//
// What it provides (drop-in signatures, file:line = reference):
//   B200::CalculateFluxes<recon_override>    FluxFun_t            src/hydro/hydro.hpp:43-46
//   B200::FirstOrderFluxCorrect              FOFC function        src/hydro/hydro.hpp:48-51
//   B200::UpdateWithFluxDivergence           task body            hydro_driver.cpp:534-537
//   B200::DednerSource                       GLMMHD::SourceFun_t  glmmhd/glmmhd.hpp:22
//   B200::ConsToPrim / B200::EstimateTimestep  pkg->FillDerivedMesh / EstimateTimestepMesh
//                                                                 hydro.cpp:707-713
//   B200::CalcDiffFluxes / ResetFluxes       CalcDiffFluxes (diffusion/diffusion.cpp:19-53), ResetFluxes
//   B200::RKL2StepFirst / RKL2StepOther      (hydro_driver.cpp:36-166), fixed-coefficient diffusion
//   B200::FusedStage                         calc_flux + update + source + fill_derived in one task
//   B200::StartHalo / TestHalo / FinishHalo  the three tasks of AddBoundaryExchangeTasks
//                                            (bvals/comms/boundary_communication.cpp:405-445)
//   B200::CheckRefinementMesh<criterion>     pkg->CheckRefinementMesh (state_descriptor.hpp:396-397)
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <parthenon/package.hpp>

#include "../include/apk_b200.h"
#include "eos/adiabatic_glmmhd.hpp"
#include "eos/adiabatic_hydro.hpp"
#include "hydro/diffusion/diffusion.hpp"
#include "main.hpp"

namespace Hydro::B200 {

using parthenon::MeshData;
using parthenon::Real;
using parthenon::TaskStatus;

// ---- descriptor cache: one per MeshData (rebuilt when the mesh changed: pointers move on remesh)
struct PackCache {
  apk_pack_desc_t desc{};
  parthenon::ParArray1D<apk_block_t> d_blocks;  // device copy of the per-block raw pointers
  std::vector<apk_block_t> h_blocks;
  // fused mode: second prim allocation per block (device array of pointers, stored as integers:
  // a Kokkos::View<double **> would be a rank-2 view of doubles)
  parthenon::ParArray1D<std::uintptr_t> d_prim_out;
  parthenon::ParArray1D<double> d_scalars;       // [0] dt min, fofc count in [2]
  parthenon::ParArray1D<int> d_status;
  apk_halo_plan_t *halo = nullptr;
  std::size_t mesh_version = static_cast<std::size_t>(-1);
};

inline void *Stream() {
#ifdef KOKKOS_ENABLE_CUDA
  return reinterpret_cast<void *>(Kokkos::Cuda().cuda_stream());
#else
  return nullptr;  // Serial build: only the syntax of this TU is checked
#endif
}

inline void Check(int rc) {
  if (rc != 0) PARTHENON_FAIL(apk_last_error());
}

inline int ReconId(Reconstruction r) {
  switch (r) {
  case Reconstruction::dc: return APK_RECON_DC;
  case Reconstruction::plm: return APK_RECON_PLM;
  case Reconstruction::ppm: return APK_RECON_PPM;
  case Reconstruction::wenoz: return APK_RECON_WENOZ;
  case Reconstruction::weno3: return APK_RECON_WENO3;
  case Reconstruction::limo3: return APK_RECON_LIMO3;
  default: PARTHENON_FAIL("apk_b200: unknown reconstruction");
  }
  return -1;
}

inline int RiemannId(RiemannSolver r) {
  switch (r) {
  case RiemannSolver::none: return APK_RIEMANN_NONE;
  case RiemannSolver::hlle: return APK_RIEMANN_HLLE;
  case RiemannSolver::llf: return APK_RIEMANN_LLF;
  case RiemannSolver::hllc: return APK_RIEMANN_HLLC;
  case RiemannSolver::hlld: return APK_RIEMANN_HLLD;
  default: PARTHENON_FAIL("apk_b200: unknown Riemann solver");
  }
  return -1;
}

template <class EOS>
void FillEos(const EOS &eos, apk_pack_desc_t &d) {
  d.gamma = eos.GetGamma();
  d.dfloor = eos.GetDensityFloor();
  d.pfloor = eos.GetPressureFloor();
  d.efloor = eos.GetInternalEFloor();
  d.vceil = eos.GetVelocityCeiling();
  d.eceil = eos.GetInternalECeiling();
}

// Raw device pointers of Parthenon's per-block arrays: "cons"/"prim" are [nvar][n3][n2][n1]
// ParArrayNDs (LayoutRight), "bnd_flux::cons" is [3][nvar][n3][n2][n1] with the cell extents
// (interface/metadata.cpp:174-205), the u1 register lives in mesh_data "u1".
inline PackCache &GetCache(MeshData<Real> *md, MeshData<Real> *md_u1) {
  static std::map<MeshData<Real> *, PackCache> cache;
  PackCache &c = cache[md];
  auto pmb0 = md->GetBlockData(0)->GetBlockPointer();
  auto pm = md->GetMeshPointer();
  const std::size_t version = pm->GetNumMeshBlocksThisRank() * 1000003u + pm->nbtotal;
  if (c.mesh_version == version && c.h_blocks.size() == static_cast<std::size_t>(md->NumBlocks()))
    return c;
  c.mesh_version = version;
  auto pkg = pmb0->packages.Get("Hydro");
  const int nb = md->NumBlocks();
  c.h_blocks.assign(nb, apk_block_t{});
  for (int b = 0; b < nb; ++b) {
    auto &mbd = md->GetBlockData(b);
    auto pmb = mbd->GetBlockPointer();
    apk_block_t &h = c.h_blocks[b];
    auto &cons = mbd->Get("cons").data;
    h.cons = cons.data();
    h.prim = mbd->Get("prim").data.data();
    h.cons_u1 = md_u1 ? md_u1->GetBlockData(b)->Get("cons").data.data() : nullptr;
    double *flx = mbd->Get("bnd_flux::cons").data.data();
    const std::size_t per = cons.size();
    for (int d = 0; d < 3; ++d) h.flux[d] = flx + d * per;
    h.dx[0] = pmb->coords.Dxc<1>();
    h.dx[1] = pmb->coords.Dxc<2>();
    h.dx[2] = pmb->coords.Dxc<3>();
    h.xmin[0] = pmb->coords.Xf<1>(0);
    h.xmin[1] = pmb->coords.Xf<2>(0);
    h.xmin[2] = pmb->coords.Xf<3>(0);
    for (int n = 0; n < 27; ++n) h.nbr[n] = APK_NBR_NONE;
    // same-level neighbours on this rank -> index into this pack (lookup by gid)
    for (auto &nbr : pmb->neighbors) {
      if (nbr.rank != parthenon::Globals::my_rank || nbr.origin_loc.level() != pmb->loc.level()) continue;
      const int idx = (nbr.offsets[2] + 1) * 9 + (nbr.offsets[1] + 1) * 3 + (nbr.offsets[0] + 1);
      for (int q = 0; q < nb; ++q)
        if (md->GetBlockData(q)->GetBlockPointer()->gid == nbr.gid) h.nbr[idx] = q;
    }
    // Physical boundaries are `BoundaryFlag::user` functions registered by name in Parthenon
    // (defs.hpp:133, src/bvals/boundary_conditions_apk.hpp): they stay Parthenon's
    // ApplyBoundaryConditionsOnCoarseOrFineMD task, so every face is APK_BC_BLOCK here.  (A deck
    // that wants apk_apply_physical_bcs translates its ix*_bc strings to APK_BC_OUTFLOW/REFLECT.)
    for (int f = 0; f < 6; ++f) h.bc[f] = APK_BC_BLOCK;
  }
  c.d_blocks = parthenon::ParArray1D<apk_block_t>("apk_blocks", nb);
  auto hb = Kokkos::create_mirror_view(c.d_blocks);
  for (int b = 0; b < nb; ++b) hb(b) = c.h_blocks[b];
  Kokkos::deep_copy(c.d_blocks, hb);
  c.d_scalars = parthenon::ParArray1D<double>("apk_scalars", 8);
  c.d_status = parthenon::ParArray1D<int>("apk_status", 1);

  apk_pack_desc_t &d = c.desc;
  const auto &cb = pmb0->cellbounds;
  const auto ib = cb.GetBoundsI(parthenon::IndexDomain::interior);
  const auto jb = cb.GetBoundsJ(parthenon::IndexDomain::interior);
  const auto kb = cb.GetBoundsK(parthenon::IndexDomain::interior);
  d.nblocks = nb;
  d.nscalars = pkg->Param<int>("nscalars");
  d.nvar = pkg->Param<int>("nhydro") + d.nscalars;
  d.nx[0] = ib.e - ib.s + 1;
  d.nx[1] = jb.e - jb.s + 1;
  d.nx[2] = kb.e - kb.s + 1;
  d.ng = parthenon::Globals::nghost;
  d.ndim = pm->ndim;
  const auto fluid = pkg->Param<Fluid>("fluid");
  d.fluid = (fluid == Fluid::glmmhd) ? APK_FLUID_GLMMHD : APK_FLUID_EULER;
  d.recon = ReconId(pkg->Param<Reconstruction>("reconstruction"));
  d.riemann = RiemannId(pkg->Param<RiemannSolver>("riemann"));
  d.fp_mode = APK_FP_FAST;
  if (fluid == Fluid::glmmhd) FillEos(pkg->Param<AdiabaticGLMMHDEOS>("eos"), d);
  else FillEos(pkg->Param<AdiabaticHydroEOS>("eos"), d);
  return c;
}

inline MeshData<Real> *U1(MeshData<Real> *md) {
  // hydro_driver.cpp:480-484: the u1 register of the same partition
  auto pm = md->GetMeshPointer();
  return pm->mesh_data.GetOrAdd("u1", md->partition).get();
}

// ---- FluxFun_t (hydro.hpp:43-46); RECON_OVERRIDE = APK_RECON_DC for flux_first_stage of VL2
template <int RECON_OVERRIDE>
TaskStatus CalculateFluxes(std::shared_ptr<MeshData<Real>> &md) {
  auto pkg = md->GetBlockData(0)->GetBlockPointer()->packages.Get("Hydro");
  PackCache &c = GetCache(md.get(), U1(md.get()));
  const Real c_h = (c.desc.fluid == APK_FLUID_GLMMHD) ? pkg->Param<Real>("c_h") : 0.0;
  Check(apk_calc_fluxes(&c.desc, c.d_blocks.data(), c_h, RECON_OVERRIDE, Stream()));
  return TaskStatus::complete;
}
template TaskStatus CalculateFluxes<-1>(std::shared_ptr<MeshData<Real>> &);
template TaskStatus CalculateFluxes<APK_RECON_DC>(std::shared_ptr<MeshData<Real>> &);

// ---- first_order_flux_correct_fun (hydro.hpp:48-51, hydro.cpp:1223-1342)
TaskStatus FirstOrderFluxCorrect(MeshData<Real> *u0, MeshData<Real> *u1, const Real gam0,
                                 const Real gam1, const Real beta_dt) {
  auto pkg = u0->GetBlockData(0)->GetBlockPointer()->packages.Get("Hydro");
  PackCache &c = GetCache(u0, u1);
  static std::map<MeshData<Real> *, parthenon::ParArray1D<char>> work;
  auto &w = work[u0];
  const std::size_t need = apk_fofc_workspace_bytes(&c.desc);
  if (w.size() < need) w = parthenon::ParArray1D<char>("apk_fofc_work", need);
  const Real c_h = (c.desc.fluid == APK_FLUID_GLMMHD) ? pkg->Param<Real>("c_h") : 0.0;
  long long corrected = 0;
  Check(apk_first_order_flux_correct(&c.desc, c.d_blocks.data(), gam0, gam1, beta_dt, c_h, w.data(),
                                     &corrected, Stream()));
  return TaskStatus::complete;
}

// ---- Update::UpdateWithFluxDivergence<MeshData<Real>> (interface/update.cpp:115-141)
TaskStatus UpdateWithFluxDivergence(MeshData<Real> *u0, MeshData<Real> *u1, const Real gam0,
                                    const Real gam1, const Real beta_dt) {
  PackCache &c = GetCache(u0, u1);
  Check(apk_flux_div_update(&c.desc, c.d_blocks.data(), gam0, gam1, beta_dt, Stream()));
  return TaskStatus::complete;
}

// ---- GLMMHD::SourceFun_t (glmmhd/glmmhd.hpp:22, dedner_source.cpp:17-75)
template <bool extended>
void DednerSource(MeshData<Real> *md, const Real beta_dt) {
  auto pkg = md->GetBlockData(0)->GetBlockPointer()->packages.Get("Hydro");
  PackCache &c = GetCache(md, U1(md));
  Check(apk_dedner_source(&c.desc, c.d_blocks.data(), pkg->Param<Real>("c_h"),
                          pkg->Param<Real>("mindx"), pkg->Param<Real>("glmmhd_alpha"), beta_dt,
                          extended ? 1 : 0, Stream()));
}
template void DednerSource<false>(MeshData<Real> *, const Real);
template void DednerSource<true>(MeshData<Real> *, const Real);

// ---- <diffusion>: CalcDiffFluxes (diffusion/diffusion.cpp:19-53) for the fixed-coefficient
// processes; Spitzer coefficients stay with the reference's kernels
static bool DiffusionSupported(parthenon::StateDescriptor *pkg) {
  const auto cond = pkg->Param<Conduction>("conduction");
  if (cond != Conduction::none &&
      pkg->Param<ThermalDiffusivity>("thermal_diff").GetCoeffType() != ConductionCoeff::fixed)
    return false;
  return true;
}

static apk_diffusion_t DiffusionParams(parthenon::StateDescriptor *pkg) {
  apk_diffusion_t d{};
  const auto cond = pkg->Param<Conduction>("conduction");
  if (cond != Conduction::none) {
    d.conduction = (cond == Conduction::isotropic) ? APK_CONDUCTION_ISOTROPIC : APK_CONDUCTION_ANISOTROPIC;
    d.thermal_diff_coeff = pkg->Param<ThermalDiffusivity>("thermal_diff").Get(0.0, 0.0);
    d.conduction_sat_prefac = pkg->Param<Real>("conduction_sat_prefac");
  }
  if (pkg->Param<Viscosity>("viscosity") != Viscosity::none) {
    d.viscosity = 1;
    d.mom_diff_coeff = pkg->Param<MomentumDiffusivity>("mom_diff").Get(0.0, 0.0);
  }
  if (pkg->Param<Resistivity>("resistivity") != Resistivity::none) {
    d.resistivity = 1;
    d.ohm_diff_coeff = pkg->Param<OhmicDiffusivity>("ohm_diff").Get(0.0, 0.0);
  }
  return d;
}

TaskStatus CalcDiffFluxes(parthenon::StateDescriptor *pkg, MeshData<Real> *md) {
  if (!DiffusionSupported(pkg)) return ::CalcDiffFluxes(pkg, md);
  PackCache &c = GetCache(md, U1(md));
  const apk_diffusion_t d = DiffusionParams(pkg);
  Check(apk_calc_diff_fluxes(&c.desc, c.d_blocks.data(), &d, Stream()));
  return TaskStatus::complete;
}

// ---- ResetFluxes (hydro_driver.cpp:36-91)
TaskStatus ResetFluxes(MeshData<Real> *md) {
  PackCache &c = GetCache(md, U1(md));
  Check(apk_reset_fluxes(&c.desc, c.d_blocks.data(), Stream()));
  return TaskStatus::complete;
}

// per-block "cons" pointers of an RKL2 register (MeshData "MY0" / "Yjm2", hydro_driver.cpp:211-212)
// (pointer table stored as integers: a ParArray1D<double *> would be a rank-2 view of doubles)
static double *const *RegisterPointers(MeshData<Real> *md) {
  static std::map<MeshData<Real> *, parthenon::ParArray1D<std::uintptr_t>> cache;
  auto &p = cache[md];
  if (static_cast<int>(p.size()) != md->NumBlocks()) {
    p = parthenon::ParArray1D<std::uintptr_t>("apk_reg_ptrs", md->NumBlocks());
    auto h = Kokkos::create_mirror_view(p);
    for (int b = 0; b < md->NumBlocks(); ++b)
      h(b) = reinterpret_cast<std::uintptr_t>(md->GetBlockData(b)->Get("cons").data.data());
    Kokkos::deep_copy(p, h);
  }
  return reinterpret_cast<double *const *>(p.data());
}

// ---- RKL2StepFirst (hydro_driver.cpp:93-126) incl. the Y0 <- base copy (:196-208) and
// Update::FluxDivergence(base -> MY0) (:255); md_Yjm1 is "base", md_Y0 is "u1"
TaskStatus RKL2StepFirst(MeshData<Real> *md_Y0, MeshData<Real> *md_Yjm1, MeshData<Real> *md_Yjm2,
                         MeshData<Real> *md_MY0, const int s_rkl, const Real tau) {
  PackCache &c = GetCache(md_Yjm1, md_Y0);
  Check(apk_rkl2_step_first(&c.desc, c.d_blocks.data(), RegisterPointers(md_MY0),
                            RegisterPointers(md_Yjm2), s_rkl, tau, Stream()));
  return TaskStatus::complete;
}

// ---- RKL2StepOther (hydro_driver.cpp:128-166)
TaskStatus RKL2StepOther(MeshData<Real> *md_Y0, MeshData<Real> *md_Yjm1, MeshData<Real> *md_Yjm2,
                         MeshData<Real> *md_MY0, const Real mu_j, const Real nu_j,
                         const Real mu_tilde_j, const Real gamma_tilde_j, const Real tau) {
  PackCache &c = GetCache(md_Yjm1, md_Y0);
  Check(apk_rkl2_step_other(&c.desc, c.d_blocks.data(), RegisterPointers(md_MY0),
                            RegisterPointers(md_Yjm2), mu_j, nu_j, mu_tilde_j,
                            gamma_tilde_j, tau, Stream()));
  return TaskStatus::complete;
}

// ---- pkg->FillDerivedMesh (hydro.cpp:707,712): ConservedToPrimitive over the entire range
void ConsToPrim(MeshData<Real> *md) {
  PackCache &c = GetCache(md, U1(md));
  Check(apk_cons_to_prim(&c.desc, c.d_blocks.data(), c.d_status.data(), Stream()));
}

// ---- pkg->EstimateTimestepMesh (hydro.cpp:708,713 -> :913-976): blocking reduction like the
// reference's parallel_reduce; keeps the dt_hyp Param logic of hydro.cpp:903-908
Real EstimateTimestep(MeshData<Real> *md) {
  auto pkg = md->GetBlockData(0)->GetBlockPointer()->packages.Get("Hydro");
  PackCache &c = GetCache(md, U1(md));
  const Real cfl = pkg->Param<Real>("cfl");
  double dt = 0.0;
  Check(apk_min_dt(&c.desc, c.d_blocks.data(), cfl, c.d_scalars.data(), &dt, Stream()));
  if (c.desc.fluid == APK_FLUID_GLMMHD && dt < pkg->Param<Real>("dt_hyp"))
    pkg->UpdateParam("dt_hyp", dt);
  // negative density / pressure without an active floor: the reference aborts inside ConsToPrim
  auto h_status = Kokkos::create_mirror_view_and_copy(parthenon::HostMemSpace(), c.d_status);
  if (h_status(0) == 1) PARTHENON_FAIL("Got negative density.");
  if (h_status(0) == 2) PARTHENON_FAIL("Got negative pressure.");
  return dt;
}

// ---- fused stage: replaces calc_flux -> update -> source_unsplit (+ fill_derived of the interior)
// when nothing else consumes cons.flux (hydro_driver.cpp:517-531)
TaskStatus FusedStage(MeshData<Real> *u0, MeshData<Real> *u1, const Real gam0, const Real gam1,
                      const Real beta_dt, const int stage, const int nstages, const bool vl2) {
  auto pkg = u0->GetBlockData(0)->GetBlockPointer()->packages.Get("Hydro");
  PackCache &c = GetCache(u0, u1);
  PARTHENON_REQUIRE(c.d_prim_out.size() > 0, "apk_b200: register the second prim allocation first");
  apk_stage_args_t a{};
  a.gam0 = gam0;
  a.gam1 = gam1;
  a.beta_dt = beta_dt;
  const bool mhd = c.desc.fluid == APK_FLUID_GLMMHD;
  a.c_h = mhd ? pkg->Param<Real>("c_h") : 0.0;
  a.mindx = mhd ? pkg->Param<Real>("mindx") : 1.0;
  a.alpha = mhd ? pkg->Param<Real>("glmmhd_alpha") : 0.0;
  a.recon_override = (stage == 1 && vl2) ? APK_RECON_DC : -1;
  a.flags = APK_FUSED_PUSH_PRIM | (stage == nstages ? APK_FUSED_PUSH_CONS : 0) |
            (vl2 ? (stage == 1 ? APK_FUSED_PRIM_ONLY : APK_FUSED_U1_FROM_CONS)
                 : (stage == 1 ? APK_FUSED_SAVE_U1 : 0));
  a.d_dtmin = (stage == nstages) ? c.d_scalars.data() : nullptr;
  a.d_status = c.d_status.data();
  Check(apk_stage_fused_ex(&c.desc, c.d_blocks.data(),
                           reinterpret_cast<double *const *>(c.d_prim_out.data()), &a, Stream()));
  return TaskStatus::complete;
}

// ---- the three exchange tasks (boundary_communication.cpp:405-445) for blocks on other GPUs.
// The plan (slab lists per peer, apk_halo_plan_create) is built from pmb->neighbors after every
// remesh, next to Parthenon's own BuildBoundaryBuffers.
TaskStatus StartHalo(std::shared_ptr<MeshData<Real>> &md) {
  PackCache &c = GetCache(md.get(), U1(md.get()));
  if (c.halo) Check(apk_halo_start(c.halo, &c.desc, c.d_blocks.data(), 0, Stream()));
  return TaskStatus::complete;
}
TaskStatus TestHalo(std::shared_ptr<MeshData<Real>> &md) {
  PackCache &c = GetCache(md.get(), U1(md.get()));
  const int st = c.halo ? apk_halo_test(c.halo) : 0;
  if (st > 1) PARTHENON_FAIL(apk_last_error());
  return st == 1 ? TaskStatus::incomplete : TaskStatus::complete;  // polled like ReceiveBoundBufs
}
TaskStatus FinishHalo(std::shared_ptr<MeshData<Real>> &md) {
  PackCache &c = GetCache(md.get(), U1(md.get()));
  if (c.halo) Check(apk_halo_finish(c.halo, &c.desc, c.d_blocks.data(), 0, Stream()));
  Check(apk_copy_ghosts_local(&c.desc, c.d_blocks.data(), 0, Stream()));
  return TaskStatus::complete;
}

// ---- refinement tagging: pkg->CheckRefinementMesh (interface/state_descriptor.hpp:396-397, called
// by Refinement::CheckAllRefinement(md), amr_criteria/refinement_package.cpp:53-68) instead of the
// per-block CheckRefinementBlock functions of src/refinement (hydro.cpp:788-816): ONE launch tags
// every block of the MeshData.  AmrTag is an int-sized enum (basic_types.hpp:64), so the kernel
// writes the tag array in place; Parthenon's scatter-max semantics (tags start at `derefine`, a
// criterion can only raise them) hold because this is the only criterion of the Hydro package.
template <int CRITERION>
void CheckRefinementMesh(MeshData<Real> *md, parthenon::ParArray1D<parthenon::AmrTag> &delta_level) {
  PackCache &c = GetCache(md, nullptr);
  auto pkg = md->GetBlockData(0)->GetBlockPointer()->packages.Get("Hydro");
  static_assert(sizeof(parthenon::AmrTag) == sizeof(int), "AmrTag is stored as int");
  double above, below;
  if (CRITERION == APK_REFINE_PRESSURE_GRADIENT) {
    above = pkg->Param<Real>("refinement/threshold_pressure_gradient");
    below = 0.25 * above;  // gradient.cpp:59
  } else if (CRITERION == APK_REFINE_XYVELOCITY_GRADIENT) {
    above = pkg->Param<Real>("refinement/threshold_xyvelocity_gradient");
    below = 0.5 * above;   // gradient.cpp:92
  } else {
    above = pkg->Param<Real>("refinement/maxdensity_refine_above");
    below = pkg->Param<Real>("refinement/maxdensity_deref_below");
  }
  parthenon::ParArray1D<double> measure("apk_refinement_measure", md->NumBlocks());
  Check(apk_tag_refinement(&c.desc, c.d_blocks.data(), CRITERION, above, below,
                           reinterpret_cast<int *>(delta_level.data()), measure.data(), Stream()));
}

// ---- registration (called from Hydro::Initialize instead of the flux_functions lookup,
// hydro.cpp:381-480, and the FillDerivedMesh / EstimateTimestepMesh assignments :707-713)
void Register(parthenon::StateDescriptor *pkg, const bool vl2, const bool extended_dedner) {
  using FluxFun_t = TaskStatus(std::shared_ptr<MeshData<Real>> &);
  FluxFun_t *first = vl2 ? &CalculateFluxes<APK_RECON_DC> : &CalculateFluxes<-1>;
  FluxFun_t *other = &CalculateFluxes<-1>;
  pkg->AddParam<FluxFun_t *>("apk_b200_flux_first_stage", first);
  pkg->AddParam<FluxFun_t *>("apk_b200_flux_other_stage", other);
  pkg->FillDerivedMesh = ConsToPrim;
  pkg->EstimateTimestepMesh = EstimateTimestep;
  using SourceFun_t = std::function<void(MeshData<Real> *, const Real)>;
  SourceFun_t src = extended_dedner ? SourceFun_t(DednerSource<true>) : SourceFun_t(DednerSource<false>);
  pkg->AddParam<SourceFun_t>("apk_b200_glmmhd_source", src);
  // <refinement> type (hydro.cpp:788-816): the mesh-level hook replaces CheckRefinementBlock
  if (pkg->AllParams().hasKey("refinement/threshold_pressure_gradient")) {
    pkg->CheckRefinementBlock = nullptr;
    pkg->CheckRefinementMesh = CheckRefinementMesh<APK_REFINE_PRESSURE_GRADIENT>;
  } else if (pkg->AllParams().hasKey("refinement/threshold_xyvelocity_gradient")) {
    pkg->CheckRefinementBlock = nullptr;
    pkg->CheckRefinementMesh = CheckRefinementMesh<APK_REFINE_XYVELOCITY_GRADIENT>;
  } else if (pkg->AllParams().hasKey("refinement/maxdensity_refine_above")) {
    pkg->CheckRefinementBlock = nullptr;
    pkg->CheckRefinementMesh = CheckRefinementMesh<APK_REFINE_MAXDENSITY>;
  }
}

}  // namespace Hydro::B200
