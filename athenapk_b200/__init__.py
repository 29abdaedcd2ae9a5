"""athenapk_b200 -- B200-native AthenaPK hot path (reconstruct -> Riemann -> flux divergence ->
Dedner source -> cons2prim + ghost exchange) behind AthenaPK's own task API.

This Python package is only the binding: it loads ``lib/libapk_b200.so`` (hand-written sm_100a
CUDA kernels + C-ABI of ``include/apk_b200.h`` + the C++ host driver mirror) with ctypes.  There
is NO CPU fallback: importing works without a GPU (so the C-ABI can be inspected), but creating
a simulation or calling a kernel without the library / a CUDA device raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# APK_B200_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("APK_B200_LIB", os.path.join(_HERE, "lib", "libapk_b200.so"))

FLUID = {"euler": 0, "glmmhd": 1}
RECON = {"dc": 0, "plm": 1, "ppm": 2, "wenoz": 3, "weno3": 4, "limo3": 5}
RIEMANN = {"none": 0, "hlle": 1, "llf": 2, "hllc": 3, "hlld": 4}
INTEGRATOR = {"rk1": 0, "rk2": 1, "vl2": 2, "rk3": 3}
BC = {"periodic": 0, "outflow": 1, "reflecting": 2}
FP_MODE = {"fast": 0, "strict": 1}
FIELD = {"cons": 0, "prim": 1, "u1": 2, "flux1": 3, "flux2": 4, "flux3": 5}


class SimConfig(C.Structure):
    """Mirror of apk_sim_config_t (include/apk_b200_sim.h)."""
    _fields_ = [
        ("fluid", C.c_int), ("recon", C.c_int), ("riemann", C.c_int), ("integrator", C.c_int),
        ("nx", C.c_int * 3), ("mb", C.c_int * 3), ("ng", C.c_int), ("bc", C.c_int * 6),
        ("xmin", C.c_double * 3), ("xmax", C.c_double * 3),
        ("gamma", C.c_double), ("cfl", C.c_double),
        ("dfloor", C.c_double), ("pfloor", C.c_double), ("efloor", C.c_double),
        ("vceil", C.c_double), ("eceil", C.c_double),
        ("fofc", C.c_int), ("dedner_extended", C.c_int), ("glmmhd_alpha", C.c_double),
        ("nscalars", C.c_int), ("fp_mode", C.c_int), ("fused", C.c_int),
        ("rank", C.c_int), ("nranks", C.c_int), ("plan_only", C.c_int),
        ("nref", C.c_int), ("ref_level", C.c_int * 4),
        ("ref_xmin", (C.c_double * 3) * 4), ("ref_xmax", (C.c_double * 3) * 4),
        ("diffint", C.c_int), ("conduction", C.c_int), ("viscosity", C.c_int),
        ("resistivity", C.c_int),
        ("thermal_diff_coeff", C.c_double), ("mom_diff_coeff", C.c_double),
        ("ohm_diff_coeff", C.c_double), ("cfl_diff", C.c_double),
        ("rkl2_max_dt_ratio", C.c_double), ("conduction_sat_phi", C.c_double),
        ("adaptive", C.c_int), ("numlevel", C.c_int), ("derefine_count", C.c_int),
        ("refine_threshold", C.c_double),
        ("refine_type", C.c_int), ("refine_threshold2", C.c_double),
    ]


REFINE_TYPE = {"pressure_gradient": 0, "xyvelocity_gradient": 1, "maxdensity": 2}


class HaloItem(C.Structure):
    """Mirror of apk_halo_item_t (include/apk_b200.h)."""
    _fields_ = [("block", C.c_int), ("lo", C.c_int * 3), ("ext", C.c_int * 3),
                ("offset", C.c_longlong)]


class PackDesc(C.Structure):
    """Mirror of apk_pack_desc_t (include/apk_b200.h)."""
    _fields_ = [
        ("nblocks", C.c_int), ("nvar", C.c_int), ("nscalars", C.c_int), ("nx", C.c_int * 3),
        ("ng", C.c_int), ("ndim", C.c_int), ("fluid", C.c_int), ("recon", C.c_int),
        ("riemann", C.c_int), ("fp_mode", C.c_int), ("gamma", C.c_double),
        ("dfloor", C.c_double), ("pfloor", C.c_double), ("efloor", C.c_double),
        ("vceil", C.c_double), ("eceil", C.c_double),
    ]


class AmrItem(C.Structure):
    """Mirror of apk_amr_item_t (include/apk_b200.h)."""
    _fields_ = [
        ("fine", C.c_void_p), ("coarse", C.c_void_p), ("nf", C.c_int * 3), ("nc", C.c_int * 3),
        ("lo", C.c_int * 3), ("hi", C.c_int * 3), ("xmin", C.c_double * 3), ("dx", C.c_double * 3),
    ]


class Block(C.Structure):
    """Mirror of apk_block_t (include/apk_b200.h)."""
    _fields_ = [
        ("cons", C.c_void_p), ("prim", C.c_void_p), ("cons_u1", C.c_void_p),
        ("flux", C.c_void_p * 3), ("dx", C.c_double * 3), ("xmin", C.c_double * 3),
        ("nbr", C.c_int * 27), ("bc", C.c_int * 6),
    ]


# every symbol include/apk_b200.h and include/apk_b200_sim.h declare
ABI_SYMBOLS = [
    "apk_last_error", "apk_abi_version", "apk_stage_fused_ex", "apk_halo_plan_create",
    "apk_halo_plan_destroy", "apk_halo_start", "apk_halo_test", "apk_halo_finish", "apk_p2p_map_peers",
    "apk_p2p_unmap", "apk_stage_barrier", "apk_allreduce_scalar", "apk_tma_maps_create",
    "apk_tma_maps_destroy", "apk_box_copy", "apk_launch_count", "apk_reset_launch_count",
    "apk_calc_fluxes", "apk_fofc_workspace_bytes", "apk_first_order_flux_correct",
    "apk_flux_div_update", "apk_dedner_source", "apk_cons_to_prim", "apk_min_dt",
    "apk_min_dt_async", "apk_fill_ghosts_local", "apk_apply_physical_bcs",
    "apk_copy_ghosts_local", "apk_halo_pack", "apk_halo_unpack", "apk_stage_fused", "apk_stage_fused_p2p",
    "apk_save_u1", "apk_fp64_probe", "apk_restrict_average", "apk_prolongate_minmod_multid",
    "apk_tag_pressure_gradient", "apk_tag_refinement", "apk_history_workspace_bytes", "apk_history_sums",
    "apk_reset_fluxes", "apk_calc_diff_fluxes", "apk_rkl2_step_first", "apk_rkl2_step_other",
    "apk_min_dt_conduction",
]
SIM_SYMBOLS = [
    "apk_sim_last_error", "apk_sim_default_config", "apk_sim_create", "apk_sim_destroy",
    "apk_sim_nccl_unique_id", "apk_sim_init_comm", "apk_sim_pgen", "apk_sim_initialize",
    "apk_sim_execute", "apk_sim_step", "apk_sim_num_local_blocks", "apk_sim_num_global_blocks",
    "apk_sim_nvar", "apk_sim_block_dims", "apk_sim_block_level", "apk_sim_block_info", "apk_sim_copy_to_host",
    "apk_sim_copy_from_host", "apk_sim_device_ptr", "apk_sim_time", "apk_sim_dt",
    "apk_sim_ncycle", "apk_sim_c_h", "apk_sim_fofc_corrections", "apk_sim_last_s_rkl", "apk_sim_ml_num_items", "apk_sim_ml_item", "apk_sim_fofc_redone_stages", "apk_sim_fused_active", "apk_sim_p2p_active",
    "apk_sim_linear_wave_period", "apk_sim_linear_wave_errors", "apk_sim_stream",
    "apk_sim_pack_desc", "apk_sim_device_blocks", "apk_sim_synchronize", "apk_sim_task",
    "apk_sim_set_c_h", "apk_sim_upload_all", "apk_sim_download_all",
    "apk_sim_interior_elems", "apk_sim_load_interior", "apk_sim_store_interior", "apk_sim_history",
    "apk_sim_enable_task_timing", "apk_sim_task_timing", "apk_sim_num_peers", "apk_sim_peer",
    "apk_sim_num_halo_items", "apk_sim_halo_item", "apk_sim_neighbor", "apk_sim_push_target",
    "apk_sim_amr_counts", "apk_sim_amr_tag", "apk_sim_amr_remesh", "apk_sim_amr_flag",
    "apk_sim_amr_update_tree", "apk_sim_amr_timing", "apk_sim_enable_history",
    "apk_sim_finalize_outputs",
]

_lib = None


def build(verbose=False):
    """Compile lib/libapk_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-j8", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libapk_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib():
    """Load the CUDA library. Raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "athenapk_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (there is no CPU fallback)" % LIB_PATH)
    # torch first: its bundled libnccl.so.2 then satisfies our NEEDED entry (one NCCL per process)
    try:
        import torch  # noqa: F401
    except Exception:
        pass
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.apk_last_error.restype = C.c_char_p
    L.apk_launch_count.restype = C.c_longlong
    L.apk_calc_fluxes.argtypes = [vp, vp, C.c_double, C.c_int, vp]
    L.apk_fofc_workspace_bytes.argtypes = [vp]
    L.apk_fofc_workspace_bytes.restype = C.c_size_t
    L.apk_first_order_flux_correct.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double,
                                               C.c_double, vp, C.POINTER(C.c_longlong), vp]
    L.apk_flux_div_update.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, vp]
    L.apk_dedner_source.argtypes = [vp, vp, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_int, vp]
    L.apk_cons_to_prim.argtypes = [vp, vp, vp, vp]
    L.apk_min_dt.argtypes = [vp, vp, C.c_double, vp, dp, vp]
    L.apk_min_dt_async.argtypes = [vp, vp, vp, vp]
    L.apk_fill_ghosts_local.argtypes = [vp, vp, C.c_uint, C.c_int, vp]
    L.apk_apply_physical_bcs.argtypes = [vp, vp, C.c_uint, C.c_int, vp]
    L.apk_copy_ghosts_local.argtypes = [vp, vp, C.c_int, vp]
    L.apk_halo_pack.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp]
    L.apk_halo_unpack.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp]
    L.apk_stage_fused.argtypes = [vp, vp, vp] + [C.c_double] * 6 + [C.c_int, C.c_int, vp, vp]
    L.apk_save_u1.argtypes = [vp, vp, vp]
    L.apk_fp64_probe.argtypes = [C.c_int, C.c_int, C.c_int, vp, dp, vp]

    L.apk_sim_last_error.restype = C.c_char_p
    L.apk_sim_default_config.argtypes = [C.POINTER(SimConfig)]
    L.apk_sim_create.argtypes = [C.POINTER(SimConfig)]
    L.apk_sim_create.restype = vp
    L.apk_sim_destroy.argtypes = [vp]
    L.apk_sim_nccl_unique_id.argtypes = [vp, C.c_int]
    L.apk_sim_init_comm.argtypes = [vp, vp, C.c_int]
    L.apk_sim_pgen.argtypes = [vp, C.c_char_p, dp, C.c_int]
    L.apk_sim_initialize.argtypes = [vp, C.c_double]
    L.apk_sim_execute.argtypes = [vp, C.c_double, C.c_int]
    L.apk_sim_step.argtypes = [vp]
    for f in ("apk_sim_num_local_blocks", "apk_sim_num_global_blocks", "apk_sim_nvar",
              "apk_sim_ncycle", "apk_sim_fused_active", "apk_sim_p2p_active", "apk_sim_synchronize"):
        getattr(L, f).argtypes = [vp]
    L.apk_sim_block_dims.argtypes = [vp, ip]
    L.apk_sim_block_level.argtypes = [vp, C.c_int]
    L.apk_sim_block_info.argtypes = [vp, C.c_int, ip, ip, dp, dp]
    L.apk_sim_copy_to_host.argtypes = [vp, C.c_int, C.c_int, vp]
    L.apk_sim_copy_from_host.argtypes = [vp, C.c_int, C.c_int, vp]
    L.apk_sim_device_ptr.argtypes = [vp, C.c_int, C.c_int]
    L.apk_sim_device_ptr.restype = vp
    for f in ("apk_sim_time", "apk_sim_dt", "apk_sim_c_h", "apk_sim_linear_wave_period"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_double
    L.apk_sim_last_s_rkl.argtypes = [vp]
    L.apk_sim_ml_num_items.argtypes = [vp]
    L.apk_sim_ml_item.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    L.apk_sim_fofc_corrections.argtypes = [vp]
    L.apk_sim_fofc_corrections.restype = C.c_longlong
    L.apk_sim_fofc_redone_stages.argtypes = [vp]
    L.apk_sim_fofc_redone_stages.restype = C.c_longlong
    L.apk_sim_linear_wave_errors.argtypes = [vp, dp]
    for f in ("apk_sim_stream", "apk_sim_pack_desc", "apk_sim_device_blocks"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = vp
    L.apk_sim_task.argtypes = [vp, C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_double]
    L.apk_sim_set_c_h.argtypes = [vp, C.c_double, C.c_double]
    L.apk_sim_upload_all.argtypes = [vp, C.c_int, vp]
    L.apk_sim_download_all.argtypes = [vp, C.c_int, vp]
    L.apk_sim_interior_elems.argtypes = [vp]
    L.apk_sim_interior_elems.restype = C.c_longlong
    L.apk_sim_load_interior.argtypes = [vp, vp]
    L.apk_sim_history.argtypes = [vp, dp]
    L.apk_sim_store_interior.argtypes = [vp, vp]
    L.apk_sim_enable_task_timing.argtypes = [vp, C.c_int]
    L.apk_sim_task_timing.argtypes = [vp, C.c_char_p, dp, C.POINTER(C.c_longlong)]
    llp = C.POINTER(C.c_longlong)
    L.apk_sim_num_peers.argtypes = [vp]
    L.apk_sim_peer.argtypes = [vp, C.c_int, ip, llp, llp, llp, llp]
    L.apk_sim_num_halo_items.argtypes = [vp, C.c_int]
    L.apk_sim_halo_item.argtypes = [vp, C.c_int, C.c_int, C.POINTER(HaloItem)]
    L.apk_sim_neighbor.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.apk_tag_refinement.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double, vp, vp, vp]
    L.apk_sim_amr_counts.argtypes = [vp, ip, ip, ip]
    L.apk_sim_amr_tag.argtypes = [vp]
    L.apk_sim_amr_remesh.argtypes = [vp]
    L.apk_sim_amr_flag.argtypes = [vp, C.c_int]
    L.apk_sim_enable_history.argtypes = [vp, C.c_char_p, C.c_double, C.c_char_p]
    L.apk_sim_finalize_outputs.argtypes = [vp]
    L.apk_sim_amr_timing.argtypes = [vp, dp]
    L.apk_sim_amr_timing.restype = C.c_longlong
    L.apk_sim_amr_update_tree.argtypes = [C.POINTER(SimConfig), C.c_int, ip, ip, C.c_int, ip, ip, ip]
    _lib = L
    return L


def make_config(fluid="euler", recon="plm", riemann="hlle", integrator="vl2", nx=(64, 32, 32),
                mb=None, ng=2, bc="periodic", xmin=(0.0, 0.0, 0.0), xmax=(3.0, 1.5, 1.5),
                gamma=5.0 / 3.0, cfl=0.3, fofc=False, dedner_extended=False, glmmhd_alpha=0.1,
                nscalars=0, dfloor=-1.0, pfloor=-1.0, efloor=-1.0, vceil=float("inf"),
                eceil=float("inf"), fp_mode="fast", fused=True, rank=0, nranks=1,
                plan_only=False, refine=(), diffint="none", conduction="none",
                thermal_diff_coeff=0.0, viscosity="none", mom_diff_coeff=0.0, resistivity="none",
                ohm_diff_coeff=0.0, cfl_diff=-1.0, rkl2_max_dt_ratio=-1.0, adaptive=None):
    """Input-deck style configuration (same keys/meaning as docs/input.md of the reference).
    refine: static refinement regions [(level, (x1min, x2min, x3min), (x1max, x2max, x3max)), ...]
    = the <parthenon/static_refinementN> blocks of a deck.
    adaptive: dict(numlevel=, threshold=, derefine_count=10) = parthenon/mesh/refinement = adaptive
    with <refinement> type = pressure_gradient."""
    c = SimConfig()
    lib().apk_sim_default_config(C.byref(c))
    c.fluid, c.recon, c.riemann = FLUID[fluid], RECON[recon], RIEMANN[riemann]
    c.integrator = INTEGRATOR[integrator]
    mb = mb or nx
    for d in range(3):
        c.nx[d], c.mb[d] = nx[d], mb[d]
        c.xmin[d], c.xmax[d] = xmin[d], xmax[d]
    c.ng = ng
    bcs = [bc] * 6 if isinstance(bc, str) else list(bc)
    for f in range(6):
        c.bc[f] = BC[bcs[f]]
    c.gamma, c.cfl = gamma, cfl
    c.fofc, c.dedner_extended, c.glmmhd_alpha = int(fofc), int(dedner_extended), glmmhd_alpha
    c.nscalars = nscalars
    c.dfloor, c.pfloor, c.efloor, c.vceil, c.eceil = dfloor, pfloor, efloor, vceil, eceil
    c.fp_mode = FP_MODE[fp_mode]
    c.fused = int(fused)
    c.rank, c.nranks = rank, nranks
    c.plan_only = int(plan_only)
    c.nref = len(refine)
    for r, (lev, lo, hi) in enumerate(refine):
        c.ref_level[r] = lev
        for d in range(3):
            c.ref_xmin[r][d], c.ref_xmax[r][d] = lo[d], hi[d]
    # <diffusion> block of a deck (docs/input.md), fixed coefficients
    c.diffint = {"none": 0, "unsplit": 1, "rkl2": 2}[diffint]
    c.conduction = {"none": 0, "isotropic": 1, "anisotropic": 2}[conduction]
    c.viscosity = {"none": 0, "isotropic": 1}[viscosity]
    c.resistivity = {"none": 0, "ohmic": 1}[resistivity]
    c.thermal_diff_coeff, c.mom_diff_coeff, c.ohm_diff_coeff = thermal_diff_coeff, mom_diff_coeff, ohm_diff_coeff
    c.cfl_diff, c.rkl2_max_dt_ratio = cfl_diff, rkl2_max_dt_ratio
    if adaptive:
        c.adaptive, c.numlevel = 1, adaptive["numlevel"]
        c.derefine_count = adaptive.get("derefine_count", 10)
        c.refine_threshold = adaptive["threshold"]
        c.refine_type = REFINE_TYPE[adaptive.get("type", "pressure_gradient")]
        c.refine_threshold2 = adaptive.get("threshold2", 0.0)
    return c


def amr_update_tree(cfg, leaves, flags):
    """Mesh::UpdateMeshBlockTree of the host driver on plain lists (no GPU needed): leaves =
    [(level, lx1, lx2, lx3)], flags = refine_flag_ per leaf.  Returns (new leaves or None when
    nothing changes, blocks created, blocks destroyed)."""
    n = len(leaves)
    a = (C.c_int * (4 * n))(*[v for l in leaves for v in l])
    f = (C.c_int * n)(*flags)
    cap = 8 * n + 64
    out = (C.c_int * (4 * cap))()
    nnew, ndel = C.c_int(), C.c_int()
    rc = lib().apk_sim_amr_update_tree(C.byref(cfg), n, a, f, cap, out, C.byref(nnew), C.byref(ndel))
    if rc < 0:
        raise RuntimeError("apk_sim_amr_update_tree: " + lib().apk_sim_last_error().decode())
    new = [tuple(out[4 * b:4 * b + 4]) for b in range(rc)] if rc else None
    return new, nnew.value, ndel.value


class HydroSim:
    """Python handle on the C++ host driver (athenapk_b200/host/hydro_sim.hpp), whose tasks
    carry the reference's names (hydro_driver.cpp:499-613)."""

    def __init__(self, **kw):
        self.L = lib()
        self.cfg = make_config(**kw)
        self.h = self.L.apk_sim_create(C.byref(self.cfg))
        if not self.h:
            raise RuntimeError("apk_sim_create: " + self.L.apk_sim_last_error().decode())
        n = (C.c_int * 3)()
        self.L.apk_sim_block_dims(self.h, n)
        self.n = tuple(n)
        self.nvar = self.L.apk_sim_nvar(self.h)
        self.shape = (self.nvar, self.n[2], self.n[1], self.n[0])

    # (adaptive meshes change both whenever the mesh does)
    nblocks = property(lambda s: s.L.apk_sim_num_local_blocks(s.h))
    nblocks_global = property(lambda s: s.L.apk_sim_num_global_blocks(s.h))

    def enable_history(self, fname, dt, data_format="%12.5e"):
        """append Parthenon-format .hst rows during execute(): <parthenon/outputN> file_type = hst"""
        self._chk(self.L.apk_sim_enable_history(self.h, fname.encode(), dt, data_format.encode()),
                  "enable_history")

    def finalize_outputs(self):
        self._chk(self.L.apk_sim_finalize_outputs(self.h), "finalize_outputs")

    def amr_counts(self):
        """(meshblocks created, destroyed, [blocks per level]) -- Mesh::nbnew / nbdel"""
        a, b, lv = C.c_int(), C.c_int(), (C.c_int * 8)()
        self.L.apk_sim_amr_counts(self.h, C.byref(a), C.byref(b), lv)
        return a.value, b.value, list(lv)

    def amr_timing(self):
        """(#remeshes, {phase: host ms summed}) of the AMR step"""
        ms = (C.c_double * 8)()
        n = self.L.apk_sim_amr_timing(self.h, ms)
        return n, dict(zip(("tag", "tree", "rebuild", "move", "fill", "rebuild_mesh",
                            "rebuild_neighbors", "rebuild_device"), list(ms)))

    def amr_tag(self):
        self._chk(self.L.apk_sim_amr_tag(self.h), "amr_tag")

    def amr_remesh(self):
        return self._chk(self.L.apk_sim_amr_remesh(self.h), "amr_remesh")

    def amr_flags(self):
        return [self.L.apk_sim_amr_flag(self.h, b) for b in range(self.nblocks)]

    def close(self):
        if getattr(self, "h", None):
            self.L.apk_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc, what):
        if rc < 0:
            raise RuntimeError("%s: %s" % (what, self.L.apk_sim_last_error().decode()))
        return rc

    def init_comm(self, uid_bytes):
        buf = C.create_string_buffer(uid_bytes, len(uid_bytes))
        self._chk(self.L.apk_sim_init_comm(self.h, buf, len(uid_bytes)), "init_comm")

    def pgen(self, name, pars=()):
        a = np.asarray(pars, dtype=np.float64)
        p = a.ctypes.data_as(C.POINTER(C.c_double)) if a.size else None
        self._chk(self.L.apk_sim_pgen(self.h, name.encode(), p, a.size), "pgen")

    def initialize(self, tlim=float("inf")):
        t = 1.0e300 if tlim == float("inf") else tlim
        self._chk(self.L.apk_sim_initialize(self.h, t), "initialize")

    def execute(self, tlim=None, nlim=-1):
        t = 1.0e300 if tlim is None else tlim
        return self._chk(self.L.apk_sim_execute(self.h, t, nlim), "execute")

    def step(self):
        self._chk(self.L.apk_sim_step(self.h), "step")

    def task(self, name, stage=1, g0=0.0, g1=1.0, bdt=0.0):
        self._chk(self.L.apk_sim_task(self.h, name.encode(), stage, g0, g1, bdt), "task " + name)

    def set_c_h(self, c_h, mindx):
        self.L.apk_sim_set_c_h(self.h, c_h, mindx)

    def block_info(self, b):
        gid = C.c_int()
        loc = (C.c_int * 3)()
        xmin = (C.c_double * 3)()
        dx = (C.c_double * 3)()
        self.L.apk_sim_block_info(self.h, b, C.byref(gid), loc, xmin, dx)
        return gid.value, tuple(loc), tuple(xmin), tuple(dx)

    def block_level(self, b):
        return self.L.apk_sim_block_level(self.h, b)

    def get(self, field, b):
        out = np.empty(self.shape, dtype=np.float64)
        self._chk(self.L.apk_sim_copy_to_host(self.h, FIELD[field], b, out.ctypes.data), "copy")
        return out

    def set(self, field, b, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        assert a.shape == self.shape
        self._chk(self.L.apk_sim_copy_from_host(self.h, FIELD[field], b, a.ctypes.data), "copy")

    def get_all(self, field):
        return np.stack([self.get(field, b) for b in range(self.nblocks)])

    def gids(self):
        return [self.block_info(b)[0] for b in range(self.nblocks)]

    def locs(self):
        return [self.block_info(b)[1] for b in range(self.nblocks)]

    def upload_all(self, field, host_ptr):
        """async H2D of a whole field ([nblocks][nvar][n3][n2][n1] doubles at host_ptr)"""
        self._chk(self.L.apk_sim_upload_all(self.h, FIELD[field], host_ptr), "upload_all")

    def history(self):
        """HydroHst volume sums of the local blocks: mass, mom1-3, KE, E, ME, relDivB"""
        out = np.zeros(8)
        self._chk(self.L.apk_sim_history(self.h, out.ctypes.data_as(C.POINTER(C.c_double))), "history")
        return out

    @property
    def interior_elems(self):
        return int(self.L.apk_sim_interior_elems(self.h))

    def load_interior(self, host_ptr):
        """async: dense interior-only cons [nblocks][nvar][nx3][nx2][nx1] -> device, ghosts, prim"""
        self._chk(self.L.apk_sim_load_interior(self.h, host_ptr), "load_interior")

    def store_interior(self, host_ptr):
        """async: device cons interiors -> dense host buffer"""
        self._chk(self.L.apk_sim_store_interior(self.h, host_ptr), "store_interior")

    def download_all(self, field, host_ptr):
        self._chk(self.L.apk_sim_download_all(self.h, FIELD[field], host_ptr), "download_all")

    def enable_task_timing(self, on=True):
        self.L.apk_sim_enable_task_timing(self.h, int(on))

    def task_timing(self, name):
        ms, cnt = C.c_double(), C.c_longlong()
        if self.L.apk_sim_task_timing(self.h, name.encode(), C.byref(ms), C.byref(cnt)) != 0:
            return 0.0, 0
        return ms.value, cnt.value

    def peers(self):
        out = []
        for p in range(self.L.apk_sim_num_peers(self.h)):
            r = C.c_int()
            v = [C.c_longlong() for _ in range(4)]
            self.L.apk_sim_peer(self.h, p, C.byref(r), *[C.byref(x) for x in v])
            out.append(dict(rank=r.value, send_off=v[0].value, send_cnt=v[1].value,
                            recv_off=v[2].value, recv_cnt=v[3].value))
        return out

    def halo_items(self, recv):
        out = []
        for i in range(self.L.apk_sim_num_halo_items(self.h, int(recv))):
            it = HaloItem()
            self.L.apk_sim_halo_item(self.h, int(recv), i, C.byref(it))
            out.append(dict(block=it.block, lo=tuple(it.lo), ext=tuple(it.ext), offset=it.offset))
        return out

    def neighbor(self, block, o1, o2, o3):
        return self.L.apk_sim_neighbor(self.h, block, o1, o2, o3)

    def push_target(self, block, o1, o2, o3):
        """(rank, local block index) that receives this block's ghost push in direction (o1,o2,o3),
        or None (physical boundary / unused dimension)"""
        r, l = C.c_int(), C.c_int()
        rc = self.L.apk_sim_push_target(self.h, block, o1, o2, o3, C.byref(r), C.byref(l))
        return None if rc != 0 else (r.value, l.value)

    def synchronize(self):
        self._chk(self.L.apk_sim_synchronize(self.h), "sync")

    time = property(lambda s: s.L.apk_sim_time(s.h))
    dt = property(lambda s: s.L.apk_sim_dt(s.h))
    ncycle = property(lambda s: s.L.apk_sim_ncycle(s.h))
    c_h = property(lambda s: s.L.apk_sim_c_h(s.h))
    fofc_corrections = property(lambda s: s.L.apk_sim_fofc_corrections(s.h))
    last_s_rkl = property(lambda s: s.L.apk_sim_last_s_rkl(s.h))

    def ml_items(self):
        """items of the multilevel exchange / flux-correction plan executed by this rank (refined
        meshes): tuples (list, exec, src, dst, lo3..., ext3...) with global block ids"""
        out = []
        buf = (C.c_int * 10)()
        for i in range(self.L.apk_sim_ml_num_items(self.h)):
            self.L.apk_sim_ml_item(self.h, i, buf)
            out.append(tuple(buf))
        return out
    fofc_redone_stages = property(lambda s: s.L.apk_sim_fofc_redone_stages(s.h))
    fused_active = property(lambda s: bool(s.L.apk_sim_fused_active(s.h)))
    p2p_active = property(lambda s: bool(s.L.apk_sim_p2p_active(s.h)))
    lw_tlim = property(lambda s: s.L.apk_sim_linear_wave_period(s.h))
    stream = property(lambda s: s.L.apk_sim_stream(s.h))

    def linear_wave_errors(self):
        """[RMS-L1, per-variable L1 ...]: 5 variables (linear_wave) or 8 (linear_wave_mhd)"""
        out = np.zeros(9)
        self._chk(self.L.apk_sim_linear_wave_errors(self.h, out.ctypes.data_as(
            C.POINTER(C.c_double))), "linear_wave_errors")
        return out


def nccl_unique_id():
    buf = C.create_string_buffer(256)
    n = lib().apk_sim_nccl_unique_id(buf, 256)
    if n <= 0:
        raise RuntimeError("ncclGetUniqueId failed")
    return buf.raw[:n]


def launch_count():
    return lib().apk_launch_count()


def reset_launch_count():
    lib().apk_reset_launch_count()
