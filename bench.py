#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native AthenaPK hot path.

Metric (BASELINE.json): cell-updates/s (== Parthenon zone-cycles/s, driver.cpp:53-64) of uniform
256^3 GLM-MHD, VL2 + PPM + HLLD, 32^3 meshblocks, on 1/2/4/8 B200 (strong scaling: the 256^3 mesh
is split over the ranks; ghost zones of remote neighbours are peer stores of the stage kernel over
CUDA-IPC mapped NVLink memory -- `config.halo` says which exchange ran; NCCL carries the 4-byte stage
barrier and the dt min).  One "step" = one full cycle (both VL2 stages, Dedner source, cons->prim,
ghost exchange, dt reduction).

  python bench.py [--gpus N] [--steps K] [--warmup W]      # product arm (this repo's CUDA path)
  python bench.py --impl reference ...                      # reference arm: the UNMODIFIED reference,
                                                            # Kokkos-Serial build, one copy per core
  python bench.py --config ot2d                             # BASELINE config 3 (Orszag-Tang, FOFC on)
  python bench.py --config blast_amr --steps 40             # BASELINE config 5 (adaptive AMR blast)

Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for the byte model.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell-updates/sec at 256^3 MHD (PPM+HLLD, VL2)"
UNIT = "cell-updates/s"
ALG_BYTES_PER_UPDATE = 504.0       # MHD nv=9, 2-stage: 216 + 288 B
ALG_BYTES_STAGE = {"fused_stage_first": 216.0, "fused_stage_other": 288.0}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md): one
    long-running `nvidia-smi -lms 100` child, parsed after the region."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
        self.t0 = time.time()

    def start(self):
        time.sleep(0.25)  # let the first sample land before the timed region starts
        return self

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.12)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        samples, reasons, sm_max = [], set(), None
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            try:
                samples.append(float(f[0]))
                sm_max = float(f[1])
                for n, v in zip(self.NAMES, f[2:]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        s = sorted(samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": sm_max,
                "reasons": sorted(reasons), "samples": len(s)}


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, BEFORE the pinned
    staging buffers are allocated (first touch places them on that NUMA node).  Without it the
    H2D/D2H legs of the e2e number of N ranks share one socket's memory controllers (measured:
    14 GB/s per rank and direction at N=8 instead of 46 GB/s at N=1)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))[:1] + [len(os.sched_getaffinity(0))]
    except Exception:
        return None


def workload(args):
    n = args.n
    if args.config == "ot2d":  # BASELINE config 3: inputs/orszag_tang.in scaled up, FOFC on
        return dict(fluid="glmmhd", recon=args.recon, riemann="hlld", integrator="vl2", nx=(n, n, 1),
                    mb=(args.mb, args.mb, 1), ng=3, xmin=(-0.5, -0.5, -0.5), xmax=(0.5, 0.5, 0.5),
                    gamma=5.0 / 3.0, cfl=0.4, fofc=True)
    kw = dict(fluid="glmmhd", recon=args.recon, riemann="hlld", integrator="vl2", nx=(n, n, n),
              mb=(args.mb,) * 3, ng=3, xmin=(0.0, 0.0, 0.0), xmax=(1.0, 1.0, 1.0),
              gamma=5.0 / 3.0, cfl=0.3)
    if args.pgen == "uniform":       # the deck's own state and gamma (turbulence.in:63-69)
        kw["gamma"] = 1.0001
    if args.pgen == "orszag_tang":   # baseline/decks/orszag_tang_3d.in
        kw.update(xmin=(-0.5, -0.5, -0.5), xmax=(0.5, 0.5, 0.5))
    return kw


REF_DECK = {"orszag_tang": "orszag_tang_3d", "uniform": "uniform_mhd"}


def cpu_sample(cycles, n=64, mb=32):
    """time the CPU oracle (C port of the reference algorithm, single thread) on a bounded sample --
    only used when the reference binary did not travel (baseline/_ref/serial missing)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    o = Oracle(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(n, n, n),
               mb=(mb,) * 3, ng=3, xmin=(-0.5, -0.5, -0.5), xmax=(0.5, 0.5, 0.5), gamma=5.0 / 3.0,
               cfl=0.3)
    o.pgen("orszag_tang")
    o.initialize()
    o.run(None, 1)  # warm-up cycle
    t0 = time.time()
    o.run(None, 1 + cycles)
    dt = time.time() - t0
    return n ** 3 * cycles / dt


def _cpu_worker(q, cycles):
    q.put(cpu_sample(cycles))


def reference_serial(pgen, copies, cycles, skip, n=64, mb=32):
    """The UNMODIFIED reference (Kokkos-Serial build, baseline/_ref/serial) on the product arm's
    deck: `copies` independent runs pinned to cores 0..copies-1 (the Serial backend is single
    threaded and there is no MPI in the image), each a bounded n^3 sample with the product arm's
    meshblock size and method.  Returns (aggregate zone-cycles/s over the cycles after `skip`,
    mean wall seconds per cycle) or None when the binary is not there."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import run_reference as rr
    if pgen not in REF_DECK or not rr.available("serial"):
        return None
    agg, rates = rr.run_copies(REF_DECK[pgen], n, mb, skip + cycles, copies, skip)
    if not agg:
        return None
    per_cycle = sum((n ** 3) / r for r in rates if r > 0) / max(1, sum(1 for r in rates if r > 0))
    return agg, per_cycle


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path -- the unmodified
    reference binary (Kokkos Serial) on the same deck as the product arm, one copy per host core,
    aggregate zone-cycles/s (= cell-updates/s).  One step = one cycle of every copy; the copies
    run `warmup + steps` cycles on 64^3 (8 meshblocks of 32^3; the per-cell cost of the reference
    does not depend on the mesh size) so the run ends within a minute."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 64)
    ref = reference_serial(args.pgen, cores, args.steps, args.warmup) if args.config == "uni256" else None
    if ref is not None:
        value, sec = ref
        kind = "reference"
        sample = ("%d pinned copies (one per core) of baseline/_ref/serial/bin/athenaPK, deck "
                  "baseline/decks/%s.in at 64^3 (8 meshblocks of 32^3), VL2+PPM+HLLD, %d cycles "
                  "timed after %d warm-up cycles; aggregate of Parthenon's own zone-cycles/wsec" %
                  (cores, REF_DECK[args.pgen], args.steps, args.warmup))
    else:  # the binary did not travel: oracle port, one replica per core
        import multiprocessing as mp
        vals = []
        for step in range(2):
            q = mp.Queue()
            ps = [mp.Process(target=_cpu_worker, args=(q, 2)) for _ in range(cores)]
            t0 = time.time()
            for p in ps:
                p.start()
            for p in ps:
                p.join()
            vals.append((sum(q.get() for _ in ps), time.time() - t0))
        value, sec = vals[-1][0], vals[-1][1] / 3
        kind = "port"
        sample = ("%d replicas (one per core) of the oracle port on 64^3 orszag_tang, 2 timed "
                  "cycles each (reference binary not present)" % cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, None),
                       "sample": "bounded 64^3 samples of the same deck and method, one per core"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(args, sim):
    if args.config == "ot2d":
        return ("Orszag-Tang %d^2 GLM-MHD (inputs/orszag_tang.in scaled up), VL2+%s+HLLD+FOFC, "
                "meshblock %d^2, nghost 3, cfl 0.4" % (args.n, args.recon.upper(), args.mb))
    what = {"orszag_tang": "Orszag-Tang vortex state (baseline/decks/orszag_tang_3d.in)",
            "uniform": "static deck state (baseline/decks/uniform_mhd.in)",
            "ot3d": "closed-form ot3d state", "random": "white noise"}[args.pgen]
    return ("uniform-grid %d^3 GLM-MHD, turbulence.in geometry without forcing, %s, VL2+%s+HLLD, "
            "meshblock %d^3, %d meshblocks, nghost 3, cfl 0.3" %
            (args.n, what, args.recon.upper(), args.mb, (args.n // args.mb) ** 3))


def run_product(args):
    import torch
    import torch.distributed as dist
    import athenapk_b200 as apk

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kw = workload(args)
    sim = apk.HydroSim(fp_mode=args.fp_mode, fused=not args.unfused, rank=rank, nranks=world, **kw)
    if world > 1:
        uid = [apk.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sim.init_comm(uid[0])
    pgen_pars = (3, 0.2) if args.pgen == "random" else ()
    sim.pgen(args.pgen, pgen_pars)
    sim.initialize()
    cells = float(kw["nx"][0]) * kw["nx"][1] * kw["nx"][2]
    stream = torch.cuda.ExternalStream(sim.stream)

    # ---------------- device-resident throughput ("value") ----------------
    sim.execute(None, sim.ncycle + args.warmup)
    barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    sim.enable_task_timing(False)
    apk.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    sim.execute(None, sim.ncycle + args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = apk.launch_count()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = cells * args.steps / (ms * 1e-3)

    # ---------------- per-kernel timing for the roofline (separate pass, events per task) ----
    sim.enable_task_timing(True)
    sim.execute(None, sim.ncycle + max(2, min(args.steps, 5)))
    names = ["fused_stage_first", "fused_stage_other", "send", "set", "fbound", "new_dt", "init_u1",
             "calc_flux", "update", "source_unsplit", "fill_derived", "first_order_flux_correct"]
    tt = {n: sim.task_timing(n) for n in names}
    sim.enable_task_timing(False)
    peak, peak_kind = peaks()
    local_cells = cells / world
    roof = None
    dom = "fused_stage_other" if tt["fused_stage_other"][1] else "calc_flux"
    if tt[dom][1]:
        dur_ms = tt[dom][0] / tt[dom][1]
        bytes_launch = ALG_BYTES_STAGE.get(dom, 288.0) * local_cells
        achieved = bytes_launch / (dur_ms * 1e-3) / 1e9
        ndim = 3 if kw["nx"][2] > 1 else 2
        roof = {"bound": "hbm", "kernel": ("k_stage_fused<glmmhd,%s,hlld,%dd> (stage 2)" % (args.recon, ndim))
                if dom == "fused_stage_other" else dom, "achieved": achieved, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "traffic_source": None, "alg_bytes_per_launch": bytes_launch,
                "launch_ms": dur_ms}
        # ncu dram__bytes of ONE launch of this kernel: a static figure from profiles/ (ncu cannot
        # run inside the timed bench); only valid for the configuration it was captured on
        tf = os.path.join(ROOT, "profiles", "traffic_r02.json")
        if (os.path.exists(tf) and world == 1 and args.config == "uni256" and args.n == 256 and
                args.recon == "ppm" and dom == "fused_stage_other"):
            try:
                tj = json.load(open(tf))
                roof["traffic"] = tj.get("dram_bytes_per_launch_256")
                roof["traffic_source"] = "static: " + tj.get("source", "profiles/traffic_r02.json")
            except Exception:
                pass
    task_ms = {n: (v[0] / v[1] if v[1] else None) for n, v in tt.items() if v[1]}
    # FP64 context for the roofline: the kernel is bound by the half-rate FP64 pipe, not by HBM
    # (DESIGN.md 3.1).  Live DFMA peak of this GPU from the library's probe kernel, and the FP64 /
    # issue utilisation ncu measured for the same kernel (profiles/ncu_r01_final_256.csv).
    if roof is not None and rank == 0:
        try:
            import ctypes as C
            L = apk.lib()
            scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
            flops = C.c_double()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 0.0
            for it in range(3):
                a.record(stream)
                L.apk_fp64_probe(148 * 8, 256, 20000, C.c_void_p(scratch.data_ptr()), C.byref(flops),
                                 C.c_void_p(sim.stream))
                b.record(stream)
                torch.cuda.synchronize()
                best = max(best, flops.value / (a.elapsed_time(b) * 1e-3) / 1e12)
            roof["fp64_dfma_peak_tflops_measured"] = best
        except Exception:
            pass
        # FP64-pipe roofline: (FP64 instructions per cell of this kernel, a static ncu count) x
        # cells per launch / live launch time, against the live DFMA issue rate of this GPU
        tf = os.path.join(ROOT, "profiles", "traffic_r02.json")
        try:
            tj = json.load(open(tf))
            per_cell = tj["fp64_thread_instr_per_cell"].get(dom if args.recon == "ppm" else "", None)
            if per_cell and roof.get("fp64_dfma_peak_tflops_measured"):
                rate = roof["fp64_dfma_peak_tflops_measured"] * 1e12 / 2.0  # DFMA = 2 flop
                roof["fp64_frac"] = per_cell * local_cells / (dur_ms * 1e-3) / rate
                roof["fp64_frac_source"] = ("live launch time and live DFMA rate; FP64 instructions "
                                            "per cell static from " + tj.get("source", "profiles/"))
        except Exception:
            pass
        roof["note"] = ("HBM is the contract roofline; the kernel is bound by FP64 issue and "
                        "dependency latency (fp64_frac = share of the measured DFMA issue rate)")

    # state fingerprint for the parity check below, BEFORE the e2e phase reloads host states
    total_cycles = sim.ncycle
    hist = torch.tensor(sim.history()[:7], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    hist = hist.cpu().numpy()

    # ---------------- end-to-end through the public API with HOST buffers ----------------
    # Every step: pinned-host conserved state (ghosts included) -> H2D -> FillDerived (cons->prim)
    # -> one full cycle -> D2H of the updated conserved state into a pinned host buffer.
    # Two simulation contexts alternate so that the H2D of step i+1 and the D2H of step i-1 overlap
    # the compute of step i (independent steps of a throughput run; PCIe is full duplex).  The
    # strictly serial variant (one context, copy-compute-copy back to back) is reported beside it.
    e2e = None
    if not args.no_e2e:
        nel = sim.interior_elems
        h_in = torch.empty(nel, dtype=torch.float64).pin_memory()
        h_out = [torch.empty(nel, dtype=torch.float64).pin_memory() for _ in range(3)]
        sim.store_interior(h_in.data_ptr())
        sim.synchronize()
        k2 = max(2, min(args.steps, 6))

        def one_step(x, out):
            x.step()                       # one full cycle (blocks the host on the dt reduction)
            x.store_interior(out.data_ptr())  # async gather + D2H

        # serial: copy in, compute, copy out, wait
        for it in range(1 + k2):
            if it == 1:
                barrier()
                t0 = time.time()
            sim.load_interior(h_in.data_ptr())
            one_step(sim, h_out[0])
            sim.synchronize()
        barrier()
        serial_ms = (time.time() - t0) * 1e3 / k2
        # pipelined over three contexts (round robin): the H2D of step i+1 and the D2H of step i-1
        # run on other contexts' streams while step i computes; within one context load, cycle and
        # store are stream-ordered, so two contexts would serialise a D2H with the next H2D
        sims = [sim]
        for _ in range(2):
            sx = apk.HydroSim(fp_mode=args.fp_mode, fused=not args.unfused, rank=rank, nranks=world, **kw)
            if world > 1:
                uid = [apk.nccl_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                sx.init_comm(uid[0])
            sx.pgen(args.pgen, pgen_pars)
            sx.initialize()
            sims.append(sx)
        nctx = len(sims)
        k2 = max(3, min(args.steps, 9))
        sims[0].load_interior(h_in.data_ptr())
        for it in range(nctx + k2):
            if it == nctx:   # pipeline is in steady state: start the clock
                for x in sims:
                    x.synchronize()
                barrier()
                sims[it % nctx].load_interior(h_in.data_ptr())
                sims[it % nctx].synchronize()
                t0 = time.time()
            x, y = sims[it % nctx], sims[(it + 1) % nctx]
            y.load_interior(h_in.data_ptr())   # next step's input: overlaps this step
            one_step(x, h_out[it % 3])
        for x in sims:
            x.synchronize()
        barrier()
        pipe_ms = (time.time() - t0) * 1e3 / k2
        t = torch.tensor([pipe_ms, serial_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pipe_ms, serial_ms = float(t[0]), float(t[1])
        e2e = {"value": cells / (pipe_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(nel * 8 * world), "d2h_bytes_per_step": int(nel * 8 * world),
               "ms_per_step": pipe_ms, "serial_value": cells / (serial_ms * 1e-3),
               "serial_ms_per_step": serial_ms, "serial_over_pipelined": pipe_ms / serial_ms,
               "what": "every step: pinned-host conserved state (interior cells, 72 B/cell) -> H2D -> "
                       "scatter + ghost exchange + FillDerived -> 1 full cycle -> gather -> D2H; "
                       "value = three contexts round robin (copies of neighbouring steps overlap compute), "
                       "serial_value = one context, copy-compute-copy back to back; host wall clock; "
                       "PCIe-bound (1.2 GB each way per step at 256^3)"}
        for sx in sims[1:]:
            sx.close()

    # ---------------- strict-FP throughput and a live parity check -------------------------------
    # N = 1: a second context in APK_FP_STRICT mode (the bit-exact arm of the parity tests) runs the
    #        same cycles: its throughput, and the relative difference of the history sums (mass,
    #        momenta, KE, E, ME) between the FMA mode that was timed and the strict mode.
    # N > 1: rank 0 repeats the same cycles on ONE GPU; the all-reduced history sums of the
    #        partitioned run must agree with it to 1e-13 (the data path is bit-identical; only the
    #        summation order of the reducers differs).
    parity, strict = None, None
    if rank == 0 and not args.no_parity and args.config == "uni256":
        import numpy as np
        chk = apk.HydroSim(fp_mode="strict" if world == 1 else args.fp_mode, fused=not args.unfused,
                           **kw)
        chk.pgen(args.pgen, pgen_pars)
        chk.initialize()
        chk.execute(None, args.warmup)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cstream = torch.cuda.ExternalStream(chk.stream)
        c0.record(cstream)
        chk.execute(None, total_cycles)
        c1.record(cstream)
        torch.cuda.synchronize()
        ref_hist = np.array(chk.history()[:7])
        scale = np.maximum(np.abs(ref_hist), np.abs(ref_hist[[0, 4, 4, 4, 4, 5, 6]]))
        rel = float(np.max(np.abs(hist - ref_hist) / np.maximum(scale, 1e-300)))
        if world == 1:
            strict = {"value": cells * (total_cycles - args.warmup) / (c0.elapsed_time(c1) * 1e-3),
                      "unit": UNIT, "fp_mode": "strict",
                      "what": "same workload with the -fmad=false kernels (bit-identical to the "
                              "reference's CPU build); %d cycles" % (total_cycles - args.warmup)}
            parity = {"vs": "APK_FP_STRICT run of the same %d cycles (history sums mass, momenta, KE, "
                            "E, ME)" % total_cycles, "max_rel_diff": rel, "tol": 1e-6, "ok": rel <= 1e-6}
        else:
            parity = {"vs": "single-GPU run of the same %d cycles on rank 0 (history sums, all-reduced)"
                            % total_cycles, "max_rel_diff": rel, "tol": 1e-13, "ok": rel <= 1e-13}
        chk.close()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu and args.config == "uni256":
            t0 = time.time()
            ref = reference_serial(args.pgen, 1, args.cpu_cycles, 2)
            if ref is not None:
                cpu = {"value": ref[0], "unit": UNIT, "cores": 1, "kind": "reference",
                       "sample": "baseline/_ref/serial/bin/athenaPK (unmodified reference, Kokkos "
                                 "Serial), deck baseline/decks/%s.in at 64^3 (8 meshblocks of 32^3), "
                                 "VL2+PPM+HLLD, %d cycles after 2 warm-up cycles, 1 core, %.1f s"
                                 % (REF_DECK[args.pgen], args.cpu_cycles, time.time() - t0)}
            else:
                v = cpu_sample(args.cpu_cycles)
                cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                       "sample": "oracle (C port of the reference's Kokkos-Serial path), 64^3 "
                                 "orszag_tang, %d timed cycles, 1 thread, %.1f s (reference binary "
                                 "not present or deck not runnable by it)"
                                 % (args.cpu_cycles, time.time() - t0)}
        halo = ("ipc-peer-store (CUDA-IPC mapped NVLink stores inside the stage kernel; NCCL: 4-byte "
                "stage barrier + dt min)" if sim.p2p_active else
                ("nccl-packed send/recv" if world > 1 else "same-device in-kernel push"))
        static_ref = None
        try:
            rj = json.load(open(os.path.join(ROOT, "profiles", "reference_cuda_vs_b200_r02.json")))
            key = REF_DECK.get(args.pgen)
            if key in rj and args.config == "uni256" and args.n == 256:
                static_ref = {"value": rj[key]["reference_cuda_zone_cycles_per_s"], "unit": UNIT,
                              "source": "static: profiles/reference_cuda_vs_b200_r02.json -- the "
                                        "unmodified reference's Kokkos-CUDA sm_100 build on one B200, "
                                        "same deck, 256^3 (measured in round 2)"}
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args, sim), "pgen": args.pgen,
                           "fp_mode": args.fp_mode, "fused": sim.fused_active, "halo": halo,
                           "l2": "state (%.1f GB/GPU) >> 126 MB L2; no flush needed" %
                                 (sim.nblocks * sim.nvar * sim.n[0] * sim.n[1] * sim.n[2] * 8 * 4 / 1e9),
                           "partition": "Morton-ordered meshblock list split in %d contiguous ranges" % world},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks, "task_ms": task_ms,
                "strict": strict, "parity_check": parity, "reference_cuda": static_ref,
                "fofc_redone_stages": int(sim.fofc_redone_stages),
                "host_affinity": {"first_cpu_and_count": numa},
                "p2p_halo": bool(sim.p2p_active),
                "alg_bytes_per_cell_update": ALG_BYTES_PER_UPDATE,
                "hbm_frac_whole_cycle": value * ALG_BYTES_PER_UPDATE / (peak * 1e9 * world)}
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def run_blast_amr(args):
    """--config blast_amr: BASELINE config 5 as the reference runs it (baseline/decks/blast_amr.in:
    inputs/blast_3d_amr.in with refinement = adaptive, GLM-MHD WENOZ + HLLD, nghost 4, 3 levels) on
    ONE GPU.  A parity-test configuration, not the headline; metric = Parthenon's zone-cycles per
    wall second including tagging and remeshing, with the unmodified reference beside it: its
    Kokkos-Serial build timed here on one host core (bounded: 6 cycles) and its Kokkos-CUDA build on
    the same B200 (static record, profiles/reference_cuda_vs_b200_config5_r02.json)."""
    import torch
    import athenapk_b200 as apk
    assert args.gpus == 1 and int(os.environ.get("WORLD_SIZE", "1")) == 1, "blast_amr runs on one GPU"
    torch.cuda.set_device(0)
    g = float("1.666666666666667")
    kw = dict(fluid="glmmhd", recon="wenoz", riemann="hlld", integrator="vl2", nx=(32, 32, 32),
              mb=(8, 8, 8), ng=4, xmin=(-0.5,) * 3, xmax=(0.5,) * 3, gamma=g, cfl=0.3,
              adaptive=dict(numlevel=3, threshold=0.1, derefine_count=10))
    sim = apk.HydroSim(fp_mode=args.fp_mode, **kw)
    sim.pgen("blast", (0.12, 0.1, 100.0, 1.0, 0.22, 0.1, -0.05))
    sim.initialize(0.1)
    nb0 = sim.nblocks
    sim.execute(0.1, args.warmup)
    sim.synchronize()
    apk.reset_launch_count()
    n_pre, ms_pre = sim.amr_timing()
    sampler = ClockSampler(0).start()
    zones, t0 = 0, time.time()
    for c in range(args.warmup, args.warmup + args.steps):
        zones += sim.nblocks * 512
        sim.execute(0.1, c + 1)
    sim.synchronize()
    sec = time.time() - t0
    clocks = sampler.stop()
    launches = apk.launch_count()
    n_amr, ms_amr = sim.amr_timing()
    created, destroyed, per_level = sim.amr_counts()
    value = zones / sec
    line = {"metric": "zone-cycles/sec, adaptive AMR blast (BASELINE config 5: WENOZ+HLLD, 3 levels)",
            "value": value, "unit": "zone-cycles/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3 / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "baseline/decks/blast_amr.in: 32^3 root grid, 8^3 meshblocks, 3 levels, "
                                   "refinement = adaptive (pressure gradient), GLM-MHD VL2+WENOZ+HLLD, "
                                   "nghost 4; tagging + remeshing inside the timed region (host wall clock)",
                       "fp_mode": args.fp_mode, "fused": False, "blocks_start": nb0,
                       "blocks_end": sim.nblocks, "created": created, "destroyed": destroyed,
                       "per_level": per_level[:3], "remeshes": n_amr - n_pre,
                       "l2": "state (200 MB) > 126 MB L2"},
            "amr_host_ms_per_cycle": {k: (ms_amr[k] - ms_pre[k]) / args.steps for k in ms_amr},
            "gpu_launches": launches, "clocks": clocks, "roofline": None, "e2e": None}
    ref = os.path.join(ROOT, "profiles", "reference_cuda_vs_b200_config5_r02.json")
    if os.path.exists(ref):
        r = json.load(open(ref))["reference_kokkos_cuda_sm100"]
        line["reference_cuda"] = {"value": r["zone_cycles_per_s_steady"], "unit": "zone-cycles/s",
                                  "source": "static: profiles/reference_cuda_vs_b200_config5_r02.json"}
    sim.close()
    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import run_reference as rr
        if rr.available("serial"):
            res = rr.run("serial", "blast_amr", 32, 8, 6, cpu=0)
            line["cpu_baseline"] = {"value": res["zone_cycles_per_wallsecond"], "unit": "zone-cycles/s",
                                    "cores": 1, "kind": "reference",
                                    "sample": "baseline/_ref/serial/bin/athenaPK on the same deck, 6 "
                                              "cycles, 1 core, %.0f s" % res["wall_s"]}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--n", type=int, default=256, help="mesh cells per direction")
    ap.add_argument("--mb", type=int, default=32, help="meshblock cells per direction")
    ap.add_argument("--pgen", default="orszag_tang", choices=["orszag_tang", "uniform", "ot3d", "random"],
                    help="orszag_tang (default) and uniform are decks the reference runs too "
                         "(baseline/decks); ot3d / random are closed-form states of this repo")
    ap.add_argument("--config", default="uni256", choices=["uni256", "ot2d", "blast_amr"],
                    help="uni256: BASELINE config 4 (headline); ot2d: config 3, 2-D Orszag-Tang + FOFC; "
                         "blast_amr: config 5, adaptive AMR blast (one GPU)")
    ap.add_argument("--recon", default="ppm", choices=["ppm", "wenoz", "plm", "weno3", "limo3"],
                    help="reconstruction of the corrector stage (headline metric: ppm)")
    ap.add_argument("--fp-mode", dest="fp_mode", default="fast", choices=["fast", "strict"])
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-cycles", type=int, default=12)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "blast_amr":
        run_blast_amr(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
