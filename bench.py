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


def cpu_sample(cycles, n=64, mb=32):
    """time the CPU oracle (port of the reference algorithm, single thread) on a bounded sample"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    o = Oracle(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", nx=(n, n, n),
               mb=(mb,) * 3, ng=3, xmin=(0.0, 0.0, 0.0), xmax=(1.0, 1.0, 1.0), gamma=5.0 / 3.0,
               cfl=0.3)
    o.pgen("ot3d")
    o.initialize()
    o.run(None, 1)  # warm-up cycle
    t0 = time.time()
    o.run(None, 1 + cycles)
    dt = time.time() - t0
    return n ** 3 * cycles / dt


def _cpu_worker(q, cycles):
    q.put(cpu_sample(cycles))


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle port of the
    Kokkos-Serial code; the reference itself needs cmake+Kokkos to build), one
    independent 64^3 (8 meshblocks of 32^3) replica per host core, aggregate zone-cycles/s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = min(os.cpu_count() or 1, 64)
    cycles = 2
    vals = []
    for step in range(args.warmup + args.steps):
        q = mp.Queue()
        ps = [mp.Process(target=_cpu_worker, args=(q, cycles)) for _ in range(cores)]
        t0 = time.time()
        for p in ps:
            p.start()
        for p in ps:
            p.join()
        wall = time.time() - t0
        each = [q.get() for _ in ps]
        agg = sum(each)
        if step >= args.warmup:
            vals.append((agg, wall))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(w for _, w in vals) / len(vals)
    sample = ("%d independent replicas (one per core) of 64^3 GLM-MHD VL2+PPM+HLLD, 8 meshblocks "
              "of 32^3, %d timed cycles each; aggregate zone-cycles/s" % (cores, cycles))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "uniform 256^3 GLM-MHD ot3d, VL2+PPM+HLLD, meshblock 32^3 "
                                   "(CPU arm: bounded 64^3 samples, per-cell cost is size independent)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


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
    cells = float(args.n) ** 3
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
        roof = {"bound": "hbm", "kernel": ("k_stage_fused<glmmhd,%s,hlld,3d> (stage 2)" % args.recon) if dom ==
                "fused_stage_other" else dom, "achieved": achieved, "peak": peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "alg_bytes_per_launch": bytes_launch, "launch_ms": dur_ms}
        tf = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tf):
            try:
                roof["traffic"] = json.load(open(tf)).get("dram_bytes_per_launch_256")
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
        try:
            import csv
            rows = {r[0]: r[1:] for r in csv.reader(open(os.path.join(ROOT, "profiles",
                                                                      "ncu_r01_final_256.csv")))}
            roof["ncu_fp64_pipe_active_pct"] = float(
                rows["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][-1])
            roof["ncu_issue_active_pct"] = float(
                rows["smsp__issue_active.avg.pct_of_peak_sustained_active"][-1])
            roof["note"] = ("HBM is the contract roofline; the kernel is limited by the half-rate "
                            "FP64 pipe and FP64 dependency latency (ncu figures from profiles/)")
        except Exception:
            pass

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
               "serial_ms_per_step": serial_ms,
               "what": "every step: pinned-host conserved state (interior cells, 72 B/cell) -> H2D -> "
                       "scatter + ghost exchange + FillDerived -> 1 full cycle -> gather -> D2H; "
                       "value = three contexts round robin (copies of neighbouring steps overlap compute), "
                       "serial_value = one context, copy-compute-copy back to back; host wall clock; "
                       "PCIe-bound (1.2 GB each way per step at 256^3)"}
        for sx in sims[1:]:
            sx.close()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            t0 = time.time()
            v = cpu_sample(args.cpu_cycles)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "oracle (C port of the reference's Kokkos-Serial path), 64^3 GLM-MHD "
                             "VL2+PPM+HLLD ot3d, 8 meshblocks of 32^3, %d timed cycles, 1 thread, "
                             "%.1f s" % (args.cpu_cycles, time.time() - t0)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "uniform %d^3 GLM-MHD (%s), VL2+%s+HLLD, meshblock %d^3, "
                                       "%d meshblocks, nghost 3, cfl 0.3" %
                                       (args.n, args.pgen, args.recon.upper(), args.mb,
                                        sim.nblocks_global),
                           "fp_mode": args.fp_mode, "fused": sim.fused_active,
                           "l2": "state (%.1f GB/GPU) >> 126 MB L2; no flush needed" %
                                 (sim.nblocks * sim.nvar * sim.n[0] * sim.n[1] * sim.n[2] * 8 * 4 / 1e9),
                           "partition": "Morton-ordered meshblock list split in %d contiguous ranges" % world},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": int(launches), "clocks": clocks, "task_ms": task_ms,
                "host_affinity": {"first_cpu_and_count": numa},
                "p2p_halo": bool(sim.p2p_active),
                "alg_bytes_per_cell_update": ALG_BYTES_PER_UPDATE,
                "hbm_frac_whole_cycle": value * ALG_BYTES_PER_UPDATE / (peak * 1e9 * world)}
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--n", type=int, default=256, help="mesh cells per direction")
    ap.add_argument("--mb", type=int, default=32, help="meshblock cells per direction")
    ap.add_argument("--pgen", default="ot3d", choices=["ot3d", "uniform", "random", "orszag_tang"])
    ap.add_argument("--config", default="uni256", choices=["uni256", "ot2d"],
                    help="uni256: BASELINE config 4 (headline); ot2d: config 3, 2-D Orszag-Tang + FOFC")
    ap.add_argument("--recon", default="ppm", choices=["ppm", "wenoz", "plm", "weno3", "limo3"],
                    help="reconstruction of the corrector stage (headline metric: ppm)")
    ap.add_argument("--fp-mode", dest="fp_mode", default="fast", choices=["fast", "strict"])
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-cycles", type=int, default=12)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
