"""Worker for test_gpu_multi.py: one rank per GPU.  Runs a deck through HydroSim with the block
list split over the ranks (packed NCCL halo exchange), then compares every local block with the
single-process CPU oracle, bit for bit (strict FP mode)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import athenapk_b200 as apk  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from parity_util import oracle_block  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = json.loads(os.environ["APK_CASE"])
    if "refine" in kw:
        kw["refine"] = [(lev, tuple(lo), tuple(hi)) for lev, lo, hi in kw["refine"]]
    ncyc = int(os.environ.get("APK_CYCLES", "3"))
    pgen, pars = kw.pop("pgen"), kw.pop("pgen_pars", [])
    fused = kw.pop("fused", True)
    sim = apk.HydroSim(fp_mode="strict", fused=fused, rank=rank, nranks=world, **kw)
    uid = [apk.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim.init_comm(uid[0])
    want_p2p = os.environ.get("APK_B200_P2P", "1") != "0"
    # fused decks exchange remote ghosts as peer stores inside the stage kernel (CUDA IPC);
    # APK_B200_P2P=0 keeps the packed ncclSend/ncclRecv path
    assert sim.p2p_active == (want_p2p and sim.fused_active), (sim.p2p_active, want_p2p)
    sim.pgen(pgen, pars)
    sim.initialize()
    sim.execute(None, ncyc)
    o = Oracle(**kw)
    o.pgen(pgen, pars)
    o.initialize()
    o.run(None, ncyc)
    assert (sim.ncycle, sim.time, sim.dt) == (o.ncycle, o.time, o.dt), (sim.time, o.time, sim.dt, o.dt)
    for b in range(sim.nblocks):
        ob = oracle_block(sim, o, b)  # (refined meshes: matched by level and logical location)
        for f, ref in (("cons", o.cons(ob)), ("prim", o.prim(ob))):
            got = sim.get(f, b)
            assert np.array_equal(got, ref), (rank, b, f, float(np.max(np.abs(got - ref))))
    dist.barrier()
    if rank == 0:
        print("MULTIGPU_OK blocks/rank=%d p2p=%d" % (sim.nblocks, sim.p2p_active))
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
