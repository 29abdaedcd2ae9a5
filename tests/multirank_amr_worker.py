"""Worker for test_multirank_cpu.py::test_multilevel_plan_over_gloo: one process per rank, gloo,
CPU only (plan_only mode: no CUDA).  A refined (static AMR) mesh is cut into contiguous ranges of
the leaf list; every rank builds the items of the multilevel ghost exchange and of the flux
correction that IT executes (athenapk_b200/host/hydro_sim.cpp, BuildMultilevelPlan).  Checked:
  * the union over the ranks of every list equals the plan of the single-rank run, item for item
    (nothing lost, nothing done twice);
  * an item is executed by the owner of the block whose memory it WRITES (pull model);
  * the same-level cross-rank pulls are exactly the same-level relations the partition cuts, and
    they are symmetric (b pulls from n  <=>  n pulls from b)."""
import json
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athenapk_b200 as apk  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    kw = json.loads(os.environ["APK_CASE"])
    kw["refine"] = [(lev, tuple(lo), tuple(hi)) for lev, lo, hi in kw["refine"]]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = apk.HydroSim(rank=rank, nranks=world, plan_only=True, fused=False, **kw)
    one = apk.HydroSim(rank=0, nranks=1, plan_only=True, fused=False, **kw)
    mine = sim.ml_items()
    gids = set(sim.gids())
    everything = [None] * world
    dist.all_gather_object(everything, mine)
    owners = [None] * world
    dist.all_gather_object(owners, sorted(gids))
    owner_of = {g: r for r, gl in enumerate(owners) for g in gl}
    assert sorted(owner_of) == list(range(sim.nblocks_global))
    assert sim.nblocks_global == one.nblocks_global

    # pull model: the executing block is the destination block and is local
    for it in mine:
        lst, ex, src, dst = it[:4]
        assert ex in gids and dst == ex, it
    union = sorted(it for r in everything for it in r if it[0] != 5)
    ref = sorted(one.ml_items())
    assert all(it[0] != 5 for it in ref)       # one rank: no remote same-level neighbours
    assert union == ref, (len(union), len(ref))
    # cross-rank relations exist on this partition and cover coarse <-> fine relations
    cross = [it for r in everything for it in r if owner_of[it[2]] != owner_of[it[3]]]
    same = sorted((it[3], it[2]) for r in everything for it in r if it[0] == 5)
    for b, n in same:
        assert owner_of[b] != owner_of[n]
    pairs = set(same)
    assert all((n, b) in pairs for b, n in pairs)
    # every same-level relation cut by the partition is pulled exactly once per offset
    want = 0
    for b in range(one.nblocks):
        for o3 in (-1, 0, 1):
            for o2 in (-1, 0, 1):
                for o1 in (-1, 0, 1):
                    nb = one.neighbor(b, o1, o2, o3)
                    if (o1 or o2 or o3) and nb >= 0 and owner_of[one.gids()[b]] != owner_of[one.gids()[nb]]:
                        want += 1
    assert len(same) == want, (len(same), want)
    if rank == 0:
        kinds = sorted(set(it[0] for it in cross))
        print("MULTIRANK_AMR_OK items=%d cross_rank=%d lists=%s" % (len(union) + len(same), len(cross), kinds))
    sim.close()
    one.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
