"""Worker for test_multirank_cpu.py: one process per rank, gloo backend, CPU only.

Exercises the host-side N>1 logic of athenapk_b200/host/hydro_sim.cpp without a GPU: Morton
partition of the block list, neighbour table, and the per-peer packed halo plan (both sides
must derive identical slab order/offsets independently).  The pack/unpack kernels are emulated
in numpy with the layout of k_halo_copy ([field][var][k][j][i] per slab at item.offset)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athenapk_b200 as apk  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    nx = tuple(int(x) for x in os.environ["APK_NX"].split(","))
    mb = tuple(int(x) for x in os.environ["APK_MB"].split(","))
    ng = int(os.environ.get("APK_NG", "2"))
    bc = os.environ.get("APK_BC", "periodic")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sim = apk.HydroSim(fluid="glmmhd", recon="plm", riemann="hlld", nx=nx, mb=mb, ng=ng, bc=bc,
                       rank=rank, nranks=world, plan_only=True)
    nvar, (n1, n2, n3) = sim.nvar, sim.n
    ndim = 3 if nx[2] > 1 else (2 if nx[1] > 1 else 1)
    s = [ng, ng if ndim >= 2 else 0, ng if ndim >= 3 else 0]

    def analytic(n, gi, gj, gk):
        return n * 1.0e6 + (gk % nx[2]) * 1.0e4 + (gj % nx[1]) * 1.0e2 + (gi % nx[0])

    # all ranks agree on the global block count and own disjoint gid ranges
    gids = sim.gids()
    all_gids = [None] * world
    dist.all_gather_object(all_gids, gids)
    flat = sorted(g for gl in all_gids for g in gl)
    assert flat == list(range(sim.nblocks_global)), flat

    blocks = []
    for b in range(sim.nblocks):
        loc = sim.block_info(b)[1]
        a = np.full((nvar, n3, n2, n1), np.nan)
        for n in range(nvar):
            for k in range(mb[2]):
                for j in range(mb[1]):
                    gi = loc[0] * mb[0] + np.arange(mb[0])
                    a[n, s[2] + k, s[1] + j, s[0]:s[0] + mb[0]] = analytic(
                        n, gi, loc[1] * mb[1] + j, loc[2] * mb[2] + k)
        blocks.append(a)

    def slab(it):
        lo, ext = it["lo"], it["ext"]
        return (slice(None), slice(lo[2], lo[2] + ext[2]), slice(lo[1], lo[1] + ext[1]),
                slice(lo[0], lo[0] + ext[0]))

    # local neighbours: direct copy (what k_fill_ghosts does)
    for b in range(sim.nblocks):
        for o3 in (-1, 0, 1):
            for o2 in (-1, 0, 1):
                for o1 in (-1, 0, 1):
                    nb = sim.neighbor(b, o1, o2, o3)
                    if nb < 0 or (o1, o2, o3) == (0, 0, 0):
                        continue
                    off = (o1, o2, o3)
                    dst, src = [], []
                    for d in range(3):
                        e = s[d] + mb[d] - 1
                        if off[d] == 0:
                            dst.append(slice(s[d], e + 1)); src.append(slice(s[d], e + 1))
                        elif off[d] > 0:
                            dst.append(slice(e + 1, e + 1 + ng)); src.append(slice(s[d], s[d] + ng))
                        else:
                            dst.append(slice(s[d] - ng, s[d])); src.append(slice(e - ng + 1, e + 1))
                    blocks[b][:, dst[2], dst[1], dst[0]] = blocks[nb][:, src[2], src[1], src[0]]

    # ---- peer-push table (what apk_stage_fused_p2p is fed with): every (block, offset) names the
    # rank and local index of the receiving block; check it against the gathered block lists and
    # that the relation is symmetric (the receiver's opposite offset points back at the sender)
    mine = [(sim.block_info(b)[0], tuple(sim.block_info(b)[1])) for b in range(sim.nblocks)]
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    targets = {}
    nblk = [nx[d] // mb[d] for d in range(3)]
    for b in range(sim.nblocks):
        gid, loc = mine[b]
        for o3 in (-1, 0, 1):
            for o2 in (-1, 0, 1):
                for o1 in (-1, 0, 1):
                    off = (o1, o2, o3)
                    t = sim.push_target(b, o1, o2, o3)
                    has_nbr = off != (0, 0, 0)
                    for d in range(3):
                        l = loc[d] + off[d]
                        if off[d] != 0 and (nx[d] == 1 or ((l < 0 or l >= nblk[d]) and bc != "periodic")):
                            has_nbr = False
                    assert (t is not None) == has_nbr, (rank, b, off, t)
                    if t is None:
                        continue
                    want_loc = tuple((loc[d] + off[d]) % nblk[d] for d in range(3))
                    tgid, tloc = everyone[t[0]][t[1]]
                    assert tloc == want_loc, (rank, b, off, t, tloc, want_loc)
                    assert (t[0] == rank) == (sim.neighbor(b, o1, o2, o3) >= 0)
                    targets[(gid, off)] = tgid
    all_targets = [None] * world
    dist.all_gather_object(all_targets, targets)
    merged = {}
    for t in all_targets:
        merged.update(t)
    for (gid, off), tgid in targets.items():
        back = merged[(tgid, tuple(-o for o in off))]
        assert back == gid, (gid, off, tgid, back)

    # remote neighbours: pack -> gloo send/recv per peer -> unpack
    peers = sim.peers()
    send_items, recv_items = sim.halo_items(False), sim.halo_items(True)
    send_total = max([p["send_off"] + p["send_cnt"] for p in peers] + [1])
    recv_total = max([p["recv_off"] + p["recv_cnt"] for p in peers] + [1])
    sendbuf = np.zeros(send_total)
    recvbuf = np.full(recv_total, np.nan)
    for it in send_items:
        data = blocks[it["block"]][slab(it)].reshape(-1)
        sendbuf[it["offset"]:it["offset"] + data.size] = data
    reqs = []
    ts, tr = torch.from_numpy(sendbuf), torch.from_numpy(recvbuf)
    for p in peers:
        reqs.append(dist.isend(ts[p["send_off"]:p["send_off"] + p["send_cnt"]], p["rank"]))
        reqs.append(dist.irecv(tr[p["recv_off"]:p["recv_off"] + p["recv_cnt"]], p["rank"]))
    for r in reqs:
        r.wait()
    for it in recv_items:
        sl = slab(it)
        shape = blocks[it["block"]][sl].shape
        n = int(np.prod(shape))
        blocks[it["block"]][sl] = recvbuf[it["offset"]:it["offset"] + n].reshape(shape)

    # every ghost cell that has a (periodic / interior) neighbour must now hold the analytic value
    nb_ = [nx[d] // mb[d] for d in range(3)]
    checked = 0
    for b in range(sim.nblocks):
        loc = sim.block_info(b)[1]
        for k in range(n3):
            for j in range(n2):
                for i in range(n1):
                    idx, off, ok = (i, j, k), [0, 0, 0], True
                    for d in range(3):
                        if nx[d] == 1:
                            continue
                        if idx[d] < s[d]:
                            off[d] = -1
                        elif idx[d] > s[d] + mb[d] - 1:
                            off[d] = 1
                        l = loc[d] + off[d]
                        if (l < 0 or l >= nb_[d]) and bc != "periodic":
                            ok = False
                    if off == [0, 0, 0] or not ok:
                        continue
                    g = [loc[d] * mb[d] + (idx[d] - s[d]) for d in range(3)]
                    want = [analytic(n, g[0], g[1], g[2]) for n in range(nvar)]
                    got = blocks[b][:, k, j, i]
                    assert np.array_equal(got, want), (rank, b, idx, got[:2], want[:2])
                    checked += 1
    assert checked > 0
    tot = torch.tensor([checked], dtype=torch.int64)
    dist.all_reduce(tot)
    if rank == 0:
        print("MULTIRANK_OK ghosts_checked=%d peers=%d" % (int(tot), len(peers)))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
