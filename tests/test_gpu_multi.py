"""Multi-GPU parity (needs >= 2 GPUs on the box: `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu`): meshblocks split over ranks, ghost zones through the packed
NCCL send/recv path, results bit-identical to the single-process CPU oracle."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


CASES = [
    dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", ng=3, nx=[32, 32, 32],
         mb=[16, 16, 16], xmin=[0, 0, 0], xmax=[1, 1, 1], pgen="ot3d", fused=True),
    dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", ng=3, nx=[32, 32, 32],
         mb=[16, 16, 16], xmin=[0, 0, 0], xmax=[1, 1, 1], pgen="random", pgen_pars=[3, 0.2],
         fused=False),
    dict(fluid="euler", recon="plm", riemann="hllc", integrator="rk3", ng=2, nx=[32, 16, 1],
         mb=[8, 8, 1], xmin=[-0.5, -0.5, -0.5], xmax=[0.5, 0.5, 0.5], gamma=1.4,
         bc=["outflow"] * 4 + ["periodic"] * 2, pgen="blast", pgen_pars=[0.2, 0.1, 20.0, 1.0],
         fused=True),
    # 10 cycles through the fused p2p path with physical (outflow) boundaries on every face: the
    # last-stage barrier may not be elided here (k_bc reads neighbour-pushed ghost cells)
    dict(fluid="glmmhd", recon="plm", riemann="hlle", integrator="vl2", ng=2, nx=[32, 32, 32],
         mb=[16, 16, 16], xmin=[-0.5, -0.5, -0.5], xmax=[0.5, 0.5, 0.5], gamma=1.4,
         bc=["outflow"] * 6, pgen="blast", pgen_pars=[0.2, 0.1, 20.0, 1.0], fused=True,
         cycles=10),
    # rk3 (odd stage count: the prim ping-pong parity persists across cycles), 3-D, periodic
    dict(fluid="euler", recon="wenoz", riemann="hllc", integrator="rk3", ng=3, nx=[32, 32, 32],
         mb=[16, 16, 16], xmin=[0, 0, 0], xmax=[1, 1, 1], gamma=1.4, pgen="random",
         pgen_pars=[5, 0.2], fused=True, cycles=6),
    # world-8 shape of the benchmark partition: 4^3 meshblocks, 8 per rank at 8 GPUs
    dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", ng=3, nx=[64, 64, 64],
         mb=[16, 16, 16], xmin=[0, 0, 0], xmax=[1, 1, 1], pgen="ot3d", fused=True, cycles=2,
         min_world=4),
    # <diffusion> block over several ranks: RKL2 super time stepping (every sub-step exchanges
    # ghost zones) with anisotropic conduction, viscosity and resistivity; the anisotropic
    # conduction time step is a device reduction + all-reduce
    dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", ng=3, nx=[32, 32, 32],
         mb=[16, 16, 16], xmin=[0, 0, 0], xmax=[1, 1, 1], pgen="ot3d", fused=False, cycles=3,
         diffint="rkl2", conduction="anisotropic", thermal_diff_coeff=0.1, viscosity="isotropic",
         mom_diff_coeff=0.04, resistivity="ohmic", ohm_diff_coeff=0.05),
    # BASELINE config 5's method on a frozen 3-level mesh (176 meshblocks of 8^3) split over the
    # ranks: every cross-rank relation of the multilevel exchange and of the flux correction is a
    # box copy that pulls from the owner's IPC-mapped memory, ordered by stage barriers
    dict(fluid="glmmhd", recon="wenoz", riemann="hlld", integrator="vl2", ng=4, nx=[32, 32, 32],
         mb=[8, 8, 8], xmin=[-0.5, -0.5, -0.5], xmax=[0.5, 0.5, 0.5], gamma=1.666666666666667,
         refine=[[1, [-0.2] * 3, [0.2] * 3], [2, [-0.1] * 3, [0.1] * 3]], pgen="blast",
         pgen_pars=[0.12, 0.1, 100.0, 1.0, 0.22, 0.1, -0.05], fused=False, cycles=5, amr=True),
    dict(fluid="euler", recon="plm", riemann="hlle", integrator="rk2", ng=2, nx=[32, 32, 1],
         mb=[8, 8, 1], xmin=[-0.5, -0.5, -0.5], xmax=[0.5, 0.5, 0.5], gamma=1.666666666666667,
         refine=[[1, [-0.2] * 3, [0.2] * 3]], pgen="blast",
         pgen_pars=[0.12, 0.1, 100.0, 1.0, 0.22, 0.1, -0.05], fused=False, cycles=8, amr=True),
    # a passive scalar through the fused stage kernel: the peer stores carry nhydro + 1 variables
    dict(fluid="glmmhd", recon="ppm", riemann="hlld", integrator="vl2", ng=3, nx=[32, 32, 32],
         mb=[16, 16, 16], xmin=[0, 0, 0], xmax=[1, 1, 1], pgen="random", pgen_pars=[3, 0.2],
         fused=True, nscalars=1),
]


@pytest.mark.parametrize("p2p", [1, 0])
@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_bitexact(case, world, p2p):
    """p2p=1: remote ghosts are peer stores of the fused stage kernel (apk_stage_fused_p2p);
    p2p=0: packed ncclSend/ncclRecv exchange.  Both must reproduce the oracle bit for bit."""
    if p2p == 0 and not CASES[case]["fused"]:
        pytest.skip("un-fused decks always use the NCCL path")
    if p2p == 0 and case not in (0, 3):
        pytest.skip("packed NCCL fallback: covered by cases 0 and 3")
    if ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    if world < CASES[case].get("min_world", 2) or (
            world == 8 and "min_world" not in CASES[case] and not CASES[case].get("amr")):
        pytest.skip("case not sized for this world size")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(world):
        kw = dict(CASES[case])
        cycles = kw.pop("cycles", 3)
        kw.pop("min_world", None)
        kw.pop("amr", None)
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), APK_CASE=json.dumps(kw), APK_CYCLES=str(cycles),
                   APK_B200_P2P=str(p2p))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "multigpu_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    assert "MULTIGPU_OK" in outs[0]
