"""world_size-2 (and 4) gloo tests of the multi-GPU host logic on CPU: Morton block partition,
neighbour tables and the packed per-peer halo plan of athenapk_b200/host/hydro_sim.cpp (the
device pack/unpack kernels are emulated in numpy; see multirank_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nx,mb,ng,bc", [
    (2, "16,16,16", "8,8,8", 2, "periodic"),
    (2, "24,8,8", "4,4,4", 3, "periodic"),
    (4, "16,16,8", "4,8,4", 2, "periodic"),
    (2, "32,16,1", "8,8,1", 2, "periodic"),
    (2, "16,16,16", "8,8,8", 2, "outflow"),
    (3, "48,1,1", "8,1,1", 2, "periodic"),
])
def test_halo_plan_over_gloo(world, nx, mb, ng, bc):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), APK_NX=nx, APK_MB=mb, APK_NG=str(ng), APK_BC=bc)
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "multirank_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    assert "MULTIRANK_OK" in outs[0]


AMR3 = dict(fluid="glmmhd", recon="wenoz", riemann="hlld", integrator="vl2", ng=4, nx=[32, 32, 32],
            mb=[8, 8, 8], xmin=[-0.5] * 3, xmax=[0.5] * 3,
            refine=[[1, [-0.2] * 3, [0.2] * 3], [2, [-0.1] * 3, [0.1] * 3]])
AMR2 = dict(fluid="euler", recon="plm", riemann="hlle", integrator="rk2", ng=2, nx=[32, 32, 1],
            mb=[8, 8, 1], xmin=[-0.5] * 3, xmax=[0.5] * 3, refine=[[1, [-0.2] * 3, [0.2] * 3]])


@pytest.mark.parametrize("world,case", [(2, AMR3), (3, AMR3), (4, AMR2), (8, AMR3)])
def test_multilevel_plan_over_gloo(world, case):
    """Refined (static AMR) mesh over several ranks: the per-rank item lists of the multilevel ghost
    exchange and of the flux correction (pull model over IPC-mapped peer memory on the GPUs) add up
    to the single-rank plan, item for item; see multirank_amr_worker.py."""
    import json
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), APK_CASE=json.dumps(case))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "multirank_amr_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    assert "MULTIRANK_AMR_OK" in outs[0]
