"""Known-answer tests of the CPU oracle against the reference's own regression threshold
(tst/regression/test_suites/convergence/convergence.py:163) and the Orszag-Tang history row.  The
per-cycle (time, dt) traces and linear-wave error rows printed by the unmodified reference are in
test_oracle_traces.py.  CPU only."""
import json
import os

import pytest

from oracle_lib import Oracle
from trace_util import GAMMA_DECK

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_scalars.json")))


def run_linear_wave(N, row):
    o = Oracle(nx=(2 * N, N, N), fluid=row["fluid"], integrator=row["integrator"],
               recon=row["recon"], riemann=row["riemann"], ng=row["ng"], gamma=GAMMA_DECK,
               xmin=(0.0, 0.0, 0.0), xmax=(3.0, 1.5, 1.5))
    o.pgen("linear_wave", [0, 1e-6, 0.0])
    o.initialize(o.lw_tlim)
    o.run(o.lw_tlim, 100000)
    return o


@pytest.mark.slow
def test_reference_regression_threshold_N64():
    """convergence.py:163: VL2+PLM+HLLE at 128x64x64 must give RMS-L1 <= 1.547584e-08."""
    row = GOLD["linear_wave_rms_l1"][1]
    o = run_linear_wave(64, row)
    assert o.ncycle == GOLD["linear_wave_cycles"]["N64"]
    assert "%.6e" % o.linear_wave_errors()[0] == GOLD["convergence_py_163"]["rms_l1"]


def test_orszag_tang_50_cycles():
    o = Oracle(fluid="glmmhd", nx=(256, 256, 1), mb=(64, 256, 1), ng=3, xmin=(-0.5,) * 3,
               xmax=(0.5,) * 3, gamma=1.666666666666667, cfl=0.4, recon="ppm", riemann="hlld",
               integrator="vl2", fofc=True)
    o.pgen("orszag_tang")
    o.initialize(1.0)
    o.run(1.0, 50)
    g = GOLD["orszag_tang_50"]
    assert "%.16e" % o.time == g["time"]
    assert "%.16e" % o.dt == g["dt"]
    h = o.history()
    assert "%.5e" % h[4] == g["hst"]["KE"]
    assert "%.5e" % h[6] == g["hst"]["ME"]
    assert "%.5e" % h[7] == g["hst"]["relDivB"]
    assert "%.5e" % h[5] == g["hst"]["tot-E"]
