"""Known-answer tests of the CPU oracle against numbers produced by the unmodified reference
(tests/golden/reference_scalars.json): the linear-wave convergence table (7 printed digits,
incl. the reference's own regression threshold 1.547584e-08), the Sod dt trace end point and
the Orszag-Tang (time, dt) pair after 50 cycles at all 17 printed digits.  CPU only."""
import json
import os

import pytest

from oracle_lib import Oracle

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_scalars.json")))


def run_linear_wave(N, row):
    o = Oracle(nx=(2 * N, N, N), fluid=row["fluid"], integrator=row["integrator"],
               recon=row["recon"], riemann=row["riemann"], ng=row["ng"])
    o.pgen("linear_wave", [0, 1e-6, 0.0])
    o.initialize(o.lw_tlim)
    o.run(o.lw_tlim, 100000)
    return o


@pytest.mark.parametrize("row", GOLD["linear_wave_rms_l1"],
                         ids=lambda r: "-".join([r["fluid"], r["integrator"], r["recon"], r["riemann"]]))
def test_linear_wave_table_N16(row):
    o = run_linear_wave(16, row)
    assert o.ncycle == GOLD["linear_wave_cycles"]["N16"]
    assert "%.6e" % o.linear_wave_errors()[0] == row["N16"]


@pytest.mark.parametrize("idx", [1, 6])
def test_linear_wave_table_N32(idx):
    row = GOLD["linear_wave_rms_l1"][idx]
    o = run_linear_wave(32, row)
    assert o.ncycle == GOLD["linear_wave_cycles"]["N32"]
    assert "%.6e" % o.linear_wave_errors()[0] == row["N32"]


@pytest.mark.slow
def test_reference_regression_threshold_N64():
    """convergence.py:163: VL2+PLM+HLLE at 128x64x64 must give RMS-L1 <= 1.547584e-08."""
    row = GOLD["linear_wave_rms_l1"][1]
    o = run_linear_wave(64, row)
    assert o.ncycle == GOLD["linear_wave_cycles"]["N64"]
    assert "%.6e" % o.linear_wave_errors()[0] == GOLD["convergence_py_163"]["rms_l1"]


def test_sod_dt_trace():
    o = Oracle(nx=(256, 1, 1), mb=(64, 1, 1), ng=2,
               bc=["outflow", "outflow", "periodic", "periodic", "periodic", "periodic"],
               xmin=(0.0, -0.5, -0.5), xmax=(1.0, 0.5, 0.5), gamma=1.4, cfl=0.5, recon="plm",
               riemann="hllc", integrator="vl2")
    o.pgen("sod")
    o.initialize(0.4)
    o.run(0.4, 10000)
    assert o.ncycle == GOLD["sod_hllc"]["cycles"]
    assert "%.16e" % o.dt == GOLD["sod_hllc"]["final_dt"]


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
