import sys, json, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import athenapk_b200 as apk
from parity_util import make_pair, rel_l1, rel_l1_per_variable
GOLD = json.load(open("tests/golden/reference_scalars.json"))
for i in (1, 6):
    row = GOLD["linear_wave_rms_l1"][i]
    s, o = make_pair(apk, "linear_wave", (0, 1e-6, 0.0), tlim=None, nx=(64, 32, 32), mb=(32, 16, 16), fluid=row["fluid"], integrator=row["integrator"], recon=row["recon"], riemann=row["riemann"], ng=row["ng"], fp_mode="fast")
    o.run(o.lw_tlim, 100000); s.execute(s.lw_tlim, 100000)
    print(row["fluid"], row["recon"], row["riemann"], row["integrator"], s.ncycle, o.ncycle, rel_l1(s, o, "cons"), rel_l1_per_variable(s, o, "cons"), s.linear_wave_errors()[0] / o.linear_wave_errors()[0] - 1)
    s.close()
