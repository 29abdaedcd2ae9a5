"""Regenerates tests/golden/oracle_linear_wave_table.json: the linear-wave RMS-L1 table as
computed by the CPU oracle (oracle/apk_oracle.c) for the rows of reference_scalars.json.
Run from the repo root:  python tests/golden/make_oracle_table.py  (a few minutes on one core).
The GPU tests require the CUDA path (strict FP mode) to print exactly these digits and to agree
with the reference-run table within 2e-6 relative (14 of 16 entries agree to all 7 digits; see
DESIGN.md 'Oracle pinning' for the two last-digit deviations)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Oracle  # noqa: E402

gold = json.load(open(os.path.join(HERE, "reference_scalars.json")))
rows = []
for row in gold["linear_wave_rms_l1"]:
    out = {k: row[k] for k in ("fluid", "integrator", "recon", "riemann", "ng")}
    for N in (16, 32):
        o = Oracle(nx=(2 * N, N, N), fluid=row["fluid"], integrator=row["integrator"],
                   recon=row["recon"], riemann=row["riemann"], ng=row["ng"])
        o.pgen("linear_wave", [0, 1e-6, 0.0])
        o.initialize(o.lw_tlim)
        o.run(o.lw_tlim, 100000)
        out["N%d" % N] = "%.6e" % o.linear_wave_errors()[0]
        out["N%d_cycles" % N] = o.ncycle
    rows.append(out)
    print(out, flush=True)
json.dump({"_provenance": "oracle/apk_oracle.c via tests/golden/make_oracle_table.py",
           "linear_wave_rms_l1": rows},
          open(os.path.join(HERE, "oracle_linear_wave_table.json"), "w"), indent=1)
