"""CPU: the host driver's tree update (athenapk_b200/host/hydro_amr.cpp, UpdateMeshBlockTree -- pure
host code behind apk_sim_amr_update_tree) against the oracle's, on the remesh inputs of the
reference-pinned adaptive runs (tests/golden/reference_traces_amr.json): same new leaf list in the
same (Parthenon block-list) order, same created / destroyed counts, every cycle."""
import pytest

from oracle_lib import Oracle
from trace_util import ADAPTIVE_DECKS

pytest.importorskip("torch")


@pytest.mark.parametrize("name,ncycle", [("adaptive_3d_plm_hlle", 14), ("adaptive_2d_plm_hllc", 120)])
def test_host_tree_update_matches_oracle(name, ncycle):
    import athenapk_b200 as apk
    kw, pgen, pars, tlim = ADAPTIVE_DECKS[name]
    cfg = apk.make_config(**kw)
    o = Oracle(**kw)
    o.pgen(pgen, pars)
    o.initialize(tlim)
    changed = refined = derefined = 0
    for c in range(ncycle):
        before = o.amr_counts()
        o.run(tlim, c + 1)
        leaves, flags = o.amr_last_input()
        new, nnew, ndel = apk.amr_update_tree(cfg, leaves, flags)
        after = o.amr_counts()
        assert (nnew, ndel) == (after[0] - before[0], after[1] - before[1]), c
        now = [(o.block_loc(b)[0],) + o.block_loc(b)[1] for b in range(o.nblocks)]
        if new is None:
            assert sorted(now) == sorted(leaves)
        else:
            assert new == now, c  # same leaves in the same order
            changed += 1
        refined += nnew > 0
        derefined += ndel > 0
    assert changed > 5 and refined > 0 and derefined > 0
