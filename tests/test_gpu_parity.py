"""GPU parity: the CUDA path (through the C ABI and the reference-facing `BossRuns`) against the oracle on
the committed golden cases, batch by batch, plus the reference's own recorded outputs."""
import numpy as np
import pytest

import helpers as H
import tolerances as tol
from golden_io import CASES, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES)
def test_golden_case(case, lib):
    g = load_case(case)
    orc = H.oracle_run(g.records, g.ploidy, g.reject_refs, g.barcodes, g.bucket_threshold)
    prod = H.product_run(g.records, g.ploidy, g.reject_refs, g.barcodes, g.bucket_threshold)
    assert int(prod.ref.n_sites) == int(g.ref("n_sites"))
    for bi, (paf, seqs, bcs) in enumerate(g.batches):
        pd = H.parse_batch(paf, bcs, g.barcodes is not None)
        upd_o = H.oracle_step(orc, pd, seqs)
        upd_p = H.product_step(prod, pd, seqs)
        assert upd_o == upd_p == bool(g.ref(f"b{bi}_updated"))
        H.compare_state(prod, orc, upd_p, f"{case}/b{bi}")
        # and against what the reference itself produced
        for cname, pc in prod.contigs.items():
            want = g.ref(f"b{bi}_{cname}_strat")
            if pc.rej or not upd_p:
                assert np.array_equal(pc.strat, want)
        if upd_p:
            thr = float(g.ref(f"b{bi}_threshold"))
            assert abs(prod.threshold - thr) <= tol.THRESHOLD_RTOL * thr
