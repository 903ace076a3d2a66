"""Pins the oracle: the NumPy restatement (oracle/boss_oracle.py) must reproduce, bit for bit, what the upstream
reference itself produced on the committed inputs (tests/golden/case_*.npz, written by oracle/make_golden.py
running /root/reference's own modules). Runs on CPU; needs neither the reference nor a GPU."""
import hashlib

import numpy as np
import pytest

import helpers as H
from golden_io import CASES, load_case


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference(case):
    g = load_case(case)
    orc = H.oracle_run(g.records, g.ploidy, g.reject_refs, g.barcodes, g.bucket_threshold)
    assert orc.n_sites == int(g.ref("n_sites"))
    first = next(iter(orc.contigs_filt.values()))
    assert first.score0 == float(g.ref("contig_score0"))          # Q5: haploid constant whatever the ploidy
    assert orc.model.score0 == float(g.ref("score0"))
    n_updated = 0
    for bi, (paf, seqs, bcs) in enumerate(g.batches):
        pd = H.parse_batch(paf, bcs, g.barcodes is not None)
        updated = H.oracle_step(orc, pd, seqs)
        p = f"b{bi}_"
        assert np.array_equal(orc.rl.approx_ccl, g.ref(p + "approx_ccl"))
        tc = float(g.ref(p + "time_cost"))
        assert (np.isnan(tc) and not hasattr(orc.rl, "time_cost")) or orc.rl.time_cost == tc
        assert updated == bool(g.ref(p + "updated"))
        n_updated += updated
        if updated:
            assert orc.threshold == float(g.ref(p + "threshold"))
            if g.has(p + "benefit_adj"):
                assert np.array_equal(orc.benefit_adj, g.ref(p + "benefit_adj"))
                assert np.array_equal(orc.fhat_adj, g.ref(p + "fhat_adj"))
            shape = tuple(g.ref(p + "merged_strat_shape"))
            want = np.unpackbits(g.ref(p + "merged_strat"))[: int(np.prod(shape))].reshape(shape).astype(bool)
            assert np.array_equal(orc.merged_strat, want)
        for cname, c in orc.contigs.items():
            q = f"{p}{cname}_"
            assert np.array_equal(c.strat, g.ref(q + "strat")), f"{case} {q}strat"
            if c.rej:
                continue
            assert sha(c.coverage) == str(g.ref(q + "coverage_sha")), f"{case} {q}coverage"
            assert sha(c.scores) == str(g.ref(q + "scores_sha")), f"{case} {q}scores"
            assert np.array_equal(c.bucket_switches, g.ref(q + "bucket_switches"))
            assert np.array_equal(c.switched_on, g.ref(q + "switched_on"))
            assert np.array_equal(c.scores[::97], g.ref(q + "scores_sample"))
            assert np.array_equal(c.coverage[::97], g.ref(q + "coverage_sample"))
            if g.has(q + "coverage"):
                assert np.array_equal(c.coverage, g.ref(q + "coverage"))
                assert np.array_equal(c.scores, g.ref(q + "scores"))
            if updated:
                for name in ("scores_ds", "smu", "expected_benefit", "additional_benefit"):
                    if g.has(q + name):
                        assert np.array_equal(getattr(c, name), g.ref(q + name)), f"{case} {q}{name}"
    assert n_updated >= 2, "a golden case must exercise the strategy branch more than once"


def test_cases_cover_the_quirks():
    """The committed cases must actually trigger what they are there for: dropout zeros, frozen sites, sticky
    buckets gating the distribution, padded merged rows (reject refs) and barcodes."""
    g = load_case("hap_nb1")
    last = len(g.batches) - 1
    sc = g.ref(f"b{last}_trk1_scores")
    assert (sc == 0.0).any(), "dropout rule never fired"
    assert (sc == np.finfo(float).tiny).any(), "no site froze at depth 30"
    sw = g.ref(f"b{last}_trk1_bucket_switches")
    assert sw.any() and not sw.all(), "bucket switches should gate part of the contig"
    pad = load_case("hap_pad")
    rows = int(pad.ref("n_sites")) // 100
    merged = tuple(pad.ref("b2_merged_strat_shape"))
    assert merged[0] == rows and rows > 130_050 // 100 + 1, "hap_pad must exercise adjust_length padding"
    assert load_case("hap_nb3").barcodes == ["barcode01", "barcode02", "barcode03"]
