"""GPU parity on the shapes BASELINE.json's configs name, beyond the small golden cases:

* C2 at FULL size (4.6 Mb haploid, 4000 x ~10 kb reads, three consecutive updates) against the oracle, state by state;
* C4-shaped (24 barcodes, unclassified reads falling to index 0 — Q11; the whole-row rules Q6/Q8 across 24 planes);
* C5-shaped (hundreds of contigs 10 kb - 600 kb: those under 100 kb dropped at load, 10 % reject refs with their 4-site
  placeholders, the Q2 row shift growing to hundreds of bins) against the oracle;
* C3 at FULL size (3.1 Gb diploid, 25 contigs, 35 GB of state): size-independent properties — a contig embedded
  behind 3.1e9 other sites agrees with the oracle run on that contig alone (per-contig stages are independent upstream,
  core.py:83-121), the exponent histogram accounts for every non-zero benefit, an update over an empty batch is
  idempotent, and four virtual shards give bit-identical thresholds and masks to the one-handle run.
"""
import numpy as np
import pytest

import helpers as H
from oracle import boss_oracle as bo
from boss_runs_b200 import synth

pytestmark = pytest.mark.gpu


def _run_both(records, ploidy, reject, barcodes, bucket_threshold, batches, tag, tracked=None):
    orc = H.oracle_run(records, ploidy, reject, barcodes, bucket_threshold)
    prod = H.product_run(records, ploidy, reject, barcodes, bucket_threshold)
    assert int(prod.ref.n_sites) == int(orc.n_sites)
    assert list(prod.contigs_filt.keys()) == list(orc.contigs_filt.keys())
    n_upd = 0
    for bi, rb in enumerate(batches):
        pd = H.parse_batch(rb.paf_text, rb.barcodes, barcodes is not None)
        uo = H.oracle_step(orc, pd, rb.seqs)
        up = H.product_step(prod, pd, rb.seqs)
        assert uo == up
        H.compare_state(prod, orc, up, f"{tag}/b{bi}")
        n_upd += int(up)
    assert n_upd > 0, "the case never switched a bucket on: nothing but the counters was compared"
    return prod, orc


def test_c2_full_size(lib):
    """BASELINE config 2 as specified: one 4.6 Mb contig, seed 7, batches of 4000 reads of ~10 kb (8.7x per batch):
    dropout, bucket switches and the freeze at depth 30 all trigger by the third batch."""
    contigs = synth.random_contigs({"ecoli": 4_600_000}, seed=7)
    batches = [synth.read_batch(contigs, n_reads=4000, seed=70 + b) for b in range(3)]
    prod, orc = _run_both(list(contigs.items()), 1, [], None, 5, batches, "c2")
    c = orc.contigs["ecoli"]
    depth = c.coverage.sum(axis=1)[:, 0]
    assert (depth >= 30).any() and (c.scores == 0.0).any() and c.bucket_switches.all()


def test_c4_shape_24_barcodes(lib):
    """BASELINE config 4 at 1/5 length: 24 barcodes x 1 Mb in one launch, ~5 % of the reads unclassified (index 0)."""
    contigs = synth.random_contigs({"amplicon_ref": 1_000_000}, seed=17)
    barcodes = [f"barcode{i + 1:02d}" for i in range(24)]
    batches = []
    for b in range(2):
        rb = synth.read_batch(contigs, n_reads=3000, seed=400 + b, mean_len=4000.0, min_len=500, max_len=20_000, n_barcodes=24,
                              focus=("amplicon_ref", 200_000, 260_000, 0.25))
        # barcode INDICES, as the sampler hands them over (sampler.py:218-221): unclassified reads already sit at 0
        batches.append(rb)
    prod, orc = _run_both(list(contigs.items()), 1, [], barcodes, 0.5, batches, "c4")
    cov = orc.contigs["amplicon_ref"].coverage
    assert all(cov[:, :, b].any() for b in range(24))


def test_c5_shape_many_contigs(lib):
    """BASELINE config 5 in miniature: 400 contigs log-uniform 10 kb - 600 kb, seed 13; under 100 kb dropped by the
    loader (reference.py:319,330), 10 % of the names in reject_refs (-> `(1,)` masks and 4 phantom sites each),
    contig k's mask read k bins early (Q2)."""
    rng = np.random.default_rng(13)
    lens = np.exp(rng.uniform(np.log(10_000), np.log(600_000), size=400)).astype(np.int64)
    contigs = synth.random_contigs({f"bin{i:04d}": int(n) for i, n in enumerate(lens)}, seed=13)
    names = list(contigs)
    reject = [names[i] for i in rng.choice(len(names), size=40, replace=False)]
    tracked = {n: s for n, s in contigs.items() if len(s) >= 100_000 and n not in reject}
    assert 100 < len(tracked) < 250
    batches = [synth.read_batch(tracked, n_reads=3000, seed=500 + b, mean_len=5000.0, min_len=500, max_len=30_000) for b in range(2)]
    prod, orc = _run_both(list(contigs.items()), 1, reject, None, 0, batches, "c5")
    assert sum(1 for c in prod.contigs.values() if c.rej) == sum(1 for n in reject if len(contigs[n]) >= 100_000)
    assert not any(len(contigs[n]) < 100_000 for n in prod.contigs)


# ---------------------------------------------------------------------------------------------------------
# C3 at full size
# ---------------------------------------------------------------------------------------------------------
C3_SYNTH = dict(seed=11, mean_depth=8.0, p_ref=0.90, p_del=0.04, frac_dropout=0.02, frac_deep=0.01)


def _c3_inputs():
    lens = synth.grch38_like_lengths(3_100_000_000, 25)
    names = [f"ctg{i + 1}" for i in range(25)]
    rng = np.random.default_rng(7)
    codes = {n: rng.integers(0, 4, size=int(L), dtype=np.uint8) for n, L in zip(names, lens)}
    return names, lens, codes


def _gate(c):
    rows = c.length // 100
    return bo.adjust_length(rows, np.repeat(np.asarray(c.bucket_switches), 200, axis=0))


@pytest.fixture(scope="module")
def c3(lib):
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~80 GB of free HBM (two 3.1 Gb diploid states)")
    from boss_runs_b200.runs import BossRuns
    names, lens, codes = _c3_inputs()
    run = BossRuns(contigs=codes, ploidy=2, bucket_threshold=5, strict_upstream_asserts=False, write_debug=True)
    run.engine.synth_coverage(**C3_SYNTH)
    small = names[-1]                                    # 500 kb contig at padded offset ~3.1e9 (> 2^31 sites in)
    rbs = [synth.read_batch(codes, n_reads=4000, seed=1000 + b, focus=(small, 100_000, 300_000, 0.02)) for b in range(2)]
    return dict(run=run, names=names, lens=lens, codes=codes, small=small, batches=rbs)


def test_c3_full_size_embedded_contig_matches_oracle(c3):
    """The last contig of the 3.1 Gb run, 3.1e9 sites into the arrays, against the oracle run on that contig alone with
    the same counters and the same reads; then the global pieces the oracle cannot afford are checked by identity."""
    run, small, codes = c3["run"], c3["small"], c3["codes"]
    pc = run.contigs[small]
    hap = bo.ScoreModel(1)
    oc = bo.ContigState(small, codes[small], nb=1, score0=hap.score0, ent0=hap.ent0)
    oc.coverage[...] = pc.coverage                       # the synthetic pre-loaded state, read back from the device
    observed = oc.coverage.sum(axis=1)[:, 0] > 0
    model = bo.ScoreModel(2)
    model.build_table()
    k = len(c3["names"]) - 1
    for bi, rb in enumerate(c3["batches"]):
        pd = H.parse_batch(rb.paf_text, {}, False)
        inc = bo.convert_records(pd, rb.seqs).get(small, [])
        assert len(inc) > 10
        bo.increment_coverage(oc, inc)
        oc.change_mask[:, 0] |= observed if bi == 0 else False     # first update scores every observed site
        H.product_step(run, pd, rb.seqs)
        assert run.last.switched_on
        bo.update_scores(oc, model)
        n_drop = bo.modify_scores(oc)
        bo.check_buckets(oc, run.bucket_threshold)
        bo.calc_smu(oc)
        bo.calc_u(oc, run.rl_dist.approx_ccl)
        t = f"c3/b{bi}/{small}"
        assert np.array_equal(pc.coverage, oc.coverage), f"{t}: coverage"
        s = pc.scores
        np.testing.assert_allclose(s, oc.scores, rtol=H.tol.SCORE_RTOL, atol=0, err_msg=f"{t}: scores")
        assert np.array_equal(s == 0.0, oc.scores == 0.0) and n_drop > 0
        assert np.array_equal(pc.bucket_switches, oc.bucket_switches)
        np.testing.assert_allclose(pc.scores_ds, oc.scores_ds, rtol=H.tol.SCORE_RTOL, atol=0, err_msg=f"{t}: scores_ds")
        H.assert_close_smooth(pc.smu, oc.smu, f"{t}: smu")
        H.assert_close_smooth(pc.expected_benefit, oc.expected_benefit, f"{t}: expected_benefit")
        H.assert_close_smooth(pc.additional_benefit, oc.additional_benefit, f"{t}: additional_benefit")
        # Q2: contig k's strategy row j is merged row (own start - k + j): rows k.. are its own benefit rows 0..
        thr = run.threshold
        gate = _gate(pc)[k:, 0]
        n = pc.strat.shape[0]
        want = (oc.additional_benefit >= thr)[: n - k]
        got = np.asarray(pc.strat)[k:]
        diff = (got != want)[gate]
        near = (np.abs(oc.additional_benefit[: n - k] - thr) <= H.tol.MASK_REL * thr)[gate]
        assert not (diff & ~near).any(), f"{t}: {(diff & ~near).sum()} mask bits differ away from the threshold"
        # global identities at 3.1 Gb
        counts, f_grid = run.engine.hist()
        assert int(counts.sum()) == run.last.n_nonzero > 0
        m, _ = np.frexp(thr / run.last.normaliser)
        assert m == 0.5, "the threshold is a power of two times the largest benefit (sequences.py:636-648)"
    # an update over an empty batch changes nothing (counters, read starts and read lengths are all unchanged)
    before = (run.threshold, run.last.ubar0, run.last.n_nonzero, run.last.n_dropout, run.last.n_accept)
    masks = [np.array(c.strat) for c in run.contigs_filt.values()]
    run.process_batch_runs({}, {})
    assert before == (run.threshold, run.last.ubar0, run.last.n_nonzero, run.last.n_dropout, run.last.n_accept)
    for m0, c in zip(masks, run.contigs_filt.values()):
        assert np.array_equal(m0, c.strat)


@pytest.mark.parametrize("exchange", ["phases", "fabric"])
def test_c3_full_size_sharded_is_bit_identical(c3, exchange):
    """Four virtual shards (cuts inside contigs, bin halos, integer-limb histogram) against the one-handle run at 3.1 Gb:
    thresholds, ubar0, counters and every mask bit equal."""
    from boss_runs_b200.sharding import ShardedRun
    one = c3["run"]
    many = ShardedRun(contigs=c3["codes"], ploidy=2, bucket_threshold=5, strict_upstream_asserts=False, n_virtual=4,
                      exchange=exchange, fabric_timeout_s=30.0)
    many.synth_coverage(**C3_SYNTH)
    assert any(s.start > 0 for segs in many.plan for s in segs), "no contig was split"
    # replay what the one-handle run has seen (its state came from the previous test; replay is cheap)
    for rb in c3["batches"]:
        pd = H.parse_batch(rb.paf_text, {}, False)
        H.product_step(many, pd, rb.seqs)
    if one.batch == 0:
        for rb in c3["batches"]:
            pd = H.parse_batch(rb.paf_text, {}, False)
            H.product_step(one, pd, rb.seqs)
    assert one.last.switched_on and many.last.switched_on
    assert one.threshold == many.threshold and one.last.ubar0 == many.last.ubar0
    assert one.last.n_nonzero == many.last.n_nonzero and one.last.n_dropout == many.last.n_dropout
    for (name, a), b in zip(one.contigs_filt.items(), many.contigs_filt.values()):
        assert np.array_equal(a.strat, b.strat), f"{name}/{exchange}"
        assert np.array_equal(a.bucket_switches, b.bucket_switches), name
    del many
