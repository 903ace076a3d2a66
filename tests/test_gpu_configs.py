"""GPU parity at the sizes BASELINE.json's configs name, beyond the small golden cases
(config 1 at full size lives in tests/test_simulation.py::test_c1_full_size_against_upstream_simulator):

* C2 at full size (4.6 Mb haploid, 4000 x ~10 kb reads, three consecutive updates) against the oracle, state by state;
* C4 at full size (24 barcodes x 5 Mb, unclassified reads falling to index 0 — Q11; the whole-row rules Q6/Q8 across
  24 planes) against the oracle, state by state;
* C5: a miniature (400 contigs, all of it against the oracle) and 1 000 contigs 10 kb - 5 Mb (~570 tracked, ~0.8 Gb): per-contig
  stages on a sample of whole contigs against the oracle, the global stage (merge, F-hat, histogram, threshold, distribution
  with the Q2 row shift) restated with the oracle's functions on all contigs;
* C3 at full size (3.1 Gb diploid, 25 contigs, 35 GB of state): two whole contigs (0.5 Mb and 50.8 Mb, the first one 3.1e9
  sites into the arrays) against the oracle run on those contigs alone (per-contig stages are independent upstream,
  core.py:83-121), the global stage restated on all 31 M merged rows, an update over an empty batch is idempotent, and
  four virtual shards give bit-identical thresholds and masks to the one-handle run.
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


def test_c4_full_size_24_barcodes_x_5mb(lib):
    """BASELINE config 4 as specified: 24 barcodes x one 5 Mb reference in one launch, 4000 reads of ~10 kb per batch
    spread over the barcodes, ~5 % of them unclassified (index 0, Q11); the whole-row rules (Q6, Q8) act across 24 planes.
    Three batches with a pile-up window so that frozen sites and bucket switches occur (the depth rule stays inactive at
    0.3x per barcode and batch; the small barcoded golden cases cover it)."""
    contigs = synth.random_contigs({"amplicon_ref": 5_000_000}, seed=17)
    barcodes = [f"barcode{i + 1:02d}" for i in range(24)]
    batches = [synth.read_batch(contigs, n_reads=4000, seed=400 + b, n_barcodes=24, focus=("amplicon_ref", 1_200_000, 1_260_000, 0.3))
               for b in range(3)]
    prod, orc = _run_both(list(contigs.items()), 1, [], barcodes, 0.2, batches, "c4")
    cov = orc.contigs["amplicon_ref"].coverage
    assert all(cov[:, :, b].any() for b in range(24))
    assert (cov.sum(axis=1) >= 30).any() and orc.contigs["amplicon_ref"].bucket_switches.any()


def test_c5_shape_many_contigs(lib):
    """BASELINE config 5 in miniature, all of it against the oracle: 400 contigs log-uniform 10 kb - 600 kb, seed 13; under
    100 kb dropped by the loader (reference.py:319,330), 10 % of the names in reject_refs (-> `(1,)` masks and 4 phantom
    sites each), contig k's mask read k bins early (Q2)."""
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
# sizes the oracle cannot hold whole: per-contig stages on a sample of whole contigs + the global stage restated
# ---------------------------------------------------------------------------------------------------------
def _oracle_contig(name, codes, nb=1):
    hap = bo.ScoreModel(1)
    return bo.ContigState(name, codes, nb=nb, score0=hap.score0, ent0=hap.ent0)


def _advance_and_compare(run, oc, pd, seqs, model, first_update_scores_observed, tag):
    """One oracle contig through the per-contig stages of an update (core.py:83-121 are per-contig loops) on the reads of
    the batch that map to it, compared with the same contig inside the big GPU run."""
    pc = run.contigs[oc.name]
    inc = bo.convert_records(pd, seqs).get(oc.name, [])
    bo.increment_coverage(oc, inc)
    if first_update_scores_observed is not None:
        oc.change_mask[:, 0] |= first_update_scores_observed
    bo.update_scores(oc, model)
    bo.modify_scores(oc)
    bo.check_buckets(oc, run.bucket_threshold)
    bo.calc_smu(oc)
    bo.calc_u(oc, run.rl_dist.approx_ccl)
    assert np.array_equal(pc.coverage, oc.coverage), f"{tag}: coverage"
    s = pc.scores
    np.testing.assert_allclose(s, oc.scores, rtol=H.tol.SCORE_RTOL, atol=0, err_msg=f"{tag}: scores")
    assert np.array_equal(s == 0.0, oc.scores == 0.0), f"{tag}: dropout zeros"
    assert np.array_equal(pc.bucket_switches, oc.bucket_switches), f"{tag}: bucket switches"
    np.testing.assert_allclose(pc.scores_ds, oc.scores_ds, rtol=H.tol.SCORE_RTOL, atol=0, err_msg=f"{tag}: scores_ds")
    H.assert_close_smooth(pc.smu, oc.smu, f"{tag}: smu")
    H.assert_close_smooth(pc.expected_benefit, oc.expected_benefit, f"{tag}: expected_benefit")
    H.assert_close_smooth(pc.additional_benefit, oc.additional_benefit, f"{tag}: additional_benefit")
    return len(inc)


def _check_global_stage(run, masks_before, tag):
    """Everything after the per-contig stages (core.py:172-198), restated with the oracle's own functions on the benefit
    arrays the GPU produced: merge + adjust_length, F-hat from the read-start counts, exponent histogram, threshold, and
    the bucket-gated distribution with its row shift (Q2) — so threshold, histogram, F-hat and EVERY mask bit are checked
    at sizes where the oracle cannot hold the per-site state."""
    from types import SimpleNamespace
    filt = run.contigs_filt
    nb = run.nbarcodes
    ben = np.concatenate([c.additional_benefit for c in filt.values()])
    target = int(run.ref.n_sites) // 100
    ben_adj = bo.adjust_length(target, ben)
    rs = bo.ReadStarts(filt)
    rs.strict = False
    rs.counts = {name: np.array(run.read_starts.read_starts[name]) for name in filt}
    fh = np.repeat(rs.fhat()[:, :, np.newaxis], nb, axis=2)
    fh_adj = bo.adjust_length(target, fh)
    strat, thr, diag = bo.find_strategy(ben_adj, ben_adj, fh_adj, run.rl_dist.time_cost)      # Q1: smu := benefit
    assert abs(run.threshold - thr) <= H.tol.THRESHOLD_RTOL * thr, f"{tag}: threshold {run.threshold!r} vs restated {thr!r}"
    H.compare_hist_stage(run, SimpleNamespace(diag=diag, benefit_adj=ben_adj, fhat_adj=fh_adj), tag)
    i = 0
    for (name, c), before in zip(filt.items(), masks_before):
        rows = c.length // 100
        gate = bo.adjust_length(rows, np.repeat(np.asarray(c.bucket_switches), 200, axis=0))
        want = np.array(before)
        cs = strat[i: i + rows]
        for b in range(nb):
            want[gate[:, b], :, b] = cs[gate[:, b], :, b]
        assert np.array_equal(np.asarray(c.strat), want), f"{tag}/{name}: {(np.asarray(c.strat) != want).sum()} mask bits differ"
        i += rows
    return thr


def test_c5_full_size_thousand_contigs(lib):
    """BASELINE config 5 at half its contig count (the host holds the sequences): 1 000 contigs log-uniform 10 kb - 5 Mb
    (seed 13), 10 % of the names in reject_refs (`(1,)` masks, 4 phantom sites each), everything under 100 kb dropped by the
    loader (reference.py:319,330). ~570 tracked contigs, ~0.8 Gb: contig k's mask is read k bins early (Q2: up to ~570 bins),
    and F-hat's length drift exceeds one window — upstream itself asserts there (readstartdist.py:131), so the run uses
    strict_upstream_asserts=False and the oracle evaluates the same expressions without the assertion.
    Per-contig stages: 24 whole contigs across the size range against the oracle. Global stage: restated on all contigs."""
    rng = np.random.default_rng(13)
    n = 1000
    lens = np.exp(rng.uniform(np.log(10_000), np.log(5_000_000), size=n)).astype(np.int64)
    rej = set(int(i) for i in rng.choice(n, size=n // 10, replace=False))
    names = [f"bin{i:04d}" for i in range(n)]
    crng = np.random.default_rng(5)
    records = {names[i]: crng.integers(0, 4, size=int(lens[i]), dtype=np.uint8) for i in range(n) if lens[i] >= 100_000}
    tracked = {k: v for k, v in records.items() if int(k[3:]) not in rej}
    assert len(tracked) > 500
    from boss_runs_b200.runs import BossRuns
    run = BossRuns(contigs=records, ploidy=1, bucket_threshold=0, strict_upstream_asserts=False, write_debug=True,
                   reject_refs=",".join(k for k in records if int(k[3:]) in rej))
    assert list(run.contigs_filt) == list(tracked) and sum(1 for c in run.contigs.values() if c.rej) == len(records) - len(tracked)
    assert int(run.ref.n_sites) == sum(len(v) for v in tracked.values()) + 4 * (len(records) - len(tracked))
    order = sorted(tracked, key=lambda k: len(tracked[k]))
    sample = [order[0], order[-1], list(tracked)[0], list(tracked)[-1]] + [order[i] for i in range(7, len(order), len(order) // 20)]
    sample = list(dict.fromkeys(sample))
    ocs = {k: _oracle_contig(k, tracked[k]) for k in sample}
    model = bo.ScoreModel(1)
    model.build_table()
    for bi in range(2):
        rb = synth.read_batch(tracked, n_reads=4000, seed=500 + bi, codes=tracked, focus=(sample[2], 20_000, 60_000, 0.05))
        pd = H.parse_batch(rb.paf_text, {}, False)
        masks_before = [np.array(c.strat) for c in run.contigs_filt.values()]
        H.product_step(run, pd, rb.seqs)
        assert run.last.switched_on
        hit = sum(_advance_and_compare(run, oc, pd, rb.seqs, model, None, f"c5/b{bi}/{k}") > 0 for k, oc in ocs.items())
        assert hit >= 5
        _check_global_stage(run, masks_before, f"c5/b{bi}")
    assert any(not np.asarray(c.strat).all() for c in run.contigs_filt.values())


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


def test_c3_full_size_contigs_and_global_stage(c3):
    """Per-contig stages: the last contig of the 3.1 Gb run (0.5 Mb, 3.1e9 sites into the arrays) and a 50.8 Mb one in the
    middle, each against the oracle run on that contig alone with the same counters and the same reads. Global stage:
    merge, F-hat, exponent histogram, threshold and the shifted, bucket-gated distribution restated with the oracle's
    functions on all 31 M merged rows — every mask bit of every contig."""
    run, small, codes = c3["run"], c3["small"], c3["codes"]
    mid = c3["names"][21]
    assert len(codes[mid]) > 50_000_000
    model = bo.ScoreModel(2)
    model.build_table()
    ocs, observed = {}, {}
    for name in (small, mid):
        oc = _oracle_contig(name, codes[name])
        oc.coverage[...] = run.contigs[name].coverage    # the synthetic pre-loaded state, read back from the device
        observed[name] = oc.coverage.sum(axis=1)[:, 0] > 0
        ocs[name] = oc
    k = len(c3["names"]) - 1
    for bi, rb in enumerate(c3["batches"]):
        pd = H.parse_batch(rb.paf_text, {}, False)
        masks_before = [np.array(c.strat) for c in run.contigs_filt.values()]
        H.product_step(run, pd, rb.seqs)
        assert run.last.switched_on
        for name, oc in ocs.items():
            # the first update scores every observed site of the pre-loaded state
            n_inc = _advance_and_compare(run, oc, pd, rb.seqs, model, observed[name] if bi == 0 else None, f"c3/b{bi}/{name}")
            assert n_inc > 10 and (oc.scores == 0.0).any()
        thr = _check_global_stage(run, masks_before, f"c3/b{bi}")
        # Q2 spelled out on the last contig: its strategy row j is merged row (own start - k + j), rows k.. are its own benefit rows 0..
        pc, oc = run.contigs[small], ocs[small]
        gate = _gate(pc)[k:, 0]
        n = pc.strat.shape[0]
        want = (oc.additional_benefit >= thr)[: n - k]
        got = np.asarray(pc.strat)[k:]
        diff = (got != want)[gate]
        near = (np.abs(oc.additional_benefit[: n - k] - thr) <= H.tol.MASK_REL * thr)[gate]
        assert not (diff & ~near).any(), f"c3/b{bi}: {(diff & ~near).sum()} mask bits differ away from the threshold"
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
