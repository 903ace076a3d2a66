"""Shared drivers for the parity tests: run the same batches through the oracle and through the product."""
from __future__ import annotations

import io

import numpy as np

from oracle import boss_oracle as bo
from boss_runs_b200.hostmodel import parse_PAF
import tolerances as tol


def parse_batch(paf_text, bcs, barcoded):
    pd = parse_PAF(io.StringIO(paf_text))
    for rid, recs in pd.items():
        for r in recs:
            r.barcode = bcs.get(rid) if barcoded else None
    return pd


def oracle_run(records, ploidy, reject_refs, barcodes, bucket_threshold, strict=True):
    return bo.OracleRun(records, ploidy=ploidy, reject_refs=reject_refs, barcodes=barcodes, bucket_threshold=bucket_threshold)


def oracle_step(run, pd, seqs):
    run.rl.update({rid: recs[0].qlen for rid, recs in pd.items()})
    run.ingest(pd, seqs)
    run.read_starts.count(pd)
    return run.update()


def product_run(records, ploidy, reject_refs, barcodes, bucket_threshold, **kw):
    from boss_runs_b200.runs import BossRuns
    return BossRuns(contigs=records, ploidy=ploidy, barcodes=barcodes,
                    reject_refs=",".join(reject_refs) if reject_refs else None, bucket_threshold=bucket_threshold,
                    write_debug=True, **kw)


def product_step(run, pd, seqs):
    run.rl_dist.update({rid: recs[0].qlen for rid, recs in pd.items()})
    run.process_batch_runs(pd, seqs)
    return bool(run.last.switched_on)


def assert_close_smooth(got, want, what):
    atol = tol.SMOOTH_ATOL_FRAC * float(np.max(np.abs(want))) if want.size else 0.0
    err = np.abs(got - want)
    bad = err > (tol.SMOOTH_RTOL * np.abs(want) + atol)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} outside tolerance; worst abs err {err.max():.3e} (atol {atol:.3e})"


def assert_masks_match(got, want, benefit, thr, what):
    """Masks must agree everywhere except where the benefit is within MASK_REL of the threshold."""
    diff = got != want
    if diff.any():
        near = np.abs(benefit - thr) <= tol.MASK_REL * abs(thr)
        assert not (diff & ~near).any(), f"{what}: {(diff & ~near).sum()} mask bits differ away from the threshold"


def compare_state(prod, orc, updated: bool, tag: str):
    """Product (GPU) vs oracle after one update."""
    for (name, pc), (oname, oc) in zip(prod.contigs.items(), orc.contigs.items()):
        assert name == oname
        if oc.rej:
            assert pc.strat.shape == (1,) and not pc.strat.any()
            continue
        t = f"{tag}/{name}"
        assert np.array_equal(pc.coverage, oc.coverage), f"{t}: coverage differs"
        s = pc.scores
        np.testing.assert_allclose(s, oc.scores, rtol=tol.SCORE_RTOL, atol=0, err_msg=f"{t}: scores")
        assert np.array_equal(s == 0.0, oc.scores == 0.0), f"{t}: dropout zeros differ"
        assert np.array_equal(pc.bucket_switches, oc.bucket_switches), f"{t}: bucket switches"
        assert np.array_equal(pc.switched_on, oc.switched_on), f"{t}: switched_on"
        if updated:
            np.testing.assert_allclose(pc.scores_ds, oc.scores_ds, rtol=tol.SCORE_RTOL, atol=0, err_msg=f"{t}: scores_ds")
            assert_close_smooth(pc.smu, oc.smu, f"{t}: smu")
            assert_close_smooth(pc.expected_benefit, oc.expected_benefit, f"{t}: expected_benefit")
            assert_close_smooth(pc.additional_benefit, oc.additional_benefit, f"{t}: additional_benefit")
    if updated:
        assert abs(prod.threshold - orc.threshold) <= tol.THRESHOLD_RTOL * abs(orc.threshold), \
            f"{tag}: threshold {prod.threshold!r} vs oracle {orc.threshold!r}"
        # strategies: Q2 means contig k reads merged rows shifted by k bins; check against the oracle's own arrays
        i = 0
        for (name, pc), oc in zip(prod.contigs_filt.items(), orc.contigs_filt.values()):
            n = oc.length // 100
            ben = orc.benefit_adj[i: i + n]
            assert_masks_match(pc.strat, oc.strat, ben, orc.threshold, f"{tag}/{name}: strat")
            i += n
        compare_hist_stage(prod, orc, tag)


def compare_hist_stage(prod, orc, tag: str):
    """The histogram stage of the threshold derivation (sequences.py:584-630) against the oracle's own intermediates:
    normaliser, per-exponent counts and F-hat mass, ubar0, and the F-hat array the device expanded by itself."""
    from boss_runs_b200._lib import HIST_BINS
    d = orc.diag
    norm = float(d["normaliser"])
    assert abs(prod.last.normaliser - norm) <= tol.NORM_RTOL * norm, f"{tag}: normaliser {prod.last.normaliser!r} vs {norm!r}"
    # F-hat rows as k_hist reads them (expansion x20, tail fixes and normalisation happen on the device)
    target = orc.fhat_adj.shape[0]
    got_f = prod.engine.fhat_rows(0, target)
    for b in range(orc.fhat_adj.shape[2]):
        np.testing.assert_allclose(got_f, orc.fhat_adj[:, :, b], rtol=tol.FHAT_RTOL, atol=0, err_msg=f"{tag}: device F-hat")
    counts, f_grid = prod.engine.hist()
    want_c = np.zeros(HIST_BINS, dtype=np.int64)
    want_f = np.zeros(HIST_BINS)
    want_c[d["exponents"]] = d["counts"]
    want_f[d["exponents"]] = d["f_grid"]
    # an oracle entry may fall either side of a bin edge (a power of two times the normaliser) when it lies within the
    # tolerance of the smoothed arrays of it: relative SMOOTH_RTOL plus the absolute floor SMOOTH_ATOL_FRAC * max
    flat = orc.benefit_adj.flatten("F")
    nz = flat[flat != 0]
    ratio = nz / norm
    m, e = np.frexp(ratio)
    width = tol.SMOOTH_RTOL * ratio + tol.SMOOTH_ATOL_FRAC
    near = (ratio - np.ldexp(0.5, e) <= width) | (np.ldexp(1.0, e) - ratio <= width)
    e = np.abs(e)
    slack = np.zeros(HIST_BINS + 1, dtype=np.int64)
    if near.any():
        k = np.bincount(e[near], minlength=HIST_BINS)[:HIST_BINS]
        slack[:HIST_BINS] += k
        slack[1:HIST_BINS + 1] += k
        slack[:HIST_BINS - 1] += k[1:]
    H = tol.HIST_HEAD
    bad = np.abs(counts[:H] - want_c[:H]) > slack[:H]
    assert not bad.any(), f"{tag}: histogram counts differ at exponents {np.nonzero(bad)[0]}: {counts[:H][bad]} vs {want_c[:H][bad]}"
    clean = slack[:H] == 0
    np.testing.assert_allclose(f_grid[:H][clean], want_f[:H][clean], rtol=tol.HIST_F_RTOL, atol=0, err_msg=f"{tag}: f_grid")
    # tail: whatever the GPU still counts there must be (near-)zero in the oracle as well
    n_small = int((flat < np.ldexp(norm, -(H - 1))).sum())
    assert int(counts[H:].sum()) <= n_small, f"{tag}: {int(counts[H:].sum())} tail entries on the GPU, {n_small} near-zero in the oracle"
    assert abs(prod.last.ubar0 - float(d["ubar0"])) <= tol.UBAR_RTOL * abs(float(d["ubar0"])), \
        f"{tag}: ubar0 {prod.last.ubar0!r} vs {float(d['ubar0'])!r}"


def oracle_sim_step(orc, seqs, paf_full, paf_trunc, barcodes, all_ids):
    """One simulated batch on the oracle (simulation.py:139-190 minus sampler / read cache): decisions from the oracle's
    current strategies, read lengths and read starts from the accepted reads, coverage from every record."""
    from boss_runs_b200.simulation import filter_paf_dict, make_decisions
    paf_dict, reads_decision, n_mapped, n_unmapped, n_acc, n_rej = make_decisions(
        orc.contigs_filt, seqs, paf_full, paf_trunc, barcodes, all_read_ids=all_ids)
    acc = filter_paf_dict(paf_dict)
    orc.rl.update({n: r[0].qlen for n, r in acc.items()})
    orc.ingest(paf_dict, seqs)
    orc.read_starts.count(acc)
    updated = orc.update()
    return updated, (n_mapped, n_unmapped, n_acc, n_rej), acc, reads_decision
