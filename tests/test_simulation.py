"""The simulator's decision step against the reference's own outputs (tests/golden/sim_decisions.npz, written by
oracle/make_golden_sim.py from `BossRunsSim.make_decisions` / `filter_paf_dict` on the reference's real reads and
mappings): which record every read ends up with, who is accepted, the counts, the truncation of rejected reads.
CPU only — the lookups read `Contig.strat`, which is a host array."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

from boss_runs_b200.simulation import filter_paf_dict, make_decisions
from oracle.make_golden_sim import strategies

Z = dict(np.load(Path(__file__).resolve().parent / "golden" / "sim_decisions.npz", allow_pickle=False))


@pytest.mark.parametrize("nb,accept_unmapped", [(1, True), (1, False), (3, True)])
def test_decisions_match_the_reference(nb, accept_unmapped):
    tag = f"nb{nb}_{'acc' if accept_unmapped else 'rej'}_"
    rids = [str(r) for r in Z["rids"]]
    seqs = {r: "A" * int(n) for r, n in zip(rids, Z["read_len"])}
    lengths = {str(c): int(n) for c, n in zip(Z["contigs"], Z["contig_len"])}
    strat = strategies(lengths, nb, int(Z["seed"]) + nb)
    contigs = {k: SimpleNamespace(strat=v) for k, v in strat.items() if k != str(Z["missing"])}
    barcodes = {r: int(b) for r, b in zip(rids, Z[tag + "barcodes"])}
    paf_dict, reads_decision, n_mapped, n_unmapped, n_acc, n_rej = make_decisions(
        contigs, seqs, Z["paf_full"].tobytes().decode(), Z["paf_trunc"].tobytes().decode(), barcodes,
        accept_unmapped=accept_unmapped, all_read_ids=set(rids))
    assert [n_mapped, n_unmapped, n_acc, n_rej] == Z[tag + "counts"].tolist()
    assert list(paf_dict) == [str(k) for k in Z[tag + "keys"]]                      # same reads, same order
    assert [len(v) for v in paf_dict.values()] == Z[tag + "n_recs"].tolist()
    recs = [v[0] for v in paf_dict.values()]
    got = np.array([(r.qlen, r.qstart, r.qend, r.tstart, r.tend, r.rev, r.mapq, r.align_score) for r in recs], dtype=np.int64)
    assert np.array_equal(got, Z[tag + "rec"])
    assert [str(r.tname) for r in recs] == [str(t) for t in Z[tag + "rec_tname"]]
    assert [-1 if r.barcode is None else int(r.barcode) for r in recs] == Z[tag + "rec_barcode"].tolist()
    assert [len(reads_decision[r]) for r in rids] == Z[tag + "decision_len"].tolist()
    assert list(filter_paf_dict(paf_dict)) == [str(k) for k in Z[tag + "accepted_keys"]]
    assert 0 < n_acc < len(rids) and any(str(r.tname) == str(Z["missing"]) for r in recs)


def test_accepted_read_without_full_mapping_raises_like_upstream():
    line = "r1\t400\t0\t400\t+\tctg\t200000\t5000\t5400\t400\t400\t60\tAS:i:400\ttp:A:P\n"
    contigs = {"ctg": SimpleNamespace(strat=np.ones((2000, 2, 1), dtype=bool))}
    with pytest.raises(IndexError):                                 # choose_best_mapper([]) upstream (paf.py:721)
        make_decisions(contigs, {"r1": "A" * 900}, "", line, {"r1": 0})
    with pytest.raises(KeyError):                                   # barcodes[rec.qname]
        make_decisions(contigs, {"r1": "A" * 900}, line, line, {})
    # a start beyond the mask (IndexError) and an unknown target (KeyError) are plain rejections
    far = line.replace("\t5000\t5400\t", "\t250000\t250400\t")
    other = line.replace("\tctg\t", "\telsewhere\t")
    for text in (far, other):
        pd, dec, n_mapped, n_unmapped, n_acc, n_rej = make_decisions(contigs, {"r1": "A" * 900}, line, text, {"r1": 0})
        assert (n_acc, n_rej) == (0, 1) and len(dec["r1"]) == 400 and pd["r1"][0].qlen == 400


# ---- the whole simulated flow on the reference's real data --------------------------------------------------------
import hashlib  # noqa: E402

import helpers as H  # noqa: E402
from golden_io import GOLDEN, load_case, unpack2bit  # noqa: E402

FLOW = dict(np.load(Path(__file__).resolve().parent / "golden" / "sim_flow_zymo.npz", allow_pickle=False))


def flow_inputs():
    g = load_case("real_zymo")
    all_ids = set(r for _, seqs, _ in g.batches for r in seqs)
    truncs = [FLOW[f"b{bi}_paf_trunc"].tobytes().decode() for bi in range(len(g.batches))]
    return g, all_ids, truncs


def test_simulated_flow_oracle_equals_reference():
    """Oracle + this package's decision step reproduce what the reference's BossRuns produced when driven by its own
    `make_decisions` (oracle/make_golden_sim.py:main_flow): counts, accepted reads, coverage, thresholds and every
    strategy bit, batch after batch — including batches where most reads are rejected and contribute only their
    first 400 bases (Q12)."""
    g, all_ids, truncs = flow_inputs()
    orc = H.oracle_run(g.records, 1, [], None, 0)
    for bi, (paf, seqs, _) in enumerate(g.batches):
        updated, counts, acc, dec = H.oracle_sim_step(orc, seqs, paf, truncs[bi], {r: 0 for r in seqs}, all_ids)
        p = f"b{bi}_"
        assert updated
        assert list(counts) == FLOW[p + "counts"].tolist()
        assert list(acc) == [str(k) for k in FLOW[p + "accepted_keys"]]
        assert [len(dec[r]) for r in seqs] == FLOW[p + "decision_len"].tolist()
        assert np.array_equal(orc.rl.approx_ccl, FLOW[p + "approx_ccl"])
        for name, c in orc.contigs_filt.items():
            assert hashlib.sha256(np.ascontiguousarray(c.coverage).tobytes()).hexdigest() == str(FLOW[f"{p}{name}_coverage_sha"]), name
            assert np.array_equal(np.packbits(c.strat.ravel()), FLOW[f"{p}{name}_strat"]), f"b{bi}/{name}"
    assert FLOW["b2_counts"][3] > FLOW["b2_counts"][2] > 0               # mostly rejections by the third batch


@pytest.mark.gpu
@pytest.mark.parametrize("text_decisions", [False, True])
def test_simulated_flow_gpu_equals_oracle(lib, text_decisions):
    """`BossRunsSim.process_batch_runs_sim` on the GPU against the oracle flow above, state by state, and against the
    reference's recorded strategies."""
    from boss_runs_b200.simulation import BossRunsSim
    g, all_ids, truncs = flow_inputs()
    orc = H.oracle_run(g.records, 1, [], None, 0)
    sim = BossRunsSim(contigs=g.records, ploidy=1, bucket_threshold=0, write_debug=True, text_decisions=text_decisions)
    for bi, (paf, seqs, _) in enumerate(g.batches):
        updated, counts, acc, dec_o = H.oracle_sim_step(orc, seqs, paf, truncs[bi], {r: 0 for r in seqs}, all_ids)
        dec_p = sim.process_batch_runs_sim(seqs, None, {r: "" for r in seqs}, paf, truncs[bi], all_read_ids=all_ids)
        assert (sim.n_accepted, sim.n_rejected) == counts[2:]
        assert dec_p == dec_o
        assert bool(sim.last.switched_on) == updated
        H.compare_state(sim, orc, updated, f"sim/b{bi}")
        for name, c in sim.contigs_filt.items():
            assert np.array_equal(np.packbits(np.asarray(c.strat).ravel()), FLOW[f"b{bi}_{name}_strat"]), f"b{bi}/{name}"


# ---- decisions + batch in one C pass over the PAF texts (fastconv.decide_text) --------------------------------------
def _rows(b):
    return [(int(b.contig[i]), int(b.tstart[i]), int(b.tend[i]), int(b.barcode[i]), int(b.rev[i]), b.cigar_bytes(i), b.slice_bytes(i))
            for i in range(len(b))]


@pytest.mark.parametrize("nb,accept_unmapped", [(1, False), (1, True), (3, False), (3, True)])
def test_text_decisions_equal_object_decisions(nb, accept_unmapped):
    """On the reference's real reads with their full-length and truncated records (CIGARs included), against seeded
    strategies: the C pass yields the same batch rows, accepted flags, read lengths, truncations and counts as
    `make_decisions` + `filter_paf_dict` + the Python `convert_records` loop."""
    from boss_runs_b200 import build
    from boss_runs_b200.runs import CoverageConverter
    from boss_runs_b200.simulation import decide_text
    build.build_fastconv()
    g, all_ids, truncs = flow_inputs()
    tracked = [n for n, s in g.records if len(s) >= 100_000]
    lengths = {n: len(s) for n, s in g.records if n in tracked}
    cc = CoverageConverter({n: i for i, n in enumerate(tracked[:1])})        # the second tracked contig is "not tracked" here
    rng = np.random.default_rng(nb)
    for bi, (paf, seqs, _) in enumerate(g.batches):
        strat = strategies(lengths, nb, 20 + bi + nb)
        contigs = {k: SimpleNamespace(strat=v) for k, v in strat.items()}
        if bi == 1:
            contigs[tracked[0]] = SimpleNamespace(strat=np.zeros(1, dtype=bool))     # a reject ref's zeros(1): IndexError upstream -> reject
        barcodes = {r: int(rng.integers(nb)) for r in seqs}
        paf_dict, dec, n_mapped, n_unmapped, n_acc, n_rej = make_decisions(
            contigs, seqs, paf, truncs[bi], barcodes, accept_unmapped=accept_unmapped, all_read_ids=all_ids)
        acc = filter_paf_dict(paf_dict)
        want = cc._convert_records_py(paf_dict, seqs)
        got = decide_text(cc.contig_index, contigs, seqs, paf, truncs[bi], barcodes, accept_unmapped=accept_unmapped)
        assert got is not None
        b, row_acc, acc_qlen, dec2, n_mapped2, mapped, n_acc2, n_rej2 = got
        assert _rows(b) == _rows(want) and b.n_skipped == want.n_skipped and len(b) > 100
        assert (n_mapped2, n_acc2, n_rej2) == (n_mapped, n_acc, n_rej) and len(all_ids - mapped) == n_unmapped
        assert dec2 == dec
        assert acc_qlen == [r[0].qlen for r in acc.values()]
        on_tracked = [rid for rid, recs in paf_dict.items() if recs[0].tname in cc.contig_index]
        assert row_acc.tolist() == [rid in acc for rid in on_tracked]
        assert n_acc < len(seqs) and (n_acc > 0 or bi == 1)


def test_text_decisions_errors_like_object_path():
    from boss_runs_b200.runs import CoverageConverter
    from boss_runs_b200.simulation import decide_text
    line = "r1\t400\t0\t400\t+\tctg\t200000\t5000\t5400\t400\t400\t60\tAS:i:400\ttp:A:P\tcg:Z:400M\n"
    contigs = {"ctg": SimpleNamespace(strat=np.ones((2000, 2, 1), dtype=bool))}
    cc = CoverageConverter({"ctg": 0})
    with pytest.raises(IndexError):                                 # accepted, but never mapped in full
        decide_text(cc.contig_index, contigs, {"r1": "A" * 900}, "", line, {"r1": 0})
    with pytest.raises(KeyError):                                   # barcodes[rec.qname]
        decide_text(cc.contig_index, contigs, {"r1": "A" * 900}, line, line, {})
    far = line.replace("\t5000\t5400\t", "\t250000\t250400\t")
    b, row_acc, acc_qlen, dec, n_mapped, mapped, n_acc, n_rej = decide_text(cc.contig_index, contigs, {"r1": "A" * 900}, line, far, {"r1": 0})
    assert (n_acc, n_rej, len(dec["r1"]), row_acc.tolist(), acc_qlen) == (0, 1, 400, [False], [])
    # tend == 0 on the reverse strand: start = -1 -> row -1 wraps to the last row, like NumPy
    wrap = line.replace("\t+\t", "\t-\t").replace("\t5000\t5400\t", "\t-400\t0\t")
    last_off = {"ctg": SimpleNamespace(strat=np.ones((2000, 2, 1), dtype=bool))}
    last_off["ctg"].strat[-1, 1, 0] = False
    full900 = line.replace("r1\t400\t0\t400", "r1\t900\t0\t400")
    for cont, want_acc in ((contigs, 1), (last_off, 0)):
        pd, *_rest, a, r = make_decisions(cont, {"r1": "A" * 900}, full900, wrap, {"r1": 0})
        got = decide_text(cc.contig_index, cont, {"r1": "A" * 900}, full900, wrap, {"r1": 0})
        assert (a, got[6]) == (want_acc, want_acc)


# ---- BASELINE config 1 at its stated size: upstream's own simulator run recorded by oracle/make_golden_c1.py ---------------
def _c1_full():
    path = GOLDEN / "c1_full.npz"
    if not path.exists():
        pytest.skip("tests/golden/c1_full.npz not generated (python -m oracle.make_golden_c1)")
    z = np.load(path, allow_pickle=False)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    records = []
    for name, L in zip(z["ref_names"], z["ref_lengths"]):
        records.append((str(name), acgt[unpack2bit(z[f"ref_{name}"], int(L))].tobytes().decode()))
    lens = z["pool_len"]
    allb = acgt[unpack2bit(z["pool_2bit"], int(lens.sum()))].tobytes().decode()
    off = np.concatenate(([0], np.cumsum(lens)))
    pool = {str(r): allb[off[i]: off[i + 1]] for i, r in enumerate(z["pool_rids"])}
    return z, records, pool


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["A", "B"])
def test_c1_full_size_against_upstream_simulator(tag, lib):
    """The whole of data/BOSS_test_data (zymo.fa: 9 tracked contigs, 31 012 581 sites; ERR3152366 reads with their full-length
    and mu-truncated minimap2 records) through `BossRunsSim.process_batch_runs_sim`, batch for batch what upstream's sampler
    handed out, against what upstream's `BossRunsSim` computed: run A = batchsize 4000, maxb 1 (BASELINE.json configs[0]),
    run B = 1000 x 8. Decision counts, read-length staircase, counters (digest), dropout zeros, score sums, threshold and
    every mask bit; run A additionally state by state against the oracle."""
    from boss_runs_b200.simulation import BossRunsSim
    z, records, pool = _c1_full()
    assert int(z[f"{tag}_n_sites"]) == 31_012_581 and len([1 for _, s in records if len(s) >= 100_000]) == 9   # test_reference.py:51-67
    sim = BossRunsSim(contigs=records, ploidy=1, bucket_threshold=0, write_debug=(tag == "A"))
    assert int(sim.ref.n_sites) == int(z[f"{tag}_n_sites"]) and list(sim.contigs) == [str(x) for x in z[f"{tag}_contigs"]]
    orc = H.oracle_run(records, 1, [], None, 0) if tag == "A" else None
    for bi in range(int(z[f"{tag}_maxb"])):
        p = f"{tag}{bi}_"
        seqs = {str(r): pool[str(r)] for r in z[p + "rids"]}
        assert len(seqs) == int(z[f"{tag}_batchsize"])
        paf_f, paf_t = z[p + "paf_full"].tobytes().decode(), z[p + "paf_trunc"].tobytes().decode()
        if orc is not None:
            updated, counts, acc, dec_o = H.oracle_sim_step(orc, seqs, paf_f, paf_t, {r: 0 for r in seqs}, set(seqs))
        dec_p = sim.process_batch_runs_sim(seqs, None, {r: "" for r in seqs}, paf_f, paf_t)
        n_mapped, n_unmapped, n_acc, n_rej = (int(x) for x in z[p + "counts"])
        assert (sim.n_accepted, sim.n_rejected) == (n_acc, n_rej), f"{p}: decisions differ from upstream's"
        assert sum(len(s) for s in dec_p.values()) > 0
        assert np.array_equal(sim.rl_dist.approx_ccl, z[p + "approx_ccl"]) and sim.rl_dist.time_cost == float(z[p + "time_cost"])
        assert bool(sim.last.switched_on) == bool(z[p + "updated"])
        if sim.last.switched_on:
            thr = float(z[p + "threshold"])
            assert abs(sim.threshold - thr) <= H.tol.THRESHOLD_RTOL * thr, f"{p}: threshold {sim.threshold!r} vs upstream {thr!r}"
            assert abs(sim.last.normaliser - float(z[p + "normaliser"])) <= H.tol.NORM_RTOL * float(z[p + "normaliser"])
        n_drop = 0
        for name, c in sim.contigs_filt.items():
            q = f"{p}{name}_"
            cov = c.coverage
            assert hashlib.sha256(np.ascontiguousarray(cov).tobytes()).hexdigest() == str(z[q + "coverage_sha"]), f"{q}: counters"
            assert int(cov.sum(dtype=np.int64)) == int(z[q + "depth_total"])
            s = c.scores
            assert int(np.count_nonzero(s == 0.0)) == int(z[q + "n_dropout"]), f"{q}: dropout zeros"
            n_drop += int(z[q + "n_dropout"])
            assert abs(float(s.sum()) - float(z[q + "scores_sum"])) <= 1e-9 * abs(float(z[q + "scores_sum"])), f"{q}: scores"
            assert np.array_equal(np.packbits(np.asarray(c.bucket_switches).ravel()), z[q + "switches"]), f"{q}: bucket switches"
            got, want = np.packbits(np.asarray(c.strat).ravel()), z[q + "strat"]
            assert np.array_equal(got, want), f"{q}: {int(np.unpackbits(got ^ want).sum())} mask bits differ from upstream's"
        assert sim.last.n_dropout == n_drop
        if orc is not None:
            assert counts == (n_mapped, n_unmapped, n_acc, n_rej) and dec_p == dec_o and updated
            H.compare_state(sim, orc, updated, f"c1/{p}")
