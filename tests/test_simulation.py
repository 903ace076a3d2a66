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
from golden_io import load_case  # noqa: E402

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
def test_simulated_flow_gpu_equals_oracle(lib):
    """`BossRunsSim.process_batch_runs_sim` on the GPU against the oracle flow above, state by state, and against the
    reference's recorded strategies."""
    from boss_runs_b200.simulation import BossRunsSim
    g, all_ids, truncs = flow_inputs()
    orc = H.oracle_run(g.records, 1, [], None, 0)
    sim = BossRunsSim(contigs=g.records, ploidy=1, bucket_threshold=0, write_debug=True)
    for bi, (paf, seqs, _) in enumerate(g.batches):
        updated, counts, acc, dec_o = H.oracle_sim_step(orc, seqs, paf, truncs[bi], {r: 0 for r in seqs}, all_ids)
        dec_p = sim.process_batch_runs_sim(seqs, None, {r: "" for r in seqs}, paf, truncs[bi], all_read_ids=all_ids)
        assert (sim.n_accepted, sim.n_rejected) == counts[2:]
        assert dec_p == dec_o
        assert bool(sim.last.switched_on) == updated
        H.compare_state(sim, orc, updated, f"sim/b{bi}")
        for name, c in sim.contigs_filt.items():
            assert np.array_equal(np.packbits(np.asarray(c.strat).ravel()), FLOW[f"b{bi}_{name}_strat"]), f"b{bi}/{name}"
