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
