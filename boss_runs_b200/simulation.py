"""The simulator's decision step in front of the strategy update (first half).

`boss.runs.simulation.BossRunsSim` samples reads, decides for each one — from the strategy of the previous update —
whether it would have been sequenced to the end or ejected after `mu` bases, and feeds the outcome into the same update
as a live run (boss/runs/simulation.py:37-190). The sampler, the read cache and the pseudo-time bookkeeping stay
upstream's (host control plane); this module mirrors the part that touches the path: `make_decisions`,
`filter_paf_dict` and the call sequence of `process_batch_runs_sim`, on top of the GPU-backed `BossRuns`. Lookups go
to `Contig.strat`, the pinned host mirror the distribution kernel keeps current.
Pinned: tests/test_simulation.py replays tests/golden/sim_decisions.npz, written by oracle/make_golden_sim.py from the
upstream functions themselves on the reference's own reads and mappings.
"""
from __future__ import annotations

import logging
from collections import defaultdict
from io import StringIO

from .hostmodel import choose_best_mapper, parse_PAF
from .runs import BossRuns


def make_decisions(contigs_filt: dict, seqs: dict[str, str], paf_full: str, paf_trunc: str, barcodes: dict[str, int],
                   mu: int = 400, accept_unmapped: bool = False, all_read_ids: set | None = None, window: int = 100):
    """`BossRunsSim.make_decisions` (simulation.py:37-120): -> (paf_dict, reads_decision, n_mapped, n_unmapped,
    n_accepted, n_rejected). `contigs_filt[name].strat` is bool `(L//100, 2, nb)`; `all_read_ids` stands for
    `sampler.fq_stream.read_ids` (default: the ids of this batch)."""
    paf_dict = defaultdict(list)
    mapped_reads = set()
    n_rejected = n_accepted = 0
    reads_decision = dict(seqs)                                      # deepcopy of a dict of immutable strings
    paf_dict_full = parse_PAF(StringIO(paf_full))
    paf_dict_trunc = parse_PAF(StringIO(paf_trunc))
    for rid, rlist in paf_dict_trunc.items():                        # decisions come from the mu-sized mappings
        rec = choose_best_mapper(rlist)[0]
        rec.barcode = barcodes[rec.qname]
        mapped_reads.add(rid)
        start_pos = rec.tend - 1 if rec.rev else rec.tstart
        try:
            strat = contigs_filt[str(rec.tname)].strat
            decision = strat[start_pos // window, rec.rev, barcodes[rec.qname]]
        except (KeyError, IndexError):
            decision = 0                                             # no strategy for that target: reject
        if decision:
            rec_full = choose_best_mapper(paf_dict_full[str(rec.qname)])[0]      # IndexError upstream if it never mapped in full
            rec_full.barcode = barcodes[rec_full.qname]
            paf_dict[str(rec.qname)].append(rec_full)
            n_accepted += 1
        else:
            paf_dict[str(rec.qname)].append(rec)
            n_rejected += 1
            reads_decision[rid] = reads_decision[rid][:mu]
    for read_id, seq in seqs.items():                                # unmapped reads are accepted or rejected wholesale
        if read_id in mapped_reads:
            continue
        if accept_unmapped:
            reads_decision[read_id] = seq
            if read_id in paf_dict_full:
                paf_dict[read_id].append(choose_best_mapper(paf_dict_full[read_id])[0])
            n_accepted += 1
        else:
            reads_decision[read_id] = seq[:mu]
            n_rejected += 1
    ids = set(seqs) if all_read_ids is None else all_read_ids
    return paf_dict, reads_decision, len(mapped_reads), len(ids - mapped_reads), n_accepted, n_rejected


def filter_paf_dict(paf_dict: dict, mu: int = 400) -> dict:
    """`BossRunsSim.filter_paf_dict` (simulation.py:123-135): accepted reads are those whose record is not mu long."""
    return {rid: recs for rid, recs in paf_dict.items() if recs[0].qlen != mu}


class BossRunsSim(BossRuns):
    """`BossRuns` driven by sampled reads with accept/reject decisions taken from the current strategy."""

    def __init__(self, *args, accept_unmapped: bool = False, mu: int = 400, **kw):
        super().__init__(*args, **kw)
        self.accept_unmapped, self.mu = accept_unmapped, mu
        self.n_accepted = self.n_rejected = 0

    def make_decisions(self, seqs, paf_full, paf_trunc, barcodes, window: int = 100, all_read_ids=None):
        return make_decisions(self.contigs_filt, seqs, paf_full, paf_trunc, barcodes, mu=self.mu,
                              accept_unmapped=self.accept_unmapped, all_read_ids=all_read_ids, window=window)

    def filter_paf_dict(self, paf_dict):
        return filter_paf_dict(paf_dict, mu=self.mu)

    def process_batch_runs_sim(self, read_seqs: dict[str, str], read_quals: dict[str, str] | None,
                               read_barcodes_names: dict, paf_f: str, paf_t: str, all_read_ids=None):
        """simulation.py:139-190 with the sampler's output handed in: decisions, read-length update from the accepted
        reads, then the update proper — coverage from every record (rejected reads contribute their first mu bases,
        Q12), read starts from the accepted ones. Returns `reads_decision` for the caller's read cache."""
        self._prescore_begin()                                       # the GPU scores while the host decides and converts
        read_barcodes = {rid: self.barcodes_index.get(bc, 0) for rid, bc in read_barcodes_names.items()}
        paf_dict, reads_decision, n_mapped, n_unmapped, n_acc, n_rej = self.make_decisions(
            seqs=read_seqs, paf_full=paf_f, paf_trunc=paf_t, barcodes=read_barcodes, all_read_ids=all_read_ids)
        logging.info(f"mapped {n_mapped}, not mapped {n_unmapped}")
        logging.info(f"accepted {n_acc}, rejected {n_rej}")
        self.n_accepted, self.n_rejected = n_acc, n_rej
        paf_dict_acc = self.filter_paf_dict(paf_dict)
        self.rl_dist.update(read_lengths={n: r[0].qlen for n, r in paf_dict_acc.items()})
        increments = self.cc.convert_records(paf_dict=paf_dict, seqs=read_seqs, quals=read_quals, barcodes=read_barcodes)
        self._prescore(increments)
        self._effect_increments(increments=increments)
        self.count_read_starts(paf_dict_acc)
        self.update_wrapper()
        self.batch += 1
        return reads_decision
