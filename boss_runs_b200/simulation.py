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

import numpy as np

from .hostmodel import choose_best_mapper, parse_PAF
from .runs import BossRuns, PackedBatch, _best_index, _fastconv


def _lookup(contigs_filt: dict, picks: list, barcodes: dict, window: int) -> np.ndarray:
    """Decision bit of every picked record: `strat[start // window, rev, barcode]` with NumPy's index rules (negative
    indices wrap, anything outside raises IndexError upstream -> reject; a target without a strategy raises KeyError ->
    reject). One gather per target instead of one Python lookup per read."""
    n = len(picks)
    out = np.zeros(n, dtype=bool)
    by_target: dict[str, list[int]] = defaultdict(list)
    for i, rec in enumerate(picks):
        by_target[str(rec.tname)].append(i)
    for tname, idx in by_target.items():
        cont = contigs_filt.get(tname)
        if cont is None:
            continue
        strat = np.asarray(cont.strat)
        if strat.ndim != 3:
            continue                                                 # `zeros(1)` of a reject ref: IndexError upstream
        rows_n, _, nb = strat.shape
        recs = [picks[i] for i in idx]
        row = np.array([((r.tend - 1) if r.rev else r.tstart) // window for r in recs], dtype=np.int64)
        rev = np.array([int(r.rev) for r in recs], dtype=np.int64)
        bc = np.array([barcodes[r.qname] for r in recs], dtype=np.int64)
        ok = (row >= -rows_n) & (row < rows_n) & (bc >= -nb) & (bc < nb)
        sel = np.asarray(idx)[ok]
        out[sel] = strat[row[ok], rev[ok], bc[ok]]
    return out


def make_decisions(contigs_filt: dict, seqs: dict[str, str], paf_full: str, paf_trunc: str, barcodes: dict[str, int],
                   mu: int = 400, accept_unmapped: bool = False, all_read_ids: set | None = None, window: int = 100):
    """`BossRunsSim.make_decisions` (simulation.py:37-120): -> (paf_dict, reads_decision, n_mapped, n_unmapped,
    n_accepted, n_rejected). `contigs_filt[name].strat` is bool `(L//100, 2, nb)`; `all_read_ids` stands for
    `sampler.fq_stream.read_ids` (default: the ids of this batch).

    Upstream walks the mu-sized mappings one by one; here the winning truncated record of every read is picked first,
    all decisions are gathered from the masks at once, and the outcome is assembled in upstream's order: reads in
    order of first appearance in `paf_trunc`, then (if unmapped reads are accepted) the unmapped ones that do have a
    full-length record, in `seqs` order."""
    full = parse_PAF(StringIO(paf_full))
    trunc = parse_PAF(StringIO(paf_trunc))
    rids = list(trunc)
    picks = [choose_best_mapper(trunc[rid])[0] for rid in rids]
    for rec in picks:
        rec.barcode = barcodes[rec.qname]                            # KeyError for a read without a barcode entry, as upstream
    accept = _lookup(contigs_filt, picks, barcodes, window)
    paf_dict = defaultdict(list)
    reads_decision = dict(seqs)                                      # upstream deep-copies; the values are immutable
    for rid, rec, yes in zip(rids, picks, accept):
        if yes:
            winner = choose_best_mapper(full[str(rec.qname)])[0]     # IndexError upstream if the read never mapped in full
            winner.barcode = barcodes[winner.qname]
        else:
            winner = rec
            reads_decision[rid] = reads_decision[rid][:mu]           # ejected after mu bases
        paf_dict[str(rec.qname)].append(winner)
    mapped = set(rids)
    n_accepted = int(accept.sum())
    n_rejected = len(rids) - n_accepted
    loose = [rid for rid in seqs if rid not in mapped]               # unmapped at mu bases: wholesale accept or reject
    if accept_unmapped:
        for rid in loose:
            if rid in full:
                paf_dict[rid].append(choose_best_mapper(full[rid])[0])
        n_accepted += len(loose)
    else:
        for rid in loose:
            reads_decision[rid] = seqs[rid][:mu]
        n_rejected += len(loose)
    ids = set(seqs) if all_read_ids is None else all_read_ids
    return paf_dict, reads_decision, len(mapped), len(ids - mapped), n_accepted, n_rejected


def filter_paf_dict(paf_dict: dict, mu: int = 400) -> dict:
    """`BossRunsSim.filter_paf_dict` (simulation.py:123-135): accepted reads are those whose record is not mu long."""
    return {rid: recs for rid, recs in paf_dict.items() if recs[0].qlen != mu}


def decide_text(contig_index: dict, contigs_filt: dict, seqs: dict[str, str], paf_full: str, paf_trunc: str,
                barcodes: dict[str, int], mu: int = 400, accept_unmapped: bool = False, window: int = 100):
    """`make_decisions` + `filter_paf_dict` + `convert_records` in one C pass over the two PAF texts (csrc/fastconv.c
    `decide_text`): no PafLine objects. Returns (batch, row_accepted, accepted_qlen, reads_decision, n_mapped, mapped,
    n_accepted, n_rejected) where `batch` is the PackedBatch of every record of `paf_dict` on a tracked contig,
    `row_accepted[i]` says whether row i belongs to an accepted read, and `accepted_qlen` lists the read lengths of all
    accepted reads (tracked contig or not) for the read-length distribution. None if the helper is not built."""
    fc = _fastconv()
    if fc is None or not hasattr(fc, "decide_text"):
        return None
    n = len(seqs) + 1
    contig, bc = np.empty(n, np.int32), np.empty(n, np.int32)
    tstart, tend, cl, sf, st = (np.empty(n, np.int64) for _ in range(5))
    rev, row_acc = np.empty(n, np.uint8), np.zeros(n, np.uint8)
    cp, sp = np.empty(n, np.uint64), np.empty(n, np.uint64)
    keep: list = []
    strat = {name: np.asarray(c.strat) for name, c in contigs_filt.items()}
    used, skipped, mapped, rejected, acc_qlen, n_acc, n_rej = fc.decide_text(
        paf_full, paf_trunc, seqs, contig_index, barcodes, strat, int(window), int(mu), bool(accept_unmapped), _best_index,
        (contig, tstart, tend, bc, rev, cp, cl, sp, sf, st), keep, row_acc)
    batch = PackedBatch(contig[:used], tstart[:used], tend[:used], bc[:used], rev[:used], cp[:used], cl[:used], sp[:used],
                        sf[:used], st[:used], keep, skipped)
    reads_decision = dict(seqs)
    for rid in rejected:
        reads_decision[rid] = reads_decision[rid][:mu]
    mapped_set = set(mapped)
    loose = [rid for rid in seqs if rid not in mapped_set]
    if accept_unmapped:
        n_acc += len(loose)
    else:
        for rid in loose:
            reads_decision[rid] = seqs[rid][:mu]
        n_rej += len(loose)
    return batch, row_acc[:used].astype(bool), acc_qlen, reads_decision, len(mapped), mapped_set, n_acc, n_rej


class BossRunsSim(BossRuns):
    """`BossRuns` driven by sampled reads with accept/reject decisions taken from the current strategy."""

    def __init__(self, *args, accept_unmapped: bool = False, mu: int = 400, text_decisions: bool = True, **kw):
        super().__init__(*args, **kw)
        self.accept_unmapped, self.mu = accept_unmapped, mu
        self.text_decisions = text_decisions           # decisions + batch in one C pass over the PAF texts (fastconv.decide_text)
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
        fast = decide_text(self.cc.contig_index, self.contigs_filt, read_seqs, paf_f, paf_t, read_barcodes, mu=self.mu,
                           accept_unmapped=self.accept_unmapped) if self.text_decisions else None
        if fast is not None:
            b, row_acc, acc_qlen, reads_decision, n_mapped, mapped, n_acc, n_rej = fast
            ids = set(read_seqs) if all_read_ids is None else all_read_ids
            logging.info(f"mapped {n_mapped}, not mapped {len(ids - mapped)}")
            logging.info(f"accepted {n_acc}, rejected {n_rej}")
            self.n_accepted, self.n_rejected = n_acc, n_rej
            self.rl_dist.update(read_lengths=dict(enumerate(acc_qlen)))
            self._prescore(b)
            self._effect_increments(increments=b)
            wins, strands = self.read_starts.count_read_starts_arrays(b.contig[row_acc], b.tstart[row_acc], b.tend[row_acc], b.rev[row_acc])
            self._read_starts_to_device(wins, strands)
            self.update_wrapper()
            self.batch += 1
            return reads_decision
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
