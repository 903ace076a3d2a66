"""TEST INFRASTRUCTURE — a minimal stand-in for the surface of upstream's `boss.runs.core.BossRuns` that
`boss_runs_b200.dropin._GpuMixin` sits on (the GPU box has no upstream checkout, so the drop-in wiring is exercised there
against this). It restates WHAT upstream's classes expose and the order in which `process_batch_runs` calls them
(core.py:23-55, 59-69, 202-224; reference.py:20-118, 274-373); every array operation is left to the mixin under test."""
from __future__ import annotations

from pathlib import Path
from types import SimpleNamespace

import numpy as np

from boss_runs_b200 import hostmodel
from boss_runs_b200.priors import Scoring
from boss_runs_b200.runs import seq_to_int


def make_args(name="standin", ref=None, barcodes=None, reject_refs=None, ploidy=1, bucket_threshold=5):
    """The fields of upstream's BossConfig the path reads (config.py:24-69)."""
    return SimpleNamespace(general=SimpleNamespace(name=name, ref=ref, mmi=ref, barcodes=barcodes),
                           optional=SimpleNamespace(reject_refs=reject_refs, ploidy=ploidy, bucket_threshold=bucket_threshold))


class Contig:
    """Attributes of upstream's Contig after __init__ (reference.py:20-118), host arrays included."""

    def __init__(self, name, seq, ploidy=1, rej=False, barcodes=None):
        self.name, self.seq, self.length, self.rej = name, seq, len(seq), rej
        self.barcodes = barcodes
        self.nbarcodes = len(barcodes) if barcodes is not None else 1
        self.seq_int = seq_to_int(seq)
        self.coverage = np.zeros((self.length, 5, self.nbarcodes), dtype="uint16")
        self.change_mask = np.zeros((self.length, self.nbarcodes), dtype="bool")
        self.bucket_size = 20_000
        self.bucket_switches = np.zeros((self.length // 20_000 + 1, self.nbarcodes), dtype="bool")
        self.switched_on = np.zeros(self.nbarcodes, dtype="bool")
        self.scoring = Scoring(ploidy=ploidy)
        self.score0, self.ent0 = self.scoring.score0, self.scoring.ent0
        self.scores = np.full((self.length, self.nbarcodes), self.score0[0])
        self.entropy = np.full((self.length, self.nbarcodes), self.ent0[0])
        self.strat = np.zeros(1, dtype="bool") if rej else np.ones((self.length // 100, 2, self.nbarcodes), dtype="bool")


class Reference:
    def __init__(self, records, reject_refs, barcodes):
        rej = set(reject_refs.split(",")) if reject_refs else set()
        self.contigs = {}
        for name, seq in records:
            if len(seq) < 100_000:                       # reference.py:319,330
                continue
            self.contigs[name] = Contig(name, "ACGT", rej=True) if name in rej else Contig(name, seq, barcodes=barcodes)
        self.n_sites = int(np.sum([c.length for c in self.contigs.values()]))

    def get_strategy_dict(self):
        return {n: c.strat for n, c in self.contigs.items()}


class BossRuns:
    """Call order of upstream's BossRuns; `records` (name, sequence) stands in for the FASTA, `mapper` for minimap2."""

    def __init__(self, args, records, out_dir):
        self.args, self._records = args, records
        self.name = args.general.name
        self.out_dir = str(out_dir)
        Path(self.out_dir, "masks").mkdir(parents=True, exist_ok=True)
        self.batch = 0
        self.rl_dist = hostmodel.ReadlengthDist()

    def init(self):
        bcs = self.args.general.barcodes
        self.barcodes_index = {"": 0} if not bcs else {int(bc.split("barcode")[1]): i for i, bc in enumerate(bcs)}
        self.nbarcodes = len(self.barcodes_index)
        self.ref = Reference(self._records, self.args.optional.reject_refs, bcs)
        self.contigs = self.ref.contigs
        self.contigs_filt = {n: c for n, c in self.contigs.items() if not c.rej}
        self.mapper = SimpleNamespace(map_sequences=lambda sequences: {})
        self.cc = None                                   # upstream: CoverageConverter()
        self.tracker = SimpleNamespace(update=lambda n, paf_dict: None)
        self.read_starts = None                          # upstream: ReadStartDist(contigs=self.contigs_filt)
        self.scoring = Scoring(ploidy=self.args.optional.ploidy)
        self._write_contig_strategies(contig_strats=self.ref.get_strategy_dict())

    def _write_contig_strategies(self, contig_strats):
        tmp = f"{self.out_dir}/masks/boss_tmp.npz"
        np.savez(tmp, **contig_strats)
        Path(tmp).rename(f"{self.out_dir}/masks/boss.npz")

    def process_batch_runs(self, new_reads, new_quals):
        paf_dict = self.mapper.map_sequences(sequences=new_reads)
        increments = self.cc.convert_records(paf_dict=paf_dict, seqs=new_reads, quals=new_quals)
        self._effect_increments(increments=increments)
        self.tracker.update(n=len(new_reads), paf_dict=paf_dict)
        self.read_starts.count_read_starts(paf_dict=paf_dict)
        self.update_wrapper()

    def update_wrapper(self):
        raise AssertionError("the mixin must override update_wrapper")

    def _effect_increments(self, increments):
        raise AssertionError("the mixin must override _effect_increments")
