"""Reference-facing host API of the B200 strategy update.

Mirrors the operator interface of the reference for this path — `Contig`, `Reference`,
`CoverageConverter`, `BossRuns` (boss/runs/reference.py, boss/runs/sequences.py:657-794,
boss/runs/core.py:20-224) — with the same method names, argument meaning and error behaviour, while
every array operation runs in libbossgpu's CUDA kernels. Mapping (minimap2/mappy), MinKNOW/readfish
I/O and the simulator's sampling stay on the host and are not part of this module.

Standalone use (no upstream checkout needed)::

    run = BossRuns(contigs={"chr1": "ACGT...", ...}, ploidy=1, bucket_threshold=5)
    run.rl_dist.update(read_lengths)                      # host, as upstream
    run.process_batch_runs(paf_dict, seqs)                # convert -> scatter -> read starts -> update
    run.contigs["chr1"].strat                             # bool (L//100, 2, nb), refreshed by the update

`boss_runs_b200.dropin` puts the same engine underneath upstream's own `BossRuns` / `BossRunsSim` classes (a mixin that
overrides `init`, `_effect_increments` and `update_wrapper`; unchanged TOML plus an optional `[gpu]` table).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from ._lib import BIN, BUCKET, str_ptr
from .engine import Engine, UpdateOutcome
from .hostmodel import ReadlengthDist, ReadStartDist, best_record, parse_PAF
from .priors import Scoring

_SEQ_LUT = np.zeros(256, dtype=np.uint8)
for _ch, _v in zip("ACGT", range(4)):
    _SEQ_LUT[ord(_ch)] = _v


def seq_to_int(seq: str) -> np.ndarray:
    """`Contig._seq2int`: ACGT -> 0..3, every other letter -> 0 (reference.py:46-68)."""
    return _SEQ_LUT[np.frombuffer(seq.upper().encode(), dtype=np.uint8)]


@dataclass
class PackedBatch:
    """What `CoverageConverter.convert_records` hands to `_effect_increments`: per-read scalars plus POINTERS
    to the CIGAR and read strings (no text is copied or joined on the Python side; the library tokenises the
    CIGARs and copies / reverse-complements the aligned slices in C++ threads, straight into pinned memory)."""
    contig: np.ndarray      # int32 index into contigs_filt order
    tstart: np.ndarray      # int64
    tend: np.ndarray        # int64
    barcode: np.ndarray     # int32
    rev: np.ndarray         # uint8
    cigar_ptr: np.ndarray   # uint64: address of the CIGAR characters
    cigar_len: np.ndarray   # int64
    seq_ptr: np.ndarray     # uint64: address of the read's characters
    seq_from: np.ndarray    # int64: aligned slice [from, to) of the read, original orientation
    seq_to: np.ndarray
    keep: list              # the str objects the pointers refer to
    n_skipped: int = 0      # reads whose target is not a tracked contig
    # routed batches (one process per GPU): the arrays above hold only the reads this process' range of the genome sees;
    # these list EVERY read of the batch on a tracked contig (depth totals and read starts need the whole batch)
    all_contig: np.ndarray | None = None
    all_tstart: np.ndarray | None = None
    all_tend: np.ndarray | None = None
    all_rev: np.ndarray | None = None

    def whole_batch(self):
        """(contig, tstart, tend, rev) of every tracked read of the batch, routed or not."""
        if self.all_contig is None:
            return self.contig, self.tstart, self.tend, self.rev
        return self.all_contig, self.all_tstart, self.all_tend, self.all_rev

    def __len__(self) -> int:
        return int(self.contig.shape[0])

    def cigar_bytes(self, i: int) -> bytes:
        import ctypes
        return ctypes.string_at(int(self.cigar_ptr[i]), int(self.cigar_len[i]))

    def slice_bytes(self, i: int) -> bytes:
        """The aligned slice of read i, original orientation."""
        import ctypes
        return ctypes.string_at(int(self.seq_ptr[i]) + int(self.seq_from[i]), int(self.seq_to[i] - self.seq_from[i]))

    def texts(self):
        """(cig_off, cigar_text, seq_off, seq_text): the joined-text form `bossgpu_ingest_records` takes."""
        n = len(self)
        cigs = [self.cigar_bytes(i) for i in range(n)]
        seqs = [self.slice_bytes(i) for i in range(n)]
        cig_off = np.zeros(n + 1, dtype=np.int64)
        seq_off = np.zeros(n + 1, dtype=np.int64)
        if n:
            np.cumsum([len(c) for c in cigs], out=cig_off[1:])
            np.cumsum([len(p) for p in seqs], out=seq_off[1:])
        return cig_off, b"".join(cigs), seq_off, b"".join(seqs)


class CoverageConverter:
    """Host half of the coverage update (sequences.py:657-739): picks each read's record and the aligned
    slice of the read. CIGAR expansion and counting happen on the GPU."""

    def __init__(self, contig_index: dict[str, int], qt: int = 0):
        if qt != 0:
            raise NotImplementedError("upstream fixes the quality threshold at 0 (sequences.py:659)")
        self.contig_index = contig_index
        self.qt = qt

    def convert_records(self, paf_dict, seqs: dict[str, str], quals: dict[str, str] | None = None,
                        barcodes: dict[str, int] | None = None, ranges: np.ndarray | None = None) -> PackedBatch:
        """Same walk as upstream's loop (sequences.py:694-738). Runs in the `_fastconv` C helper when it has been
        built (`__graft_entry__.build()` / `python -m boss_runs_b200.build`), else in `_convert_records_py`; both
        produce identical batches (tests/test_hostmodel.py).

        `ranges` (int64 [n tracked contigs][2], one process per GPU): only reads overlapping [ranges[k][0], ranges[k][1]) of
        their contig k are converted — the others are merely listed in the batch's `all_*` arrays."""
        fc = _fastconv()
        if fc is None:
            return self._convert_records_py(paf_dict, seqs, ranges)
        n = len(paf_dict)
        contig, bc = np.empty(n, np.int32), np.empty(n, np.int32)
        tstart, tend, cl, sf, st = (np.empty(n, np.int64) for _ in range(5))
        rev = np.empty(n, np.uint8)
        cp, sp = np.empty(n, np.uint64), np.empty(n, np.uint64)
        keep: list = []
        bufs = (contig, tstart, tend, bc, rev, cp, cl, sp, sf, st)
        if ranges is None:
            used, skipped = fc.convert(paf_dict, seqs, self.contig_index, best_record, bufs, keep)
            return PackedBatch(contig[:used], tstart[:used], tend[:used], bc[:used], rev[:used], cp[:used], cl[:used], sp[:used],
                               sf[:used], st[:used], keep, skipped)
        ranges = np.ascontiguousarray(ranges, dtype=np.int64)
        assert ranges.shape == (len(self.contig_index), 2)
        ac, ats, ate, arv = np.empty(n, np.int32), np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.uint8)
        used, skipped, m = fc.convert(paf_dict, seqs, self.contig_index, best_record, bufs, keep, (ranges, ac, ats, ate, arv))
        return PackedBatch(contig[:used], tstart[:used], tend[:used], bc[:used], rev[:used], cp[:used], cl[:used], sp[:used],
                           sf[:used], st[:used], keep, skipped, ac[:m], ats[:m], ate[:m], arv[:m])

    def convert_text(self, paf_raw: str, seqs: dict[str, str], min_len: int = 200,
                     barcodes: dict[str, int] | None = None) -> PackedBatch:
        """The batch straight from the mapper's PAF text: `Paf.parse_PAF(StringIO(paf_raw), min_len=mu/2)`
        (mapper.py:63-65, paf.py:653-672) + `convert_records` in one C pass over the text (csrc/fastconv.c
        `convert_text`); no PafLine objects are built. `barcodes`: optional {read id: barcode index}.
        Falls back to `parse_PAF` + `convert_records` when the helper has not been built."""
        fc = _fastconv()
        if fc is None or not hasattr(fc, "convert_text"):
            from io import StringIO
            pd = parse_PAF(StringIO(paf_raw), min_len=min_len)
            if barcodes:
                for rid, recs in pd.items():
                    for r in recs:
                        r.barcode = barcodes.get(rid)
            return self.convert_records(pd, seqs)
        n = len(seqs) + 1                          # every read that is used is a key of `seqs`
        contig, bc = np.empty(n, np.int32), np.empty(n, np.int32)
        tstart, tend, cl, sf, st = (np.empty(n, np.int64) for _ in range(5))
        rev = np.empty(n, np.uint8)
        cp, sp = np.empty(n, np.uint64), np.empty(n, np.uint64)
        keep: list = []
        used, skipped, _ = fc.convert_text(paf_raw, seqs, self.contig_index, int(min_len), barcodes, _best_index,
                                           (contig, tstart, tend, bc, rev, cp, cl, sp, sf, st), keep)
        return PackedBatch(contig[:used], tstart[:used], tend[:used], bc[:used], rev[:used], cp[:used], cl[:used], sp[:used],
                           sf[:used], st[:used], keep, skipped)

    def _convert_records_py(self, paf_dict, seqs: dict[str, str], ranges=None) -> PackedBatch:
        contig, tstart, tend, bc, rev, cp, cl, sp, sf, st, keep = [], [], [], [], [], [], [], [], [], [], []
        everyone = [] if ranges is not None else None
        skipped = 0
        index = self.contig_index
        for recs in paf_dict.values():
            rec = recs[0] if len(recs) == 1 else best_record(recs)
            k = index.get(rec.tname)
            if k is None:
                skipped += 1          # upstream collects these under a key nobody reads (core.py:83-86)
                continue
            if ranges is not None:
                everyone.append((k, rec.tstart, rec.tend, 1 if rec.rev else 0))
                t0, t1 = min(rec.tstart, rec.tend), max(rec.tstart, rec.tend)
                if t1 <= ranges[k][0] or t0 >= ranges[k][1]:
                    continue
            s = seqs[rec.qname]
            if rec.rev:
                # upstream slices the reverse complement of the WHOLE string with qlen-based coordinates
                # (sequences.py:707-711; Q12): rc[a:b] == revcomp(s[n-b:n-a]), with Python's slice clamping
                n = len(s)
                a, b = rec.qlen - rec.qend, rec.qlen - rec.qstart
                if a < 0 or b < 0 or rec.qstart < 0 or rec.qend < 0:
                    # qend beyond qlen: upstream's slice would wrap around from the end of the read; not a mapping
                    raise ValueError("negative query coordinates")
                lo, hi = max(n - min(b, n), 0), max(n - min(a, n), 0)
            else:
                if rec.qstart < 0 or rec.qend < 0:
                    raise ValueError("negative query coordinates")
                lo, hi = min(rec.qstart, len(s)), min(rec.qend, len(s))
            cig = rec.cigar
            assert cig is not None
            contig.append(k)
            tstart.append(rec.tstart)
            tend.append(rec.tend)
            bc.append(0 if rec.barcode is None else rec.barcode)
            rev.append(rec.rev)
            cp.append(str_ptr(cig))
            cl.append(len(cig))
            sp.append(str_ptr(s))
            sf.append(lo)
            st.append(max(hi, lo))
            keep.append(cig)
            keep.append(s)
        b = PackedBatch(np.asarray(contig, dtype=np.int32), np.asarray(tstart, dtype=np.int64),
                        np.asarray(tend, dtype=np.int64), np.asarray(bc, dtype=np.int32),
                        np.asarray(rev, dtype=np.uint8), np.asarray(cp, dtype=np.uint64), np.asarray(cl, dtype=np.int64),
                        np.asarray(sp, dtype=np.uint64), np.asarray(sf, dtype=np.int64), np.asarray(st, dtype=np.int64),
                        keep, skipped)
        if everyone is not None:
            cols = list(zip(*everyone)) if everyone else [[], [], [], []]
            b.all_contig, b.all_tstart = np.asarray(cols[0], dtype=np.int32), np.asarray(cols[1], dtype=np.int64)
            b.all_tend, b.all_rev = np.asarray(cols[2], dtype=np.int64), np.asarray(cols[3], dtype=np.uint8)
        return b


def _best_index(keys: list) -> int:
    """`Paf.choose_best_mapper` on bare (mapq, AS) pairs (paf.py:710-722): NumPy's own argsort decides."""
    arr = np.array(keys, dtype=[("q", int), ("dp", int)])
    return int(np.argsort(arr, order=["q", "dp"])[-1])


_FASTCONV = False


def _fastconv():
    """The compiled helper, or None when it has not been built (host glue only: the CUDA path has no fallback)."""
    global _FASTCONV
    if _FASTCONV is False:
        try:
            from . import _fastconv as mod
            _FASTCONV = mod
        except ImportError:
            _FASTCONV = None
    return _FASTCONV


class Contig:
    """View of one reference sequence. Small state (`strat`, switches) is mirrored on the host after every
    update; the per-site arrays are fetched from the GPU on access (tests / debugging)."""

    def __init__(self, name: str, seq: str, ploidy: int = 1, rej: bool = False, barcodes: list | None = None):
        self.name = name.strip().split(" ")[0]
        self.length = len(seq)
        self.rej = rej
        self.barcodes = barcodes
        self.nbarcodes = len(barcodes) if barcodes is not None else 1
        # `seq` may also be given already integerised (uint8 codes 0..3) — large synthetic references
        self.seq_int = seq_to_int(seq) if isinstance(seq, str) else np.asarray(seq, dtype=np.uint8)
        assert len(set(self.seq_int[:100]) | {0, 1, 2, 3}) == 4
        self.bucket_size = BUCKET
        self.bucket_switches = np.zeros(shape=(int(self.length // BUCKET) + 1, self.nbarcodes), dtype="bool")
        self.switched_on = np.zeros(shape=(self.nbarcodes), dtype="bool")
        self.scoring = Scoring(ploidy=ploidy)
        self.len_b = self.scoring.priors.len_b
        self.score0, self.ent0 = self.scoring.score0, self.scoring.ent0
        if self.rej:
            self.strat = np.zeros(dtype="bool", shape=1)
        else:
            self.strat = np.ones(dtype="bool", shape=(self.length // BIN, 2, self.nbarcodes))
        self._engine: Engine | None = None
        self._seg = -1

    def _bind(self, engine: Engine, seg: int) -> None:
        self._engine, self._seg = engine, seg

    def _need(self) -> Engine:
        if self._engine is None:
            raise RuntimeError("contig is not bound to a GPU engine (reject refs carry no state)")
        return self._engine

    # per-site / per-bin state lives on the GPU
    @property
    def coverage(self) -> np.ndarray:
        return self._need().coverage(self._seg)

    @property
    def scores(self) -> np.ndarray:
        return self._need().scores(self._seg)

    @property
    def entropy(self) -> np.ndarray:
        return self._need().scores(self._seg, entropy=True)[1]

    @property
    def scores_ds(self) -> np.ndarray:
        return self._need().scores_ds(self._seg)

    @property
    def additional_benefit(self) -> np.ndarray:
        return self._need().benefit(self._seg)

    @property
    def smu(self) -> np.ndarray:
        return self._need().benefit(self._seg, debug=True)[1]

    @property
    def expected_benefit(self) -> np.ndarray:
        return self._need().benefit(self._seg, debug=True)[2]


class Reference:
    """Contig table with upstream's filtering rules (reference.py:274-373): contigs under 1e5 are dropped,
    `reject_refs` become 4-bp placeholders whose strategy is a single False."""

    def __init__(self, ref: str | None = None, mmi: str | None = None, reject_refs: str | None = None,
                 barcodes: list | None = None, records=None):
        self.ref, self.mmi, self.barcodes = ref, mmi, barcodes
        if records is None:
            if ref is None or not Path(ref).is_file():
                raise FileNotFoundError("Reference file not found")
            if not any(r in {".fa", ".fasta"} for r in Path(ref).suffixes):
                raise ValueError("Reference needs to be either .fa or .fasta (optionally gzipped).")
            records = read_fasta(ref)
        self.reject_refs = set(reject_refs.split(",")) if reject_refs else set()
        logging.info("Reading reference file")
        self.contigs = self._load_contigs(records)
        self.n_sites = self._total_sites()

    def _load_contigs(self, records, min_len: int = int(1e5), ploidy: int = 1) -> dict[str, Contig]:
        contigs = {}
        for cname, cseq in records:
            if len(cseq) < min_len:
                continue
            if cname not in self.reject_refs:
                contigs[cname] = Contig(name=cname, seq=cseq, ploidy=ploidy, barcodes=self.barcodes)
            else:
                contigs[cname] = Contig(name=cname, seq="ACGT", ploidy=ploidy, rej=True)
        return contigs

    def _total_sites(self) -> int:
        return np.sum(list(self.contig_lengths().values()))

    def contig_lengths(self) -> dict[str, int]:
        return {c.name: c.length for c in self.contigs.values()}

    def get_strategy_dict(self) -> dict[str, np.ndarray]:
        return {cname: cont.strat for cname, cont in self.contigs.items()}


def read_fasta(path: str):
    """Minimal FASTA reader (plain or gzip) yielding (name up to the first blank, sequence)."""
    import gzip
    opener = gzip.open if str(path).endswith(".gz") else open
    name, chunks = None, []
    with opener(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    yield name, "".join(chunks)
                head = line[1:].split()
                name, chunks = (head[0] if head else ""), []
            else:
                chunks.append(line.strip())
    if name is not None:
        yield name, "".join(chunks)


class BossRuns:
    """The strategy-update half of `boss.runs.core.BossRuns`, GPU-backed.

    Constructor arguments replace the TOML fields the path reads (`general.ref/barcodes`,
    `optional.ploidy/reject_refs/bucket_threshold`); `out_dir` enables the boss.npz hand-over to readfish.
    """

    def __init__(self, ref: str | None = None, contigs=None, ploidy: int = 1, barcodes: list[str] | None = None,
                 reject_refs: str | None = None, bucket_threshold: float = 5, out_dir: str | None = None,
                 device: int = 0, stream: int | None = None, strict_upstream_asserts: bool = True,
                 write_debug: bool = False, strategy_format: str = "npz"):
        if strategy_format not in ("npz", "bits", "both"):
            raise ValueError("strategy_format must be 'npz' (upstream's file, default), 'bits' or 'both'")
        self.strategy_format = strategy_format
        if not barcodes:
            self.barcodes_index = {"": 0}
        else:
            self.barcodes_index = {int(bc.split("barcode")[1]): i for i, bc in enumerate(barcodes)}
        self.barcodes = barcodes
        self.nbarcodes = len(self.barcodes_index)
        self.bucket_threshold = bucket_threshold
        self.out_dir = out_dir
        self.write_debug = write_debug
        records = None
        if contigs is not None:
            records = list(contigs.items()) if isinstance(contigs, dict) else list(contigs)
        self.ref = Reference(ref=ref, reject_refs=reject_refs, barcodes=barcodes, records=records)
        self.contigs = self.ref.contigs
        self.contigs_filt = {n: c for n, c in self.contigs.items() if not c.rej}
        if not self.contigs_filt:
            raise ValueError("no contig of at least 100 kb left to track")
        self.scoring = Scoring(ploidy=ploidy)      # validates ploidy like upstream (ValueError)
        self.rl_dist = ReadlengthDist()
        self.read_starts = ReadStartDist(contigs=self.contigs_filt, strict=strict_upstream_asserts)
        self._names = list(self.contigs_filt.keys())
        self.cc = CoverageConverter({n: i for i, n in enumerate(self._names)})
        self.ploidy = ploidy
        self._create_engine(device=device, stream=stream)
        self.batch = 0
        self._strat_views = None
        self._switch_views = None
        self.threshold: float | None = None
        self.last: UpdateOutcome | None = None
        if self.out_dir is not None:
            Path(self.out_dir, "masks").mkdir(parents=True, exist_ok=True)
            self._write_contig_strategies(self.ref.get_strategy_dict())

    def _create_engine(self, device: int, stream) -> None:
        """One GPU holds every contig whole. `sharding.ShardedRun` overrides this with one engine per shard."""
        self.engine = Engine(contig_lengths=[c.length for c in self.contigs_filt.values()],
                             ref_codes=[c.seq_int for c in self.contigs_filt.values()],
                             n_barcodes=self.nbarcodes, ploidy=self.ploidy, n_sites_total=int(self.ref.n_sites),
                             device=device, stream=stream)
        for i, c in enumerate(self.contigs_filt.values()):
            c._bind(self.engine, i)

    # -- output (core.py:59-69) ----------------------------------------------------------------------
    def _write_contig_strategies(self, contig_strats: dict[str, np.ndarray]) -> None:
        if self.out_dir is None:
            return
        if self.strategy_format in ("npz", "both"):
            tmp = f"{self.out_dir}/masks/boss_tmp.npz"
            np.savez(tmp, **contig_strats)
            Path(tmp).rename(f"{self.out_dir}/masks/boss.npz")
        if self.strategy_format in ("bits", "both"):
            self._write_packed_strategies()

    def _write_packed_strategies(self) -> None:
        """The same masks as `boss.bits` (stratfile.py): 1 bit per entry, packed on the GPU when one engine holds the
        genome, from the host mirror otherwise."""
        from . import stratfile
        tracked = [(n, c.length // BIN) for n, c in self.contigs_filt.items()]
        if getattr(self, "engines", None) is None and self._strat_views is not None:
            packed = self.engine.strat_packed()
        else:
            packed = stratfile.pack_strategies([c.strat for c in self.contigs_filt.values()])
        stratfile.write_bits(f"{self.out_dir}/masks/boss.bits", tracked, self.nbarcodes, packed,
                             rejected=[n for n, c in self.contigs.items() if c.rej])

    # -- coverage (core.py:77-86) ----------------------------------------------------------------------
    def _effect_increments(self, increments: PackedBatch) -> None:
        b = increments
        self.engine.ingest_records_ptr(b.contig, b.tstart, b.tend, b.barcode, b.rev, b.cigar_ptr, b.cigar_len,
                                       b.seq_ptr, b.seq_from, b.seq_to)

    # -- update (core.py:160-198) ----------------------------------------------------------------------
    def update_wrapper(self) -> None:
        import time as _t
        t0 = _t.perf_counter()
        # F-hat: three scalars from the host; the per-window values are formed on the GPU from its own counts
        scalars = self.read_starts.pointmass_scalars()
        t1 = _t.perf_counter()
        # `time_cost` does not exist before the first successful read-length update (Q14). Upstream only
        # reads it once some bucket is on (core.py:172,192), so the AttributeError is raised at that point.
        time_cost = getattr(self.rl_dist, "time_cost", None)
        try:
            out = self.engine.update(approx_ccl=self.rl_dist.approx_ccl,
                                     time_cost=np.float64("nan") if time_cost is None else time_cost,
                                     bucket_threshold=self.bucket_threshold, fhat_scalars=scalars, debug=self.write_debug)
        except AttributeError:
            # a bucket is on and there is no time_cost: the library stopped after the switches (no strategy was touched)
            self._pull_switches()
            raise
        t2 = _t.perf_counter()
        self.last = out
        self._pull_switches()
        t3 = _t.perf_counter()
        if out.switched_on:
            self.threshold = out.threshold
            self._pull_strategies()
            t4 = _t.perf_counter()
            self._write_contig_strategies(self.ref.get_strategy_dict())
        else:
            t4 = t3
        self.last_host_ms = {"fhat_host": (t1 - t0) * 1e3, "engine_update": (t2 - t1) * 1e3,
                             "pull_switches": (t3 - t2) * 1e3, "pull_strategies": (t4 - t3) * 1e3}

    def _pull_switches(self) -> None:
        """`Contig.bucket_switches` are views of the library's pinned image (refreshed by the update);
        `switched_on` follows reference.py:203-207: any switch of any barcode flags the whole contig. With thousands of
        contigs the per-contig test is one segmented reduction over the flat image, and contigs already flagged are done."""
        if self._switch_views is None:
            self._switch_views = self.engine.buckets_host()
            starts, row = [], 0
            for c, v in zip(self.contigs_filt.values(), self._switch_views):
                c.bucket_switches = v
                starts.append(row)
                row += v.shape[0]
            flat = getattr(self.engine, "buckets_flat", None)
            self._switch_flat = None
            if flat is not None and flat.shape[0] == row and all(v.shape[0] > 0 for v in self._switch_views):
                self._switch_flat = flat
                self._switch_starts = np.asarray(starts, dtype=np.intp)
            self._flagged = np.zeros(len(self._switch_views), dtype=bool)
        if self._flagged.all():
            return
        if self._switch_flat is not None:
            on = np.logical_or.reduceat(self._switch_flat.any(axis=1), self._switch_starts)
            todo = np.flatnonzero(on & ~self._flagged)
        else:
            todo = [i for i, v in enumerate(self._switch_views) if not self._flagged[i] and v.any()]
        if len(todo):
            contigs = list(self.contigs_filt.values())
            for i in todo:
                contigs[i].switched_on[...] = True
                self._flagged[i] = True

    def _pull_strategies(self) -> None:
        """`Contig.strat` of every contig is a view into the library's pinned host mirror, which the update has
        just refreshed with one device->host copy; nothing is copied here. Accept fractions for the log line
        (core.py:152-154) come from counters the distribution kernel filled."""
        if self._strat_views is None:
            flat = self.engine.strat_host()
            self._strat_views = []
            row = 0
            for c in self.contigs_filt.values():
                n = c.length // BIN
                self._strat_views.append(flat[row: row + n])
                row += n
        for c, v in zip(self.contigs_filt.values(), self._strat_views):
            if c.strat is not v:
                c.strat = v
        if logging.getLogger().isEnabledFor(logging.INFO):
            acc = self.engine.seg_accept()
            for i, c in enumerate(self.contigs_filt.values()):
                rows = max(c.strat.shape[0], 1)
                logging.info(f"{c.name}: {acc[i, 0] / rows}, {acc[i, 1] / rows}")

    def count_read_starts(self, paf_dict) -> None:
        """`ReadStartDist.count_read_starts` on the host mirror AND on the GPU-side counter."""
        wins, strands = self.read_starts.count_read_starts(paf_dict=paf_dict)
        self._read_starts_to_device(wins, strands)

    # -- one batch (core.py:202-224, minus the mapping call) -------------------------------------------
    def process_batch_runs(self, paf_dict, seqs: dict[str, str], quals: dict[str, str] | None = None,
                           barcodes: dict[str, int] | None = None, paf_dict_starts=None) -> None:
        """`paf_dict` = mappings of the batch ({read id: [PafLine]}); `paf_dict_starts` (default: the same)
        is the subset that feeds the read-start distribution — the simulator passes accepted reads only
        (simulation.py:171)."""
        import time as _t
        t0 = _t.perf_counter()
        self._prescore_begin()
        increments = self._convert(paf_dict, seqs, quals, barcodes)
        t1 = _t.perf_counter()
        self._prescore(increments)
        events = None
        if paf_dict_starts is None:
            # the winning record of every read is already in the batch arrays: same windows, same drops as
            # ReadStartDist.count_read_starts(paf_dict) without walking the PafLine objects a second time. Worked out while
            # the GPU is still busy with the early pass, applied only once the ingest has accepted the batch (upstream
            # counts read starts after _effect_increments, core.py:219-222)
            events = self.read_starts.window_events_arrays(*increments.whole_batch())
        t2 = _t.perf_counter()
        self._effect_increments(increments=increments)
        t3 = _t.perf_counter()
        if events is not None:
            self.read_starts.add_events(*events)
            self._read_starts_to_device(*events)
        else:
            self.count_read_starts(paf_dict_starts)
        t4 = _t.perf_counter()
        self.update_wrapper()
        t5 = _t.perf_counter()
        self.last_batch_ms = {"convert_records": (t1 - t0) * 1e3, "announce": (t2 - t1) * 1e3, "ingest": (t3 - t2) * 1e3,
                              "count_read_starts": (t4 - t3) * 1e3, "update_wrapper": (t5 - t4) * 1e3}
        self.batch += 1

    def _convert(self, paf_dict, seqs, quals=None, barcodes=None) -> PackedBatch:
        """`self.cc.convert_records`; `sharding.ShardedRun` routes the batch here (each process converts only its reads)."""
        return self.cc.convert_records(paf_dict=paf_dict, seqs=seqs, quals=quals, barcodes=barcodes)

    def process_batch_text(self, paf_raw: str, seqs: dict[str, str], quals: dict[str, str] | None = None,
                           barcodes: dict[str, int] | None = None, min_len: int = 200) -> None:
        """One batch from the mapper's raw PAF text (`Mapper._mappy_batch`'s return value, mapper.py:63): what
        `process_batch_runs` does after `map_sequences`, without materialising `{read id: [PafLine]}` — the text is
        tokenised once in C, the winning record of every read goes to the GPU, and the read starts are counted from
        the same arrays. `min_len` = mu/2 as `Mapper.map_sequences` passes it (mapper.py:64)."""
        self._prescore_begin()
        b = self.cc.convert_text(paf_raw, seqs, min_len=min_len, barcodes=barcodes)
        self._prescore(b)
        self._effect_increments(increments=b)
        wins, strands = self.read_starts.count_read_starts_arrays(b.contig, b.tstart, b.tend, b.rev)
        self._read_starts_to_device(wins, strands)
        self.update_wrapper()
        self.batch += 1

    def _read_starts_to_device(self, wins, strands) -> None:
        self.engine.read_starts_add(wins, strands)

    use_prescore = True

    def _engines(self):
        return getattr(self, "engines", None) or [self.engine]

    def _prescore_begin(self) -> None:
        """Split score/bin pass, early half: the GPU scores every tile from the counters as they are while the host
        prepares the batch (`bossgpu_prescore_begin`). Only called when an update is certain to follow."""
        if self.use_prescore:
            for e in self._engines():
                if hasattr(e, "prescore_begin"):
                    e.prescore_begin()

    def _prescore(self, b: PackedBatch) -> None:
        """... and the announcement: the update re-scores only the tiles these intervals cover (`bossgpu_prescore`)."""
        if self.use_prescore:
            for e in self._engines():
                if hasattr(e, "prescore"):
                    e.prescore(b.contig, b.tstart, b.tend)

    # -- device-resident variants (bench.py `value` leg, pipelined callers) ----------------------------
    def local_sites(self) -> int:
        """Sites held by this process' GPU."""
        return int(sum(s.length for s in self.engine.segments))

    def pack_for_device(self, batch: PackedBatch, engine: Engine | None = None) -> dict:
        """Tokenise a batch on the host into libbossgpu's packed arrays (what `bossgpu_ingest_packed` takes),
        e.g. to upload it ahead of time and ingest with `ingest_device`. With an `engine` that holds only part of
        the genome, the reads overlapping its segments are kept and `cov_add` carries the batch's reference span
        per contig (every shard needs the contig-wide depth totals)."""
        import ctypes as C
        engine = engine or self.engine
        lib = engine.lib
        seg_of = {s.contig: (i, s.start, s.start + s.length) for i, s in enumerate(engine.segments)}
        t0s, t1s = np.minimum(batch.tstart, batch.tend), np.maximum(batch.tstart, batch.tend)
        cov_add = np.zeros(len(engine.contig_lengths), dtype=np.int64)
        np.add.at(cov_add, batch.contig, t1s - t0s)
        keep = [i for i in range(len(batch)) if batch.contig[i] in seg_of
                and t1s[i] > seg_of[batch.contig[i]][1] and t0s[i] < seg_of[batch.contig[i]][2]]
        n = len(keep)
        comp = np.arange(256, dtype=np.uint8)
        for a, b in zip(b"ATGC", b"TACG"):
            comp[a] = b
        cig_off = np.zeros(n + 1, dtype=np.int64)
        base_off = np.zeros(n + 1, dtype=np.int64)
        ops = np.empty(int(sum(int(batch.cigar_len[i]) // 2 + 1 for i in keep)) + 1, dtype=np.uint32)
        bases = np.empty(int(sum(batch.seq_to[i] - batch.seq_from[i] for i in keep)) + 1, dtype=np.uint8)
        r, q = C.c_int64(), C.c_int64()
        w = 0
        for j, i in enumerate(keep):
            text = batch.cigar_bytes(i)
            k = lib.bossgpu_tokenize_cigar(text, len(text), ops[w:].ctypes.data, len(ops) - w, C.byref(r), C.byref(q))
            if k < 0:
                raise ValueError("CIGAR tokenizer failed")
            w += k
            cig_off[j + 1] = w
            sl = np.frombuffer(batch.slice_bytes(i), dtype=np.uint8)
            if q.value != len(sl) or r.value != int(t1s[i] - t0s[i]):
                raise AssertionError(f"read {i}: CIGAR does not span the aligned slice / target interval")
            base_off[j + 1] = base_off[j] + len(sl)
            bases[base_off[j]: base_off[j + 1]] = comp[sl[::-1]] if batch.rev[i] else sl
        return dict(n=n, seg=np.array([seg_of[batch.contig[i]][0] for i in keep], dtype=np.int32),
                    tstart=t0s[keep].astype(np.int64), barcode=batch.barcode[keep].astype(np.int32), cig_off=cig_off,
                    cigar=ops[:max(w, 1)].copy(), base_off=base_off, bases=bases, cov_add=cov_add)

    def ingest_device(self, d: dict, engine: Engine | None = None) -> None:
        """`d`: the dict of `pack_for_device` with every array replaced by a CUDA tensor on this device."""
        (engine or self.engine).ingest_packed_device(
            d["n"], d["seg"].data_ptr(), d["tstart"].data_ptr(), d["barcode"].data_ptr(), d["cig_off"].data_ptr(),
            d["cigar"].data_ptr(), d["base_off"].data_ptr(), d["bases"].data_ptr(), ascii_bases=True,
            contig_cov_add=d["cov_add"].data_ptr())

    def device_update(self, approx_ccl, time_cost, bucket_threshold, fhat_windows=None) -> UpdateOutcome:
        """The strategy update without the device->host copy of the masks."""
        self.last = self.engine.update(approx_ccl=approx_ccl, time_cost=time_cost, bucket_threshold=bucket_threshold,
                                       fhat_windows=fhat_windows)
        return self.last
