"""Host-side mirrors of the small reference classes that feed the strategy update.

Same names, arguments and error behaviour as upstream so callers (and our parity tests) read like the
reference's own; the heavy arrays they used to produce now live on the GPU.

  PafLine, parse_PAF, choose_best_mapper .... boss/paf.py:12-74, 640-672, 710-722
  ReadlengthDist ............................. boss/readlengthdist.py:7-97
  ReadStartDist .............................. boss/runs/readstartdist.py:11-152 (expansion moved to the GPU)
  adjust_length, window helpers .............. boss/utils.py:192-226
"""
from __future__ import annotations

import logging
from collections import defaultdict
from io import StringIO
from pathlib import Path

import numpy as np
from scipy.special import betaln

from ._lib import RSD_WINDOW

_PAF_FIELDS = ("qname", "qlen", "qstart", "qend", "strand", "tname", "tlen", "tstart", "tend",
               "num_matches", "alignment_block_length", "mapq")
_TAG_TYPES = {"i": int, "A": str, "f": float, "Z": str}


def _conv(s: str, typ):
    try:
        return typ(s)
    except ValueError:
        return s


class PafLine:
    """One PAF alignment: 12 core columns + AS/cg/s1/tp tags (boss/paf.py:18-74)."""

    __slots__ = (*_PAF_FIELDS, "line", "rev", "align_score", "cigar", "s1", "primary", "barcode", "c")

    def __init__(self, line: str, tags: bool = True):
        self.line = line
        cols = line.strip().split("\t")
        if len(cols) < 12:
            raise IndexError("list index out of range")          # upstream indexes all 12 core columns (paf.py:50-52)
        for name, raw in zip(_PAF_FIELDS, cols[:12]):
            setattr(self, name, _conv(raw, int))
        self.qname, self.tname = str(self.qname), str(self.tname)
        self.rev = 0 if self.strand == "+" else 1
        self.c = -1
        self.barcode = None
        if tags:
            parsed = {}
            for t in cols[12:]:
                key, typ, val = t.split(":")          # upstream splits on every ':' the same way
                parsed[key] = _conv(val, _TAG_TYPES[typ])
            self.align_score = int(parsed.get("AS", 0))
            self.cigar = parsed.get("cg", None)
            self.s1 = parsed.get("s1", 0)
            self.primary = 1 if parsed.get("tp", None) == "P" else 0


def parse_PAF(paf_file, min_len: int = 1) -> dict[str, list[PafLine]]:
    """`Paf.parse_PAF`: path or StringIO -> {qname: [primary records with block length >= min_len]}."""
    if isinstance(paf_file, str) and Path(paf_file).is_file():
        fh = open(paf_file, "r")
    elif isinstance(paf_file, StringIO):
        fh = paf_file
    else:
        print("need file path or StringIO")
        return dict()
    out = defaultdict(list)
    with fh:
        for rec in fh:
            p = PafLine(rec)
            if p.alignment_block_length < min_len or not p.primary:
                continue
            out[str(p.qname)].append(p)
    return out


def choose_best_mapper(records: list) -> list:
    """Last element of an argsort by (mapq, alignment score) (boss/paf.py:710-722)."""
    keys = np.array([(r.mapq, r.align_score) for r in records], dtype=[("q", int), ("dp", int)])
    return [records[np.argsort(keys, order=["q", "dp"])[-1]]]


def best_record(recs):
    return recs[0] if len(recs) == 1 else choose_best_mapper(recs)[0]


def adjust_length(original_size: int, expanded: np.ndarray) -> np.ndarray:
    """Pad with a copy of the tail / truncate to `original_size` rows (boss/utils.py:206-226)."""
    d = original_size - expanded.shape[0]
    if d > 0:
        out = np.append(expanded, expanded[-d:], axis=0)
    elif d < 0:
        out = expanded[:-abs(d)]
    else:
        out = expanded
    assert out.shape[0] == original_size
    return out


def _sequential_sum(x: np.ndarray) -> float:
    """Left-to-right fp64 sum. Upstream normalises its length distributions with Python's built-in `sum`
    (readlengthdist.py:29,65), which adds element by element; np.sum's pairwise order differs in the last bits."""
    return float(np.cumsum(x, dtype=np.float64)[-1])


class ReadlengthDist:
    """Read-length model of the strategy update (boss/readlengthdist.py:7-97): a histogram of the lengths of accepted
    reads -> mean length `lam`, `time_cost`, and `approx_ccl`, the ten lengths at which the survival function of the
    read length drops below 0.95, 0.85, ..., 0.05 (the staircase `Contig.calc_u` convolves with, reference.py:241-260).
    Before any read has been seen the lengths follow a Gaussian prior (lam 6000, sd 4000) cut at lam + 10 sd.

    Same attributes as upstream (`read_lengths`, `L`, `ccl`, `approx_ccl`, `lam`, `longest_read`, and `time_cost` only
    after the first successful update, Q14) and bit-identical values (tests/test_oracle_kats.py re-asserts upstream's
    own vectors); the computation is array-shaped: lengths are counted with one scatter-add, the staircase is a
    binary search on the (non-increasing) survival function instead of a scanning loop."""

    MAX_LEN = 1_000_000          # longer reads are counted as MAX_LEN - 1 (readlengthdist.py:47-48)
    CUTOFF = 1e-6                # survival below this is treated as 0 (readlengthdist.py:83)

    def __init__(self, mu: int = 400, sd: int = 4000, lam: int = 6000, eta: int = 11):
        self.mu, self.sd, self.lam, self.eta = mu, sd, lam, eta
        self.read_lengths = np.zeros(self.MAX_LEN, dtype=np.uint16)
        x = np.arange(int(lam + 10 * sd), dtype="int")
        density = np.exp(-((x - lam + 1) ** 2) / (2 * (sd ** 2))) / (sd * np.sqrt(2 * np.pi))
        self.L = density / _sequential_sum(density)
        self.approx_ccl = self.ccl_approx_constant()

    def update(self, read_lengths: dict) -> None:
        """`read_lengths`: {read id: length}. Reads of at most 2 mu bases are rejected (truncated) ones and do not count."""
        lens = np.fromiter(read_lengths.values(), dtype=np.int64, count=len(read_lengths))
        lens = np.minimum(lens[lens > 2 * self.mu], self.MAX_LEN - 1)
        np.add.at(self.read_lengths, lens, np.uint16(1))                   # uint16 counters wrap like upstream's
        seen = np.flatnonzero(self.read_lengths)
        if seen.size == 0:
            logging.info("Attempted update of read lengths before observing any reads")
            return
        n = self.read_lengths[seen].astype(np.int64)
        self.lam = np.float64(int(np.dot(seen, n))) / np.float64(int(n.sum()))   # exact integers, one rounding
        self.longest_read = seen[-1]
        hist = self.read_lengths[: self.longest_read + 1].astype(np.float64)
        self.L = hist / _sequential_sum(hist)
        self.approx_ccl = self.ccl_approx_constant()
        logging.info(f"rld: {self.approx_ccl}")
        self.time_cost = self.lam - 400 - 300                              # lambda - mu - rho (readlengthdist.py:68)

    def ccl_approx_constant(self) -> np.ndarray:
        """Survival function ccl[i] = P(length > i) (kept as `self.ccl`, cut below 1e-6 and closed with one 0) and its
        `eta - 1` step positions: the first index where it is <= 1 - (k + 0.5) / (eta - 1)."""
        survival = np.empty(len(self.L) + 1)
        survival[0] = 1
        survival[1:-1] = 1 - np.cumsum(self.L[1:])
        survival[-1] = 0                                                   # 1 - 1: every read ends somewhere
        survival[survival < self.CUTOFF] = 0
        self.ccl = np.append(np.trim_zeros(survival, trim="b"), 0.0)
        levels = 1 - (np.arange(self.eta - 1) + 0.5) / (self.eta - 1)
        # non-increasing array: (entries > level) = len - (entries <= level), counted on the ascending view
        above = len(self.ccl) - np.searchsorted(self.ccl[::-1], levels, side="right")
        return above.astype("int32")


class ReadStartDist:
    """Read-start counts per 2 kb window and strand, and the Bayesian point-mass estimate F-hat.

    `update_f_pointmass()` returns the COMPACT estimate (one row per window); expansion x20, the two tail
    fixes and the normalisation (readstartdist.py:121-152, core.py:184-185) run on the GPU inside the
    strategy update. `expand()` reproduces the upstream array on the host for tests and small genomes.
    """

    def __init__(self, contigs: dict, window_size: int = RSD_WINDOW, alpha: float = 1.0, p0: float = 0.1,
                 strict: bool = True):
        self.alpha, self.p0, self.window_size = alpha, p0, window_size
        self.read_starts = {name: np.zeros(shape=(int(c.length / window_size), 2)) for name, c in contigs.items()}
        self.total_len = np.sum([a.shape[0] for a in self.read_starts.values()])
        self.target_size = int(np.sum([c.length for c in contigs.values()]) // 100)
        self.on_target = 1
        lendiff = self.target_size - int(self.total_len) * (window_size // 100)
        if strict:
            # upstream asserts this inside _expand_fhat (readstartdist.py:131); it fails for references of
            # more than ~100-200 contigs. strict=False lifts the limit (the GPU expansion handles any tail).
            assert lendiff < self.window_size

    def merge(self) -> np.ndarray:
        return self._merged_rows()

    def window_events(self, paf_dict: dict) -> tuple[np.ndarray, np.ndarray]:
        """The batch's read starts as (global window index, strand) pairs — what `count_read_starts` adds, in
        the sparse form the GPU-side counter takes. Binning follows np.histogram over [0, 2000*n_windows]:
        the right edge is closed, everything outside is dropped (readstartdist.py:68-78)."""
        if not hasattr(self, "_win_off"):
            off, acc = {}, 0
            for name, arr in self.read_starts.items():
                off[name] = (acc, int(arr.shape[0]))
                acc += int(arr.shape[0])
            self._win_off = off
        wins, strands = [], []
        for recs in paf_dict.values():
            rec = best_record(recs)
            ent = self._win_off.get(rec.tname)
            if ent is None:
                continue
            pos = rec.tend if rec.rev else rec.tstart
            base, nw = ent
            if pos < 0 or pos > self.window_size * nw:
                continue
            w = min(pos // self.window_size, nw - 1)
            wins.append(base + w)
            strands.append(1 if rec.rev else 0)
        return np.asarray(wins, dtype=np.int64), np.asarray(strands, dtype=np.uint8)

    def window_events_arrays(self, contig, tstart, tend, rev) -> tuple[np.ndarray, np.ndarray]:
        """`window_events` for records already reduced to arrays (one winning record per read, `contig` = index in
        the tracked-contig order): same bins, same drops, vectorised."""
        if not hasattr(self, "_win_base"):
            sizes = np.array([int(a.shape[0]) for a in self.read_starts.values()], dtype=np.int64)
            self._win_nw = sizes
            self._win_base = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64) if len(sizes) else sizes
        contig = np.asarray(contig, dtype=np.int64)
        rev = np.asarray(rev).astype(bool)
        pos = np.where(rev, np.asarray(tend, dtype=np.int64), np.asarray(tstart, dtype=np.int64))
        nw = self._win_nw[contig]
        ok = (pos >= 0) & (pos <= self.window_size * nw) & (nw > 0)
        w = np.minimum(pos // self.window_size, nw - 1)
        return (self._win_base[contig] + w)[ok].astype(np.int64), rev[ok].astype(np.uint8)

    def add_events(self, wins, strands) -> None:
        """Add (global window, strand) events to the host mirror of the counts."""
        if len(wins):
            np.add.at(self._merged_rows(), (wins, strands), 1.0)
            self.csum = getattr(self, "csum", 0.0) + float(len(wins))

    def count_read_starts_arrays(self, contig, tstart, tend, rev) -> tuple[np.ndarray, np.ndarray]:
        """`count_read_starts` from arrays (the text path never builds PafLine objects)."""
        wins, strands = self.window_events_arrays(contig, tstart, tend, rev)
        self.add_events(wins, strands)
        return wins, strands

    def pointmass_scalars(self, csum: float | None = None) -> tuple[float, float, float]:
        """(alpha, denom, zero_value) of `update_f_pointmass` (readstartdist.py:96-111): F-hat is
        (alpha + C) / denom where C > 0 and zero_value elsewhere. `csum` = total count (default: from the host
        mirror of the counts)."""
        nw = int(self.total_len)
        if csum is None:
            csum = getattr(self, "csum", 0.0)       # every event adds exactly 1 to exactly one window
        csum = np.float64(csum)
        denom = 2 * nw * self.alpha + csum
        rhs = self.alpha / (2 * nw * self.alpha + csum)
        beta_num = np.exp(betaln(self.alpha, ((2 * nw - 1) * self.alpha + csum)))
        beta_denom = np.exp(betaln(self.alpha, ((2 * nw - 1) * self.alpha))) or 1e-20
        p0_bit = self.p0 / (self.p0 + (1 - self.p0))
        return float(self.alpha), float(denom), float((1 - p0_bit * (beta_num / beta_denom)) * rhs)

    def count_read_starts(self, paf_dict: dict) -> tuple[np.ndarray, np.ndarray]:
        """Adds the batch's read starts to the per-contig window counts (same bins as upstream's np.histogram
        calls) and returns them as (global window, strand) events for the GPU-side counter."""
        wins, strands = self.window_events(paf_dict)
        if len(wins):
            merged_view = self._merged_rows()
            np.add.at(merged_view, (wins, strands), 1.0)
            self.csum = getattr(self, "csum", 0.0) + float(len(wins))
        return wins, strands

    def _merged_rows(self) -> np.ndarray:
        """One (total windows, 2) array whose per-contig slices ARE `read_starts[name]` (so `merge()` is free)."""
        if not hasattr(self, "_merged"):
            self._merged = np.concatenate(list(self.read_starts.values())) if self.read_starts else np.zeros((0, 2))
            row = 0
            for name, arr in list(self.read_starts.items()):
                n = arr.shape[0]
                self.read_starts[name] = self._merged[row: row + n]
                row += n
        return self._merged

    def update_f_pointmass(self) -> np.ndarray:
        """F-hat per window and strand, compact (before expansion): (alpha + C) / denom where reads have started,
        the point-mass posterior elsewhere (readstartdist.py:86-115) — the same three scalars the GPU is handed."""
        counts = self.merge()
        alpha, denom, zero_value = self.pointmass_scalars(csum=np.sum(counts[counts != 0]))
        return np.where(counts != 0, np.divide(np.add(alpha, counts), denom), zero_value)

    def expand(self, fhat_windows: np.ndarray, downsample_window: int = 100) -> np.ndarray:
        f = np.repeat(fhat_windows, int(self.window_size // downsample_window), axis=0)
        d = self.target_size - f.shape[0]
        if d > 0:
            f = np.append(f, f[-d:], axis=0)
        elif d < 0:
            f = f[:-abs(d)]
        s = np.sum(f)
        if s != 0:
            f = np.multiply(f, self.on_target / s)
        return f
