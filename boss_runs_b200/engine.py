"""`Engine`: thin object wrapper over one libbossgpu handle (one GPU shard of the genome).

All numerics run in the CUDA kernels of libbossgpu.so; this file only marshals NumPy arrays across the
C ABI (include/bossgpu.h). The reference-facing classes in `boss_runs_b200.runs` sit on top of it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import BIN, BUCKET, RSD_WINDOW, HIST_BINS, N_PATTERNS, N_STEPS, N_TIMERS, check, ptr, as_c
from .priors import Priors, Scoring


@dataclass
class SegmentSpec:
    contig: int        # index in contigs_filt order
    start: int
    length: int


@dataclass
class UpdateOutcome:
    switched_on: bool
    threshold: float
    strat_size: int
    normaliser: float
    ubar0: float
    fhat_sum: float
    n_nonzero: int
    n_dropout: int
    n_accept: tuple[int, int]
    mirror_bytes: int = 0


def staircase_mult() -> np.ndarray:
    """Weights of the ten read-length steps, computed with upstream's own expression (reference.py:253)."""
    return np.arange(0.05, 1, 0.1)[::-1].copy()


class Engine:
    """One shard: an ordered list of contig segments with their counters, switches and strategy on a GPU."""

    def __init__(self, contig_lengths, ref_codes, n_barcodes: int = 1, ploidy: int = 1, n_sites_total: int | None = None,
                 device: int = 0, stream: int | None = None, segments: list[SegmentSpec] | None = None,
                 halo_bins: int = 0):
        """
        :param contig_lengths: lengths of ALL non-rejected contigs, in contigs_filt order
        :param ref_codes: per segment (or per contig when `segments` is None) uint8 arrays of seq_int (0..3)
        :param n_sites_total: `Reference.n_sites` (adds 4 per reject ref); defaults to sum(contig_lengths)
        :param stream: raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); None = default stream
        """
        self.lib = _lib.load()
        self.contig_lengths = np.asarray(contig_lengths, dtype=np.int64)
        self.nb = int(n_barcodes)
        self.ploidy = int(ploidy)
        if segments is None:
            segments = [SegmentSpec(k, 0, int(L)) for k, L in enumerate(self.contig_lengths)]
        self.segments = segments
        if len(ref_codes) != len(segments):
            raise ValueError("one reference array per segment expected")
        for seg, rc in zip(segments, ref_codes):
            assert len(rc) == seg.length, "reference array length does not match its segment"
        self.n_sites_total = int(n_sites_total if n_sites_total is not None else self.contig_lengths.sum())
        self.n_windows_total = int(sum(int(L / RSD_WINDOW) for L in self.contig_lengths))
        # scoring constants: model of the configured ploidy; unobserved-site constant of the HAPLOID model (Q5)
        self.model = Priors(ploidy=self.ploidy)
        hap = Scoring(ploidy=1)
        self.score0_contig, self.ent0_contig = float(hap.score0[0]), float(hap.ent0[0])

        segs = (_lib.Segment * len(segments))()
        for i, s in enumerate(segments):
            segs[i].contig, segs[i].contig_len = s.contig, int(self.contig_lengths[s.contig])
            segs[i].start, segs[i].len = s.start, s.length
        ref_all = as_c(np.concatenate([np.asarray(r, dtype=np.uint8) for r in ref_codes]), np.uint8)
        phi = as_c(self.model.phi, np.float64)
        pri = as_c(self.model.priors, np.float64)
        ppow = as_c(self.model.phi_stored[:, :, :_lib.FREEZE], np.float64)
        cfg = _lib.Config()
        cfg.abi_version, cfg.device, cfg.stream = _lib.ABI_VERSION, int(device), stream
        cfg.n_segments, cfg.n_barcodes = len(segments), self.nb
        cfg.segments, cfg.ref_codes = segs, ptr(ref_all)
        cfg.n_contigs_total, cfg.halo_bins = len(self.contig_lengths), int(halo_bins)
        cfg.contig_len_all = ptr(self.contig_lengths)
        cfg.n_sites_total, cfg.n_windows_total = self.n_sites_total, self.n_windows_total
        cfg.len_g = self.model.len_g
        cfg.phi, cfg.priors, cfg.phi_pow = ptr(phi), ptr(pri), ptr(ppow)
        cfg.score0_contig, cfg.entropy0_contig = self.score0_contig, self.ent0_contig
        h = C.c_void_p()
        check(self.lib.bossgpu_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.device = int(device)
        self.halo_bins = int(halo_bins)
        self._mult = staircase_mult()

    # -- lifetime --------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.bossgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self) -> None:
        check(self.lib.bossgpu_synchronize(self.h))

    # -- geometry --------------------------------------------------------------------------------
    def seg_len(self, seg: int) -> int:
        return self.segments[seg].length

    def seg_is_tail(self, seg: int) -> bool:
        s = self.segments[seg]
        return s.start + s.length == int(self.contig_lengths[s.contig])

    def seg_bins(self, seg: int) -> int:
        s = self.segments[seg]
        L = int(self.contig_lengths[s.contig])
        return (L // BIN + 1 - s.start // BIN) if self.seg_is_tail(seg) else s.length // BIN

    def seg_strat_rows(self, seg: int) -> int:
        return int(self.lib.bossgpu_strat_rows(self.h, seg))

    def seg_switches(self, seg: int) -> int:
        s = self.segments[seg]
        L = int(self.contig_lengths[s.contig])
        return (L // BUCKET - s.start // BUCKET + 1) if self.seg_is_tail(seg) else s.length // BUCKET

    # -- coverage update -------------------------------------------------------------------------
    def ingest_packed(self, seg, tstart, barcode, cig_off, cigar, base_off, bases, ascii_bases: bool = False,
                      contig_cov_add=None) -> None:
        seg = as_c(seg, np.int32); tstart = as_c(tstart, np.int64); barcode = as_c(barcode, np.int32)
        cig_off = as_c(cig_off, np.int64); cigar = as_c(cigar, np.uint32)
        base_off = as_c(base_off, np.int64); bases = as_c(bases, np.uint8)
        n = len(seg)
        assert len(tstart) == n and len(barcode) == n and len(cig_off) == n + 1 and len(base_off) == n + 1
        add = None if contig_cov_add is None else as_c(contig_cov_add, np.int64)
        check(self.lib.bossgpu_ingest_packed(self.h, n, ptr(seg), ptr(tstart), ptr(barcode), ptr(cig_off), ptr(cigar),
                                             ptr(base_off), ptr(bases), int(ascii_bases), 0, ptr(add)))

    def ingest_packed_device(self, n, seg, tstart, barcode, cig_off, cigar, base_off, bases, ascii_bases=False,
                             contig_cov_add=None) -> None:
        """Same with raw device pointers (ints); nothing is copied and the call does not synchronise."""
        check(self.lib.bossgpu_ingest_packed(self.h, int(n), seg, tstart, barcode, cig_off, cigar, base_off, bases,
                                             int(ascii_bases), 1, contig_cov_add))

    def ingest_records(self, contig, tstart, tend, barcode, rev, cig_off, cigar_text: bytes, seq_off, seq_text: bytes,
                       n_threads: int = 0) -> None:
        contig = as_c(contig, np.int32); tstart = as_c(tstart, np.int64); tend = as_c(tend, np.int64)
        barcode = as_c(barcode, np.int32); rev = as_c(rev, np.uint8)
        cig_off = as_c(cig_off, np.int64); seq_off = as_c(seq_off, np.int64)
        n = len(contig)
        assert len(cig_off) == n + 1 and len(seq_off) == n + 1
        assert cig_off[-1] == len(cigar_text) and seq_off[-1] == len(seq_text)
        check(self.lib.bossgpu_ingest_records(self.h, n, ptr(contig), ptr(tstart), ptr(tend), ptr(barcode), ptr(rev),
                                              ptr(cig_off), C.cast(C.c_char_p(cigar_text), C.c_void_p), ptr(seq_off),
                                              C.cast(C.c_char_p(seq_text), C.c_void_p), int(n_threads)))

    def ingest_records_ptr(self, contig, tstart, tend, barcode, rev, cigar_ptr, cigar_len, seq_ptr, seq_from, seq_to,
                           n_threads: int = 0) -> None:
        """Text ingest straight from the callers' string buffers (pointers as uint64)."""
        contig = as_c(contig, np.int32); tstart = as_c(tstart, np.int64); tend = as_c(tend, np.int64)
        barcode = as_c(barcode, np.int32); rev = as_c(rev, np.uint8)
        cigar_ptr = as_c(cigar_ptr, np.uint64); cigar_len = as_c(cigar_len, np.int64)
        seq_ptr = as_c(seq_ptr, np.uint64); seq_from = as_c(seq_from, np.int64); seq_to = as_c(seq_to, np.int64)
        check(self.lib.bossgpu_ingest_records_ptr(self.h, len(contig), ptr(contig), ptr(tstart), ptr(tend), ptr(barcode),
                                                  ptr(rev), ptr(cigar_ptr), ptr(cigar_len), ptr(seq_ptr), ptr(seq_from),
                                                  ptr(seq_to), int(n_threads)))

    def ingest_records_routed(self, contig, tstart, tend, barcode, rev, cigar_ptr, cigar_len, seq_ptr, seq_from, seq_to,
                              batch_cov_add, n_threads: int = 0) -> None:
        """Text ingest of the reads this shard sees; `batch_cov_add` = reference span of the whole batch per global contig."""
        contig = as_c(contig, np.int32); tstart = as_c(tstart, np.int64); tend = as_c(tend, np.int64)
        barcode = as_c(barcode, np.int32); rev = as_c(rev, np.uint8)
        cigar_ptr = as_c(cigar_ptr, np.uint64); cigar_len = as_c(cigar_len, np.int64)
        seq_ptr = as_c(seq_ptr, np.uint64); seq_from = as_c(seq_from, np.int64); seq_to = as_c(seq_to, np.int64)
        add = as_c(batch_cov_add, np.int64)
        assert add.shape == (len(self.contig_lengths),)
        check(self.lib.bossgpu_ingest_records_routed(self.h, len(contig), ptr(contig), ptr(tstart), ptr(tend), ptr(barcode),
                                                     ptr(rev), ptr(cigar_ptr), ptr(cigar_len), ptr(seq_ptr), ptr(seq_from),
                                                     ptr(seq_to), ptr(add), int(n_threads)))

    def prescore_begin(self) -> None:
        """A batch has arrived: score every tile now, on a second stream, while the host prepares the batch
        (`bossgpu_prescore_begin`); results do not depend on it."""
        check(self.lib.bossgpu_prescore_begin(self.h))

    def prescore(self, contig, tstart, tend) -> None:
        """Announce the batch's alignment intervals: the update will only re-score the tiles they cover."""
        contig = as_c(contig, np.int32); tstart = as_c(tstart, np.int64); tend = as_c(tend, np.int64)
        check(self.lib.bossgpu_prescore(self.h, len(contig), ptr(contig), ptr(tstart), ptr(tend)))

    def read_starts_add(self, window, strand) -> None:
        window = as_c(window, np.int64); strand = as_c(strand, np.uint8)
        assert len(window) == len(strand)
        check(self.lib.bossgpu_read_starts_add(self.h, len(window), ptr(window), ptr(strand)))

    def read_starts(self) -> np.ndarray:
        out = np.empty((self.n_windows_total, 2), dtype=np.int64)
        check(self.lib.bossgpu_get_read_starts(self.h, ptr(out), out.size))
        return out

    # -- strategy update -------------------------------------------------------------------------
    def _params(self, approx_ccl, time_cost, bucket_threshold, fhat_windows, debug, fhat_scalars=None) -> tuple:
        p = _lib.UpdateParams()
        w = np.asarray(approx_ccl) // BIN                       # reference.py:252
        assert w.shape == (N_STEPS,)
        for i in range(N_STEPS):
            p.w[i] = int(w[i])
            p.mult[i] = float(self._mult[i])
        p.tc = float(time_cost // BIN)                          # sequences.py:581
        p.bucket_threshold = float(bucket_threshold)
        keep = None
        if fhat_windows is not None:
            keep = as_c(fhat_windows, np.float64)
            assert keep.shape == (self.n_windows_total, 2), "fhat_windows must be [sum int(L/2000)][2]"
            p.fhat_windows = keep.ctypes.data
        p.write_debug = int(bool(debug))
        if fhat_scalars is not None:
            assert fhat_windows is None
            p.fhat_from_counts = 1
            p.rs_alpha, p.rs_denom, p.rs_zero_value = (float(x) for x in fhat_scalars)
        return p, keep

    @staticmethod
    def _outcome(r) -> UpdateOutcome:
        return UpdateOutcome(bool(r.switched_on), r.threshold, r.strat_size, r.normaliser, r.ubar0, r.fhat_sum,
                             r.n_nonzero, r.n_dropout, (r.n_accept[0], r.n_accept[1]), r.mirror_bytes)

    def update(self, approx_ccl, time_cost, bucket_threshold, fhat_windows=None, debug: bool = False,
               fhat_scalars=None) -> UpdateOutcome:
        """`fhat_windows`: compact F-hat [n_windows_total][2] from the host, or `fhat_scalars` = (alpha, denom,
        zero_value) to derive it on the device from the counts added with `read_starts_add`; neither = reuse."""
        p, _keep = self._params(approx_ccl, time_cost, bucket_threshold, fhat_windows, debug, fhat_scalars)
        r = _lib.UpdateResult()
        check(self.lib.bossgpu_update(self.h, C.byref(p), C.byref(r)))
        return self._outcome(r)

    def update_phase(self, phase: int, params) -> UpdateOutcome | None:
        r = _lib.UpdateResult()
        check(self.lib.bossgpu_update_phase(self.h, phase, C.byref(params), C.byref(r)))
        return self._outcome(r) if phase == 4 else None

    def exchange_buffer(self, which: int) -> tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        check(self.lib.bossgpu_exchange_buffer(self.h, which, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def exchange_tensor(self, which: int):
        """The exchange buffer as a uint8 torch tensor ALIASING the library's device memory (no copy), for
        torch.distributed collectives. torch is plumbing here: it never computes on these."""
        import torch
        ptr_, nbytes = self.exchange_buffer(which)

        class _Dev:
            pass
        d = _Dev()
        d.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr_, False), "version": 2}
        d._owner = self
        return torch.as_tensor(d, device=torch.device("cuda", self.device))

    # -- peer-memory fabric (csrc/fabric.cuh) ------------------------------------------------------
    def fabric_info(self) -> tuple[int, int, bytes]:
        """(device pointer, bytes, 64-byte CUDA IPC handle) of this shard's exchange block."""
        p, n = C.c_void_p(), C.c_size_t()
        hbuf = (C.c_ubyte * 64)()
        check(self.lib.bossgpu_fabric_info(self.h, C.byref(p), C.byref(n), C.cast(hbuf, C.c_void_p)))
        return int(p.value), int(n.value), bytes(hbuf)

    def ipc_open(self, handle: bytes) -> int:
        """Map a peer process' exchange block on this engine's device; returns the device pointer."""
        assert len(handle) == 64
        hbuf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        check(self.lib.bossgpu_ipc_open(self.device, C.cast(hbuf, C.c_void_p), C.byref(p)))
        return int(p.value)

    def ipc_close(self, dev_ptr: int) -> None:
        check(self.lib.bossgpu_ipc_close(self.device, C.c_void_p(dev_ptr)))

    def fabric_attach(self, peer_ptrs, timeout_s: float = 0.0) -> None:
        pp = as_c(peer_ptrs, np.uint64)
        check(self.lib.bossgpu_fabric_attach(self.h, len(pp), ptr(pp), float(timeout_s)))

    def update_fused_begin(self, params) -> None:
        """Enqueue every kernel of one sharded update, exchanges included (no host synchronisation)."""
        check(self.lib.bossgpu_update_fused_begin(self.h, C.byref(params)))

    def update_fused_end(self) -> UpdateOutcome:
        r = _lib.UpdateResult()
        check(self.lib.bossgpu_update_fused_end(self.h, C.byref(r)))
        return self._outcome(r)

    def set_shards(self, n_shards: int, shard_index: int, row_start) -> None:
        row_start = as_c(row_start, np.int64)
        assert row_start.shape == (n_shards + 1,)
        check(self.lib.bossgpu_set_shards(self.h, n_shards, shard_index, ptr(row_start)))

    def halo_pack(self) -> None:
        check(self.lib.bossgpu_halo_pack(self.h))

    def halo_unpack(self) -> None:
        check(self.lib.bossgpu_halo_unpack(self.h))

    def params(self, approx_ccl, time_cost, bucket_threshold, debug: bool = False, fhat_windows=None, fhat_scalars=None):
        """UpdateParams for `update_phase` (keeps the F-hat array alive on the returned object)."""
        p, keep = self._params(approx_ccl, time_cost, bucket_threshold, fhat_windows, debug, fhat_scalars)
        p._keep = keep
        return p

    # -- results / state -------------------------------------------------------------------------
    def strat(self, seg: int, out: np.ndarray | None = None) -> np.ndarray:
        """Contig.strat of segment `seg`: bool [rows][2][nb] (reference.py:118)."""
        rows = self.seg_strat_rows(seg)
        if out is None:
            out = np.empty((rows, 2, self.nb), dtype=np.bool_)
        assert out.shape == (rows, 2, self.nb) and out.dtype == np.bool_ and out.flags.c_contiguous
        check(self.lib.bossgpu_get_strat(self.h, seg, ptr(out), out.size))
        return out

    def strat_all(self, out: np.ndarray | None = None) -> np.ndarray:
        rows = self.seg_strat_rows(-1)
        if out is None:
            out = np.empty((rows, 2, self.nb), dtype=np.bool_)
        check(self.lib.bossgpu_get_strat_all(self.h, ptr(out), out.size))
        return out

    def strat_host(self) -> np.ndarray:
        """Zero-copy view of the library's pinned host mirror of every mask: bool [all rows][2][nb]. The library
        rewrites it at the end of each update that derives a strategy. The view keeps this engine alive."""
        p, n = C.c_void_p(), C.c_int64()
        check(self.lib.bossgpu_strat_host(self.h, C.byref(p), C.byref(n)))
        raw = (C.c_uint8 * n.value).from_address(p.value)
        raw._owner = self                      # numpy's base chain -> this ctypes block -> the engine
        return np.frombuffer(raw, dtype=np.bool_).reshape(-1, 2, self.nb)

    def buckets_host(self) -> list[np.ndarray]:
        """Zero-copy views of the pinned image of every segment's bucket switches: bool [n_sw][nb] per segment."""
        p, n = C.c_void_p(), C.c_int64()
        check(self.lib.bossgpu_buckets_host(self.h, C.byref(p), C.byref(n)))
        raw = (C.c_uint8 * max(n.value, 1)).from_address(p.value)
        raw._owner = self
        flat = np.frombuffer(raw, dtype=np.bool_)[: n.value].reshape(-1, self.nb)
        self.buckets_flat = flat               # all segments back to back (the views below are slices of it)
        out, row = [], 0
        for i in range(len(self.segments)):
            k = self.seg_switches(i)
            out.append(flat[row: row + k])
            row += k
        return out

    def set_strat_mirror(self, buf: np.ndarray, registered: bool = False) -> None:
        """Make `buf` (uint8/bool, C-contiguous, exactly this shard's strategy bytes — typically a slice of a
        shared-memory array) the host mirror the distribution kernel maintains."""
        assert buf.flags.c_contiguous and buf.dtype.itemsize == 1
        check(self.lib.bossgpu_set_strat_mirror(self.h, buf.ctypes.data, buf.size, int(registered)))
        self._mirror = buf

    def host_register(self, buf: np.ndarray) -> None:
        """Page-lock and map a host array for this process' CUDA context (once per array; slices of it can then
        be handed to `set_strat_mirror(..., registered=True)` of several engines)."""
        check(self.lib.bossgpu_host_register(buf.ctypes.data, buf.size * buf.dtype.itemsize))

    def host_unregister(self, buf: np.ndarray) -> None:
        check(self.lib.bossgpu_host_unregister(buf.ctypes.data))

    def seg_accept(self) -> np.ndarray:
        out = np.empty((len(self.segments), 2), dtype=np.int64)
        check(self.lib.bossgpu_get_seg_accept(self.h, ptr(out), out.size))
        return out

    def strat_packed(self) -> np.ndarray:
        n = self.seg_strat_rows(-1) * 2 * self.nb
        out = np.empty((n + 7) // 8, dtype=np.uint8)
        check(self.lib.bossgpu_get_strat_packed(self.h, ptr(out), out.size))
        return out

    def coverage(self, seg: int) -> np.ndarray:
        out = np.empty((self.seg_len(seg), 5, self.nb), dtype=np.uint16)
        check(self.lib.bossgpu_get_coverage(self.h, seg, ptr(out), out.size))
        return out

    def set_coverage(self, seg: int, cov: np.ndarray) -> None:
        cov = as_c(cov, np.uint16)
        assert cov.shape == (self.seg_len(seg), 5, self.nb)
        check(self.lib.bossgpu_set_coverage(self.h, seg, ptr(cov), cov.size))

    def scores(self, seg: int, entropy: bool = False):
        s = np.empty((self.seg_len(seg), self.nb))
        e = np.empty_like(s) if entropy else None
        check(self.lib.bossgpu_get_scores(self.h, seg, ptr(s), ptr(e), s.size))
        return (s, e) if entropy else s

    def scores_ds(self, seg: int) -> np.ndarray:
        out = np.empty((self.seg_bins(seg), self.nb))
        check(self.lib.bossgpu_get_scores_ds(self.h, seg, ptr(out), out.size))
        return out

    def benefit(self, seg: int, debug: bool = False):
        shape = (self.seg_bins(seg), 2, self.nb)
        ad = np.empty(shape)
        if not debug:
            check(self.lib.bossgpu_get_benefit(self.h, seg, ptr(ad), None, None, ad.size))
            return ad
        smu, eb = np.empty(shape), np.empty(shape)
        check(self.lib.bossgpu_get_benefit(self.h, seg, ptr(ad), ptr(smu), ptr(eb), ad.size))
        return ad, smu, eb

    def buckets(self, seg: int) -> tuple[np.ndarray, np.ndarray]:
        sw = np.empty((self.seg_switches(seg), self.nb), dtype=np.bool_)
        on = np.empty(self.nb, dtype=np.bool_)
        check(self.lib.bossgpu_get_buckets(self.h, seg, ptr(sw), sw.size, ptr(on)))
        return sw, on

    def set_buckets(self, seg: int, sw: np.ndarray) -> None:
        sw = as_c(sw, np.bool_)
        assert sw.shape == (self.seg_switches(seg), self.nb)
        check(self.lib.bossgpu_set_buckets(self.h, seg, ptr(sw), sw.size))

    def hist(self) -> tuple[np.ndarray, np.ndarray]:
        counts = np.empty(HIST_BINS, dtype=np.int64)
        f_grid = np.empty(HIST_BINS)
        check(self.lib.bossgpu_get_hist(self.h, ptr(counts), ptr(f_grid)))
        return counts, f_grid

    def fhat_rows(self, row0: int, n_rows: int) -> np.ndarray:
        """F-hat [n_rows][2] of the merged, length-adjusted rows [row0, row0 + n_rows) as the last update's histogram
        read it (expanded, tail-fixed and normalised on the device)."""
        out = np.empty((int(n_rows), 2), dtype=np.float64)
        check(self.lib.bossgpu_get_fhat(self.h, int(row0), int(n_rows), ptr(out)))
        return out

    def score_table(self) -> tuple[np.ndarray, np.ndarray]:
        s = np.empty((N_PATTERNS, 4))
        e = np.empty((N_PATTERNS, 4))
        check(self.lib.bossgpu_get_score_table(self.h, ptr(s), ptr(e)))
        return s, e

    def pattern_rank(self, counts) -> int:
        c = as_c(counts, np.uint16)
        assert c.shape == (5,)
        return int(self.lib.bossgpu_pattern_rank(ptr(c)))

    def timing(self) -> dict[str, float]:
        ms = np.empty(N_TIMERS, dtype=np.float32)
        check(self.lib.bossgpu_timing(self.h, ptr(ms)))
        names = ("scatter", "score_bin", "buckets", "smooth", "hist", "threshold", "distribute", "update")
        return {k: float(v) for k, v in zip(names, ms)}

    def launch_count(self) -> int:
        return int(self.lib.bossgpu_launch_count(self.h))

    def ingest_bytes(self) -> int:
        """Host->device bytes of the last text ingest."""
        return int(self.lib.bossgpu_ingest_bytes(self.h))

    def synth_coverage(self, seed: int = 11, mean_depth: float = 8.0, p_ref: float = 0.90, p_del: float = 0.04,
                       frac_dropout: float = 0.02, frac_deep: float = 0.01) -> None:
        check(self.lib.bossgpu_synth_coverage(self.h, seed, mean_depth, p_ref, p_del, frac_dropout, frac_deep))
