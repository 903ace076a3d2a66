"""TEST INFRASTRUCTURE — a NumPy model of ONE shard with the interface of `boss_runs_b200.engine.Engine`.

It exists so that the host side of the multi-GPU path (`boss_runs_b200.sharding`: planning, read routing, the five
exchange steps, strategy gathering) can run on a CPU box under torch.distributed's gloo backend, world_size 2,
with the exchange buffers living in host memory. Scoring comes from the oracle (oracle/boss_oracle.py); the
phase structure, buffer layouts and the exact integer-limb sums restate what the CUDA kernels do
(boss_runs_b200/csrc/strategy.cuh), so that the result must equal the unsharded oracle's. Never imported by the
product.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from oracle import boss_oracle as bo
from boss_runs_b200._lib import (BIN, BUCKET, HIST_BINS, BUF_SWITCH, BUF_NORM, BUF_HIST, BUF_MASK, BUF_HALO_SEND, BUF_HALO_RECV,
                                 BUF_STRAT, BUF_COV_TOTAL, UpdateParams)
from boss_runs_b200.engine import UpdateOutcome, staircase_mult
from boss_runs_b200.priors import Scoring

_MODELS: dict[int, bo.ScoreModel] = {}
_COMP = str.maketrans("ATGC", "TACG")


def _model(ploidy):
    if ploidy not in _MODELS:
        m = bo.ScoreModel(ploidy)
        m.build_table()
        _MODELS[ploidy] = m
    return _MODELS[ploidy]


def _limbs(v: float) -> tuple[int, int]:
    """v ~= hi + lo * 2^-32 as the CUDA kernels split it (strategy.cuh:to_limbs): below 2^51 hi = rint(v) and a
    signed lo = rint((v - hi) * 2^32); above, hi = floor(v) and lo = floor of the scaled remainder."""
    if v < 2.0 ** 51:
        hi = int(np.rint(v))
        return hi, int(np.rint((v - float(hi)) * 4294967296.0))
    hi = int(v)
    return hi, int((v - float(hi)) * 4294967296.0)


def _from_limbs(hi: int, lo: int, shift: int) -> float:
    return math.ldexp(float(hi) + float(lo) * (1.0 / 4294967296.0), -shift)


class NumpyShardEngine:
    def __init__(self, contig_lengths, ref_codes, n_barcodes=1, ploidy=1, n_sites_total=None, device=0, stream=None,
                 segments=None, halo_bins=0):
        self.contig_lengths = np.asarray(contig_lengths, dtype=np.int64)
        self.nb, self.ploidy, self.segments, self.halo_bins = int(n_barcodes), int(ploidy), segments, int(halo_bins)
        self.n_sites_total = int(n_sites_total if n_sites_total is not None else self.contig_lengths.sum())
        self.windows = [int(L / 2000) for L in self.contig_lengths]
        self.n_windows_total = int(sum(self.windows))
        self.model = _model(self.ploidy)
        hap = Scoring(ploidy=1)
        self.score0 = float(hap.score0[0])
        self.ref = [np.asarray(r, dtype=np.uint8) for r in ref_codes]
        self.cov = [np.zeros((s.length, 5, self.nb), dtype=np.uint16) for s in segments]
        self.cov_total = np.zeros(len(self.contig_lengths), dtype=np.int64)
        self.rs_counts = np.zeros((self.n_windows_total, 2), dtype=np.int64)
        L = self.contig_lengths
        self.o_row = np.concatenate(([0], np.cumsum(L // BIN + 1)))
        self.o_srow = np.concatenate(([0], np.cumsum(L // BIN)))
        self.M_rows, self.target = int(self.o_row[-1]), self.n_sites_total // BIN
        self.Tf = int(L.sum()) // BIN
        self.R0 = int(self.o_row[segments[0].contig] + segments[0].start // BIN)
        self.D0 = int(self.o_srow[segments[0].contig] + segments[0].start // BIN)
        self.n_bins = [self.seg_bins(i) for i in range(len(segments))]
        self.n_srows = [self._srows(i) for i in range(len(segments))]
        self.sw = [np.zeros((self.seg_switches(i), self.nb), dtype=bool) for i in range(len(segments))]
        self.n_rows = int(sum(self.n_bins))
        self._strat = np.ones((int(sum(self.n_srows)), 2, self.nb), dtype=bool)
        lg = 0
        while (1 << lg) < self.nb:
            lg += 1
        self.shift = 60 - lg
        self._mult = staircase_mult()
        self.buf = {BUF_SWITCH: np.zeros(1, np.int32), BUF_NORM: np.zeros(1, np.int64),
                    BUF_HIST: np.zeros(3 * HIST_BINS + 4, np.int64), BUF_COV_TOTAL: self.cov_total,
                    BUF_HALO_SEND: np.zeros(2 * max(1, self.halo_bins) * self.nb), BUF_HALO_RECV: np.zeros(2 * max(1, self.halo_bins) * self.nb),
                    BUF_STRAT: self._strat.view(np.uint8).reshape(-1)}
        self.phase_done = -1

    # ---- geometry (as engine.Engine) ----------------------------------------------------------------
    def seg_len(self, i):
        return self.segments[i].length

    def seg_is_tail(self, i):
        s = self.segments[i]
        return s.start + s.length == int(self.contig_lengths[s.contig])

    def seg_bins(self, i):
        s = self.segments[i]
        L = int(self.contig_lengths[s.contig])
        return (L // BIN + 1 - s.start // BIN) if self.seg_is_tail(i) else s.length // BIN

    def _srows(self, i):
        s = self.segments[i]
        L = int(self.contig_lengths[s.contig])
        return (L // BIN - s.start // BIN) if self.seg_is_tail(i) else s.length // BIN

    def seg_strat_rows(self, i):
        return int(sum(self.n_srows)) if i < 0 else self.n_srows[i]

    def seg_switches(self, i):
        s = self.segments[i]
        L = int(self.contig_lengths[s.contig])
        return (L // BUCKET - s.start // BUCKET + 1) if self.seg_is_tail(i) else s.length // BUCKET

    def set_shards(self, n, idx, row_start):
        self.n_shards, self.shard_index = n, idx
        self.row_start = np.asarray(row_start, dtype=np.int64)
        assert self.row_start[idx] == self.R0 and self.row_start[idx + 1] == self.R0 + self.n_rows
        mx = int(np.max(np.diff(self.row_start)))
        self.stride = (-(-(mx * 2 * self.nb) // 8) + 15) // 16 * 16
        self.buf[BUF_MASK] = np.zeros(self.stride * n, np.uint8)

    def exchange_tensor(self, which):
        return torch.from_numpy(self.buf[which].view(np.uint8).reshape(-1))

    def params(self, approx_ccl, time_cost, bucket_threshold, debug=False, fhat_windows=None, fhat_scalars=None):
        p = UpdateParams()
        w = np.asarray(approx_ccl) // BIN
        for i in range(10):
            p.w[i], p.mult[i] = int(w[i]), float(self._mult[i])
        p.tc, p.bucket_threshold = float(time_cost // BIN), float(bucket_threshold)
        assert fhat_windows is None and fhat_scalars is not None
        p.fhat_from_counts = 1
        p.rs_alpha, p.rs_denom, p.rs_zero_value = (float(x) for x in fhat_scalars)
        return p

    # ---- ingest ---------------------------------------------------------------------------------------
    def ingest_records_ptr(self, contig, tstart, tend, barcode, rev, cigar_ptr, cigar_len, seq_ptr, seq_from, seq_to, n_threads=0):
        seg_of = {s.contig: i for i, s in enumerate(self.segments)}
        for r in range(len(contig)):
            t0, t1 = int(min(tstart[r], tend[r])), int(max(tstart[r], tend[r]))
            self.cov_total[contig[r]] += t1 - t0
            i = seg_of.get(int(contig[r]))
            if i is None:
                continue
            s = self.segments[i]
            if t1 <= s.start or t0 >= s.start + s.length:
                continue
            cig = ctypes.string_at(int(cigar_ptr[r]), int(cigar_len[r])).decode()
            sl = ctypes.string_at(int(seq_ptr[r]) + int(seq_from[r]), int(seq_to[r] - seq_from[r])).decode()
            if rev[r]:
                sl = sl.translate(_COMP)[::-1]
            q = bo.expand_cigar(cig, sl, 0, len(sl))
            assert len(q) == t1 - t0
            pos = np.arange(t0, t1) - s.start
            ok = (pos >= 0) & (pos < s.length)
            b = int(barcode[r]) if 0 <= int(barcode[r]) < self.nb else 0
            np.add.at(self.cov[i], (pos[ok], q[ok], b), 1)

    def ingest_records_routed(self, contig, tstart, tend, barcode, rev, cigar_ptr, cigar_len, seq_ptr, seq_from, seq_to, batch_cov_add,
                              n_threads=0):
        before = self.cov_total.copy()
        self.ingest_records_ptr(contig, tstart, tend, barcode, rev, cigar_ptr, cigar_len, seq_ptr, seq_from, seq_to)
        self.cov_total[:] = before + np.asarray(batch_cov_add, dtype=self.cov_total.dtype)

    def read_starts_add(self, wins, strands):
        np.add.at(self.rs_counts, (np.asarray(wins), np.asarray(strands)), 1)

    def halo_pack(self):
        hb, nb = self.halo_bins, self.nb
        send = self.buf[BUF_HALO_SEND]
        send[:] = 0
        if hb <= 0:
            return
        F, Lg = self.segments[0], self.segments[-1]
        for b in range(nb):
            if F.start > 0:
                n = min(hb, self.n_bins[0])
                send[b * hb: b * hb + n] = self.ds[0][:n, b]
            if not self.seg_is_tail(len(self.segments) - 1):
                n = min(hb, self.n_bins[-1])
                send[(nb + b) * hb + hb - n: (nb + b + 1) * hb] = self.ds[-1][self.n_bins[-1] - n:, b]

    def halo_unpack(self):
        pass            # the halos are read straight from the receive buffer in phase 1

    # ---- phases ---------------------------------------------------------------------------------------
    def update_phase(self, phase, p):
        assert phase == 0 or phase == self.phase_done + 1 or (phase == 4 and self.phase_done == 0)
        out = getattr(self, f"_phase{phase}")(p)
        self.phase_done = -1 if phase == 4 else phase
        return out

    def _phase0(self, p):
        self.ds, self.n_dropout = [], 0
        on = 0
        for i, s in enumerate(self.segments):
            L = int(self.contig_lengths[s.contig])
            cov = self.cov[i].astype(np.int64)
            depth = cov.sum(axis=1)                                    # (len, nb)
            seen = depth.sum(axis=1) > 0
            mean = float(self.cov_total[s.contig]) / float(L * self.nb)
            drop = np.zeros(s.length, dtype=bool)
            if mean > 5:
                drop = depth.min(axis=1) <= int(mean / 8)
                self.n_dropout += int(drop.sum())
            sc = np.empty((s.length, self.nb))
            for b in range(self.nb):
                d = depth[:, b]
                live = np.nonzero((d < 30) & seen)[0]
                col = np.full(s.length, self.score0)
                col[live] = self.model.lookup(cov[live, :, b], self.ref[i][live])[0]
                if self.nb == 1:
                    col[d == 0] = self.score0
                col[d >= 30] = bo.TINY
                col[drop] = 0.0
                sc[:, b] = col
            ds = np.zeros((self.n_bins[i], self.nb))
            np.add.at(ds, np.arange(s.length) // BIN, sc)
            self.ds.append(ds)
            nfull = self.seg_switches(i) - (1 if self.seg_is_tail(i) else 0)
            bsum = depth[: nfull * BUCKET].reshape(nfull, BUCKET, self.nb).sum(axis=1) / BUCKET
            if self.seg_is_tail(i):
                bsum = np.concatenate((bsum, bsum[-1:]))
            self.sw[i] |= bsum >= p.bucket_threshold
            on |= int(self.sw[i].any())
        self.buf[BUF_SWITCH][0] = on
        self.buf[BUF_NORM][0] = 0

    def _phase1(self, p):
        hb, nb = self.halo_bins, self.nb
        recv = self.buf[BUF_HALO_RECV]
        w = [int(x) for x in p.w]
        self.benefit = []
        best = 0.0
        row = self.R0
        for i, s in enumerate(self.segments):
            n = self.n_bins[i]
            pad = max(max(w), 4)
            ben = np.zeros((n, 2, nb))
            for b in range(nb):
                ext = np.zeros(n + 2 * pad)
                ext[pad: pad + n] = self.ds[i][:, b]
                if i == 0 and s.start > 0 and hb > 0:
                    k = min(hb, pad)
                    ext[pad - k: pad] = recv[b * hb + hb - k: (b + 1) * hb]
                if i == len(self.segments) - 1 and not self.seg_is_tail(i) and hb > 0:
                    k = min(hb, pad)
                    ext[pad + n: pad + n + k] = recv[(nb + b) * hb: (nb + b) * hb + k]

                def box(width, fwd):
                    v = np.lib.stride_tricks.sliding_window_view(ext, width).sum(axis=1)
                    return v[pad: pad + n] if fwd else v[pad - width + 1: pad - width + 1 + n]
                eb_f = sum(box(w[k], True) * p.mult[k] for k in range(10))
                eb_r = sum(box(w[k], False) * p.mult[k] for k in range(10))
                ben[:, 0, b] = np.maximum(eb_f - box(4, True), 0.0)
                ben[:, 1, b] = np.maximum(eb_r - box(4, False), 0.0)
            self.benefit.append(ben)
            inside = (row + np.arange(n)) < self.target
            if inside.any():
                best = max(best, float(ben[inside].max()))
            row += n
        self.buf[BUF_NORM][0] = np.float64(best).view(np.int64)

    def _fhat_rows(self, p, rows):
        """normalised F-hat [len(rows)][2] for global (already adjust_length-ed) row indices"""
        c = self.rs_counts
        fw = np.where(c > 0, (p.rs_alpha + c) / p.rs_denom, p.rs_zero_value)
        W, e = self.n_windows_total, 20 * self.n_windows_total
        # exact sum of the expanded, tail-fixed array (multiplicity of every window)
        mult = np.zeros(W, dtype=np.int64)
        lim = min(self.Tf, e)
        full = np.clip(lim - 20 * np.arange(W), 0, 20)
        mult += full
        if self.Tf > e:
            d = self.Tf - e
            mult += np.clip(20 * np.arange(W) + 20 - np.maximum(e - d, 20 * np.arange(W)), 0, None)
        hi = lo = 0
        for wdw in range(W):
            for s_ in range(2):
                h, l_ = _limbs(math.ldexp(float(fw[wdw, s_]), 50))
                hi += h * int(mult[wdw])
                lo += l_ * int(mult[wdw])
        self.fhat_sum = _from_limbs(hi, lo, 50)
        scale = 1.0 / self.fhat_sum if self.fhat_sum != 0.0 else 1.0
        r = np.asarray(rows, dtype=np.int64).copy()
        r = np.where(r >= self.Tf, r - (self.target - self.Tf), r)
        r = np.where(r >= e, r - (self.Tf - e), r)
        return fw[np.clip(r // 20, 0, W - 1)] * scale          # rows beyond `target` are masked out by the caller

    def _phase2(self, p):
        norm = float(self.buf[BUF_NORM].view(np.float64)[0])
        hist = self.buf[BUF_HIST]
        hist[:] = 0
        if norm == 0.0:
            return
        _, norm_e = math.frexp(norm)
        extra = max(self.target - self.M_rows, 0)
        cnt = [0] * HIST_BINS
        hi = [0] * HIST_BINS
        lo = [0] * HIST_BINS
        uh = ul = nnz = 0
        ben = np.concatenate(self.benefit)                             # (n_rows, 2, nb)
        rows = self.R0 + np.arange(self.n_rows)
        passes = [(rows < self.target, rows)]
        if extra > 0:
            passes.append((rows >= self.M_rows - extra, rows + extra))
        for keep, rr in passes:
            f = self._fhat_rows(p, rr)
            for j in np.nonzero(keep)[0]:
                for s_ in range(2):
                    for b in range(self.nb):
                        x = float(ben[j, s_, b])
                        if x == 0.0:
                            continue
                        e = abs(math.frexp(x / norm)[1])
                        fv = float(f[j, s_])
                        h, l_ = _limbs(fv * math.ldexp(1.0, self.shift))
                        cnt[e] += 1
                        hi[e] += h
                        lo[e] += l_
                        h, l_ = _limbs(math.ldexp(fv * x, self.shift - norm_e))
                        uh += h
                        ul += l_
                        nnz += 1
        hist[:HIST_BINS] = cnt
        hist[HIST_BINS:2 * HIST_BINS] = hi                                # < 2^63: sum(fhat) ~ nb and shift = 60 - log2(nb)
        hist[2 * HIST_BINS:3 * HIST_BINS] = lo
        hist[3 * HIST_BINS: 3 * HIST_BINS + 3] = [uh, ul, nnz]

    def _phase3(self, p):
        hist = [int(v) for v in self.buf[BUF_HIST]]
        norm = float(self.buf[BUF_NORM].view(np.float64)[0])
        self.empty = norm == 0.0 or hist[3 * HIST_BINS + 2] == 0
        self.threshold, self.strat_size, self.ubar0 = 0.0, 0, 0.0
        if not self.empty:
            _, norm_e = math.frexp(norm)
            self.ubar0 = _from_limbs(hist[3 * HIST_BINS], hist[3 * HIST_BINS + 1], self.shift - norm_e)
            cs_u = cs_t = 0.0
            best, best_i, occ = 0.0, -1, []
            for e in range(HIST_BINS):
                if not hist[e]:
                    continue
                counts = float(hist[e])
                f_mean = _from_limbs(hist[HIST_BINS + e], hist[2 * HIST_BINS + e], self.shift) / counts
                cs_u += (math.ldexp(1.0, -e) * norm * f_mean) * counts
                cs_t += (p.tc * counts) * f_mean
                peak = (cs_u + self.ubar0) / (cs_t + 10.0)
                if best_i < 0 or peak > best:
                    best, best_i = peak, len(occ)
                occ.append(e)
            k = best_i + 1
            self.strat_size = k
            self.threshold = math.ldexp(1.0, -(occ[k] if k < len(occ) else occ[-1])) * norm
        ben = np.concatenate(self.benefit)
        rows = self.R0 + np.arange(self.n_rows)
        m = (ben >= self.threshold) & (rows < self.target)[:, None, None] if not self.empty else np.zeros(ben.shape, bool)
        bits = np.packbits(m.reshape(-1), bitorder="little")
        mask = self.buf[BUF_MASK]
        mask[self.shard_index * self.stride: self.shard_index * self.stride + len(bits)] = bits

    def _phase4(self, p):
        on = bool(self.buf[BUF_SWITCH][0])
        if on:
            assert self.phase_done == 3
            if self.empty:
                raise ValueError("all benefits are zero")
            mask = self.buf[BUF_MASK]
            dl0 = 0
            for i, s in enumerate(self.segments):
                n = self.n_srows[i]
                r = self.D0 + dl0 + np.arange(n)
                sh = np.searchsorted(self.row_start, r, side="right") - 1
                gate = np.repeat(self.sw[i], BUCKET // BIN, axis=0)[:n]
                for st in range(2):
                    for b in range(self.nb):
                        bit = ((r - self.row_start[sh]) * 2 + st) * self.nb + b
                        m = (mask[sh * self.stride + (bit >> 3)] >> (bit & 7)) & 1
                        cur = self._strat[dl0: dl0 + n, st, b]
                        cur[gate[:, b]] = m[gate[:, b]].astype(bool)
                dl0 += n
        if getattr(self, "_mirror", None) is not None:
            self._mirror[...] = self._strat
        acc = (int(self._strat[:, 0].sum()), int(self._strat[:, 1].sum()))
        return UpdateOutcome(on, self.threshold if on else 0.0, self.strat_size if on else 0,
                             float(self.buf[BUF_NORM].view(np.float64)[0]), getattr(self, "ubar0", 0.0), getattr(self, "fhat_sum", 0.0),
                             int(self.buf[BUF_HIST][3 * HIST_BINS + 2]), self.n_dropout, acc)

    # ---- results --------------------------------------------------------------------------------------
    def strat_host(self):
        return self._strat

    def buckets_host(self):
        return self.sw

    def host_register(self, buf):
        pass

    def host_unregister(self, buf):
        pass

    def set_strat_mirror(self, buf, registered=False):
        self._mirror = buf.view(np.bool_).reshape(-1, 2, self.nb)
        self._mirror[...] = self._strat

    def close(self):
        pass

    def coverage(self, i):
        return self.cov[i]

    def buckets(self, i):
        return self.sw[i], np.full(self.nb, bool(self.sw[i].any()))
