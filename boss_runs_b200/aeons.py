"""BOSS-AEONS' benefit / threshold step on the GPU (second half).

Host-side mirror of `boss.aeons.sequences.Benefit` (sequences.py:1520-1682) plus the pool-wide step that
`ContigPool.find_threshold` (:1059-1094) and `Sequence.find_strat_m0` (:398-406) perform, all in one stateless C-ABI call
(`bossgpu_aeons_update`, kernels in csrc/aeons.cuh): AEONS has no reference — its contigs are the current assembly and
change with every batch, so nothing is kept on the device between calls.

    from boss_runs_b200.aeons import Benefit, pool_update
    b, smu_sum = Benefit.calc_fragment_benefit(scores, mu, approx_ccl, e1, e2)            # one contig, like upstream
    res = pool_update([c.scores for c in contigs], [(c.noi[0], c.noi[-1]) for c in contigs], mu=400, lam=rl.lam,
                      approx_ccl=rl.approx_ccl)                                           # benefits, threshold, masks

`Benefit.init_scoring_vec` / `score_array` (101-entry logistic table and a gather) stay NumPy one-liners, as upstream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import HIST_BINS, N_STEPS, as_c, check, ptr


@dataclass
class PoolResult:
    benefit: list          # per contig (2, n) float64: row 0 forward, row 1 reverse (Sequence.benefit)
    smu_sum: list          # per contig float (Sequence.smu_sum)
    threshold: float | None
    strat: list | None     # per contig (n, 2) bool (Sequence.find_strat_m0)
    counts: np.ndarray | None      # exponent histogram, int64 [HIST_BINS]
    normaliser: float | None
    ubar0: float | None
    n_nonzero: int = 0


def _run(scores, ends, mu: int, lam: float | None, approx_ccl, node_size: int, want_strategy: bool, device: int) -> PoolResult:
    lib = _lib.load()
    scores = [np.ascontiguousarray(s, dtype=np.float64) for s in scores]
    if len(scores) == 0 or len(scores) != len(ends):
        raise ValueError("one (e1, e2) pair per contig expected")
    off = np.zeros(len(scores) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in scores], out=off[1:])
    flat = as_c(np.concatenate(scores), np.float64)
    e1 = as_c([1 if e[0] else 0 for e in ends], np.uint8)
    e2 = as_c([1 if e[1] else 0 for e in ends], np.uint8)
    p = _lib.AeonsParams()
    p.mu_ds = int(mu // node_size)                                         # sequences.py:1575
    ccl_ds = np.asarray(approx_ccl) // node_size                           # :1576
    assert ccl_ds.shape == (N_STEPS,)
    perc = np.arange(0.1, 1.1, 0.1)[::-1]                                  # :1634, upstream's own expression
    for i in range(N_STEPS):
        p.ccl_ds[i], p.perc[i] = int(ccl_ds[i]), float(perc[i])
    p.want_strategy = int(want_strategy)
    if want_strategy:
        p.tc = float((lam - mu - 300) // node_size)                        # :1074
        p.tbar0 = float(200 // node_size + 300 // node_size + mu // node_size)     # :1072-1073,1079
    total = int(off[-1])
    ben = np.empty(2 * total, dtype=np.float64)
    ss = np.empty(len(scores), dtype=np.float64)
    strat = np.empty((total, 2), dtype=np.bool_) if want_strategy else None
    counts = np.zeros(HIST_BINS, dtype=np.int64) if want_strategy else None
    r = _lib.AeonsResult()
    check(lib.bossgpu_aeons_update(int(device), len(scores), ptr(off), ptr(flat), ptr(e1), ptr(e2), C.byref(p), ptr(ben), ptr(ss),
                                   ptr(strat), ptr(counts), C.byref(r)))
    b = [ben[2 * off[i]: 2 * off[i + 1]].reshape(2, -1) for i in range(len(scores))]
    if not want_strategy:
        return PoolResult(b, [float(x) for x in ss], None, None, None, None, None)
    return PoolResult(b, [float(x) for x in ss], float(r.threshold), [strat[off[i]: off[i + 1]] for i in range(len(scores))],
                      counts, float(r.normaliser), float(r.ubar0), int(r.n_nonzero))


def pool_update(scores, ends, mu: int, lam: float, approx_ccl, node_size: int = 100, device: int = 0) -> PoolResult:
    """Benefits of every contig, the pool-wide acceptance threshold and every contig's (n, 2) mask
    (`ContigPool` lines 1008-1010: `_contigs_benefits`, `find_threshold`, `_find_contig_strategies`)."""
    return _run(scores, ends, mu, lam, approx_ccl, node_size, True, device)


class Benefit:
    """Same static interface as upstream's `Benefit` (sequences.py:1520-1682)."""

    @staticmethod
    def init_scoring_vec(lowcov: float) -> np.ndarray:
        return 1 / (np.exp(np.arange(101) - lowcov) + 1)                   # :1531-1535

    @staticmethod
    def score_array(score_vec: np.ndarray, cov_arr: np.ndarray, node_size: int = 100) -> np.ndarray:
        return score_vec[(cov_arr // node_size).astype("int")]            # :1549-1552

    @staticmethod
    def calc_fragment_benefit(scores, mu: int, approx_ccl, e1: bool, e2: bool, node_size: int = 100, device: int = 0):
        res = _run([scores], [(e1, e2)], mu, None, approx_ccl, node_size, False, device)
        return res.benefit[0], res.smu_sum[0]
