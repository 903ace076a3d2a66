"""ORACLE / TEST INFRASTRUCTURE — NumPy restatement of BOSS-AEONS' benefit and threshold step.

Follows `boss/aeons/sequences.py` of the reference (goldman-gp-ebi/BOSS-RUNS @ a5b6af8):

  Benefit.calc_fragment_benefit  :1555-1587   scores of one contig (one value per 100-bp node) -> (2, n) benefit, smu_sum
  Benefit._expand_scores         :1590-1605   ccl_max nodes of padding on both sides, 1 where the end is marked, last 0
  Benefit._calc_smu_moving       :1608-1620   trailing mu-window sums of the padded scores and of their mirror image
  Benefit._calc_benefit_moving   :1623-1641   ten nested windows weighted 1.0 .. 0.1
  Benefit.benefit_bins           :1644-1682   binary-exponent histogram of all non-zero benefits
  ContigPool.find_threshold      :1059-1094   cumulative benefit / time ratio, threshold one bin below the peak
  Sequence.find_strat_m0        :398-406     mask = benefit >= threshold, transposed to (n, 2)

PINNED: tests/golden/aeons.npz holds the outputs of those upstream functions themselves (oracle/make_golden_aeons.py
imports them with the shims of oracle/shims/); tests/test_aeons.py asserts this restatement reproduces them bit for bit.
The window sums go through `oracle.move_sum` like the rest of the oracle (Bottleneck is not installed; parity unpinned
at exactly that function, see oracle/move_sum.py).

Index quirks kept as upstream has them (Q-AEONS):
  A1  the right-hand padding is filled up to its last-but-one element only (`scoresx[-ccl_max:-1]`);
  A2  forward benefit of window w at padded position p = sum of the NEXT w nodes (p+1 .. p+w), for p < N - w - 1;
      reverse benefit = sum of the w nodes ENDING at p (p-w+1 .. p), for w <= p < N - 1;
  A3  smu row 1 is the mirror image's trailing sums and is NOT mirrored back before it is subtracted from the reverse
      benefit: position p is paired with the window starting at the mirrored position N-1-p;
  A4  the threshold's base time is alpha + rho + mu = 2 + 3 + 4 nodes (RUNS uses 3 + 3 + 4) and every node counts as one
      (no read-start distribution).
"""
from __future__ import annotations

import numpy as np

from oracle.move_sum import move_sum


def expand_scores(scores: np.ndarray, e1: bool, e2: bool, ccl_max: int) -> np.ndarray:
    sx = np.zeros(scores.shape[0] + 2 * ccl_max, dtype=np.float64)
    sx[ccl_max: -ccl_max] = scores
    sx[0: ccl_max] = 1 if e1 else 0
    sx[-ccl_max: -1] = 1 if e2 else 0          # A1
    return sx


def smu_moving(sx: np.ndarray, mu_ds: int) -> np.ndarray:
    return np.stack((move_sum(sx, window=mu_ds, min_count=1), move_sum(sx[::-1], window=mu_ds, min_count=1)))   # A3


def benefit_moving(sx: np.ndarray, ccl_ds: np.ndarray) -> np.ndarray:
    rev = sx[::-1]
    out = np.zeros((2, sx.shape[0]), dtype=np.float64)
    perc = np.arange(0.1, 1.1, 0.1)[::-1]
    assert perc.shape == ccl_ds.shape
    for i in range(ccl_ds.shape[0]):
        w = int(ccl_ds[i])
        fwd = move_sum(sx, window=w, min_count=1)[w: -1]
        bwd = move_sum(rev, window=w, min_count=1)[w: -1]
        out[0, 0: -w - 1] += fwd * perc[i]     # A2
        out[1, w: -1] += bwd[::-1] * perc[i]
    return out


def fragment_benefit(scores: np.ndarray, mu: int, approx_ccl: np.ndarray, e1: bool, e2: bool, node_size: int = 100):
    mu_ds = mu // node_size
    ccl_ds = approx_ccl // node_size
    c = int(ccl_ds[-1])
    sx = expand_scores(np.asarray(scores, dtype=np.float64), e1, e2, c)
    smu = smu_moving(sx, mu_ds)
    ben = benefit_moving(sx, ccl_ds)
    smu_sum = float(np.sum(smu))
    b = ben - smu
    b[b < 0] = 0
    b = b[:, c: -c]
    assert b.shape[1] == np.asarray(scores).shape[0]
    return b, smu_sum


def benefit_bins(benefit: np.ndarray):
    nz = benefit[np.nonzero(benefit)]
    norm = np.max(nz)
    _, ex = np.frexp(nz / norm)
    ex = np.abs(ex)
    counts_full = np.zeros(int(ex.max()) + 1, dtype="int")
    for part in np.array_split(ex, 12):
        c = np.bincount(part)
        counts_full[: c.shape[0]] += c
    uniq = np.nonzero(counts_full)[0]
    return np.power(2.0, -uniq) * norm, counts_full[uniq]


def find_threshold(benefits: list, smu_sums: list, mu: float, lam: float, node_size: int = 100) -> float:
    flat = np.column_stack(benefits).ravel()
    ubar0 = np.sum(smu_sums)
    alpha, rho = 200 // node_size, 300 // node_size
    tc = (lam - mu - 300) // node_size
    bins, counts = benefit_bins(flat)
    tbar0 = alpha + rho + (mu // node_size)       # A4
    cs_u = np.cumsum(bins * counts) + ubar0
    cs_t = np.cumsum(tc * counts) + tbar0
    k = int(np.argmax(cs_u / cs_t)) + 1
    return float(bins[k]) if k < bins.shape[0] else float(bins[-1])


def strategies(benefits: list, threshold: float) -> list:
    return [np.where(b >= threshold, True, False).transpose() for b in benefits]


def pool_update(scores: list, ends: list, mu: int, lam: float, approx_ccl: np.ndarray, node_size: int = 100):
    """`ContigPool` lines 1008-1010 on bare arrays: per-contig benefits, pool-wide threshold, per-contig masks."""
    ben, sums = [], []
    for s, (e1, e2) in zip(scores, ends):
        b, ss = fragment_benefit(s, mu, approx_ccl, bool(e1), bool(e2), node_size)
        ben.append(b)
        sums.append(ss)
    thr = find_threshold(ben, sums, mu, lam, node_size)
    return ben, sums, thr, strategies(ben, thr)
