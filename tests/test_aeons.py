"""BOSS-AEONS benefit / threshold step (second half; boss/aeons/sequences.py:398-406,1059-1094,1520-1682).

* CPU: the oracle restatement (oracle/aeons_oracle.py) reproduces, bit for bit, what upstream's own `Benefit`,
  `ContigPool.find_threshold` and `Sequence.find_strat_m0` produced on three seeded contig pools (tests/golden/aeons.npz,
  written by oracle/make_golden_aeons.py).
* GPU (`-m gpu`): `boss_runs_b200.aeons.pool_update` (one C-ABI call, kernels in csrc/aeons.cuh) against the oracle."""
import numpy as np
import pytest

import tolerances as tol
from golden_io import GOLDEN
from oracle import aeons_oracle as ao

POOLS = ("small", "mixed", "deep")


def load_pool(name):
    z = np.load(GOLDEN / "aeons.npz", allow_pickle=False)
    ends = z[f"{name}_ends"]
    scores = [z[f"{name}_{i}_scores"] for i in range(len(ends))]
    return z, scores, [(bool(a), bool(b)) for a, b in ends], z[f"{name}_ccl"], float(z[f"{name}_lam"])


@pytest.mark.parametrize("name", POOLS)
def test_oracle_equals_upstream(name):
    z, scores, ends, ccl, lam = load_pool(name)
    ben, sums, thr, strats = ao.pool_update(scores, ends, 400, lam, ccl)
    for i in range(len(scores)):
        assert np.array_equal(ben[i], z[f"{name}_{i}_benefit"]), f"{name}/{i}: benefit"
        assert sums[i] == float(z[f"{name}_{i}_smu_sum"])
        assert np.array_equal(np.packbits(strats[i].ravel()), z[f"{name}_{i}_strat"])
        assert strats[i].shape == (scores[i].shape[0], 2)
    bins, counts = ao.benefit_bins(np.column_stack(ben).ravel())
    assert np.array_equal(bins, z[f"{name}_bins"]) and np.array_equal(counts, z[f"{name}_counts"])
    assert thr == float(z[f"{name}_threshold"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", POOLS)
def test_gpu_pool_update_matches_oracle(name, lib):
    from boss_runs_b200 import aeons
    z, scores, ends, ccl, lam = load_pool(name)
    want_b, want_s, want_thr, want_m = ao.pool_update(scores, ends, 400, lam, ccl)
    got = aeons.pool_update(scores, ends, mu=400, lam=lam, approx_ccl=ccl)
    assert abs(got.threshold - want_thr) <= tol.THRESHOLD_RTOL * want_thr
    flat = np.column_stack(want_b).ravel()
    bins, counts = ao.benefit_bins(flat)
    # histogram: exact away from bin edges (same rule as the RUNS histogram, tests/helpers.py)
    norm = flat.max()
    assert abs(got.normaliser - norm) <= tol.NORM_RTOL * norm
    ex = np.round(-np.log2(bins / norm)).astype(int)
    want_c = np.zeros(got.counts.shape[0], dtype=np.int64)
    want_c[ex] = counts
    nz = flat[flat != 0]
    ratio = nz / norm
    m, e = np.frexp(ratio)
    width = tol.SMOOTH_RTOL * ratio + tol.SMOOTH_ATOL_FRAC
    near = (ratio - np.ldexp(0.5, e) <= width) | (np.ldexp(1.0, e) - ratio <= width)
    slack = np.zeros(got.counts.shape[0] + 1, dtype=np.int64)
    k = np.bincount(np.abs(e)[near], minlength=got.counts.shape[0])[: got.counts.shape[0]]
    slack[:-1] += k; slack[1:] += k; slack[:-2] += k[1:]
    H = tol.HIST_HEAD
    assert not (np.abs(got.counts[:H] - want_c[:H]) > slack[:H]).any()
    for i in range(len(scores)):
        atol = tol.SMOOTH_ATOL_FRAC * float(np.max(np.abs(want_b[i]))) if want_b[i].size else 0.0
        err = np.abs(got.benefit[i] - want_b[i])
        assert not (err > tol.SMOOTH_RTOL * np.abs(want_b[i]) + atol).any(), f"{name}/{i}: benefit, worst {err.max():.3e}"
        assert abs(got.smu_sum[i] - want_s[i]) <= 1e-9 * max(abs(want_s[i]), 1e-300)
        diff = got.strat[i] != want_m[i]
        if diff.any():
            nearthr = np.abs(want_b[i].T - want_thr) <= tol.MASK_REL * want_thr
            assert not (diff & ~nearthr).any(), f"{name}/{i}: {(diff & ~nearthr).sum()} mask bits differ away from the threshold"
        assert got.strat[i].shape == (scores[i].shape[0], 2) and got.strat[i].dtype == np.bool_


@pytest.mark.gpu
def test_gpu_single_fragment_and_errors(lib):
    from boss_runs_b200 import aeons
    rng = np.random.default_rng(3)
    ccl = np.array([1167, 2729, 3903, 4918, 5866, 6808, 7797, 8912, 10321, 12713])
    for n, e1, e2 in ((1, True, True), (3, False, True), (129, True, False), (5000, False, False)):
        sc = rng.random(n)
        want, ss = ao.fragment_benefit(sc, 400, ccl, e1, e2)
        got, gs = aeons.Benefit.calc_fragment_benefit(sc, 400, ccl, e1, e2)
        assert got.shape == (2, n)
        np.testing.assert_allclose(got, want, rtol=tol.SMOOTH_RTOL, atol=tol.SMOOTH_ATOL_FRAC * want.max())
        assert abs(gs - ss) <= 1e-9 * ss
    with pytest.raises(ValueError):                                   # np.max of an empty array upstream (all benefits zero)
        aeons.pool_update([np.zeros(50)], [(False, False)], mu=400, lam=6000.0, approx_ccl=ccl)
    with pytest.raises(ValueError):                                   # bn.move_sum rejects windows < 1
        aeons.pool_update([rng.random(50)], [(False, False)], mu=400, lam=6000.0, approx_ccl=ccl // 100)
