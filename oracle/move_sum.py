"""ORACLE / TEST INFRASTRUCTURE — restatement of Bottleneck's `move_sum` (float64, 1-D).

PARITY UNPINNED at this function: Bottleneck (pin `~=1.3.7`, /root/reference/pyproject.toml:19)
is a third-party C extension that is neither vendored under /root/reference nor installed in the
build container, and no reference test pins a value that flows through it.
Call sites on the hot path: /root/reference/boss/runs/reference.py:233-234 (S_mu, window 4) and
:259-260 (ten staircase windows), always with `min_count=1` on a 1-D float64 view.

Published algorithm (Bottleneck 1.3.x `move_template.c`, `move_sum` for float dtypes):

    asum = 0; count = 0
    for i in [0, min_count-1):      accumulate a[i] if not NaN;            y[i] = NaN
    for i in [min_count-1, window): accumulate a[i] if not NaN;            y[i] = asum if count >= min_count else NaN
    for i in [window, n):           ai = a[i]; aold = a[i-window]
                                    both finite : asum += ai - aold
                                    only ai     : asum += ai ; count += 1
                                    only aold   : asum -= aold ; count -= 1
                                    y[i] = asum if count >= min_count else NaN

i.e. a *running accumulator*: rounding error is carried along the whole array, so outputs in
regions whose true window sum is ~0 hold residues of order eps * (largest running sum seen).
The product (CUDA) path sums each window directly; parity tests therefore compare with
rtol 1e-9 plus an absolute floor tied to that residue scale (tests/tolerances.py).

Window validation follows Bottleneck: 1 <= window <= n, else ValueError.
"""
from __future__ import annotations

import numpy as np


def _check(a: np.ndarray, window: int, min_count) -> tuple[np.ndarray, int, int]:
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("oracle move_sum restates the 1-D case only")
    n = a.shape[0]
    window = int(window)
    if window < 1 or window > n:
        raise ValueError(f"Moving window (={window}) must between 1 and {n}, inclusive")
    mc = window if min_count is None else int(min_count)
    if mc < 1 or mc > window:
        raise ValueError("min_count must be between 1 and window")
    return a, window, mc


def move_sum_loop(a, window: int, min_count=None) -> np.ndarray:
    """Scalar transcription of the recurrence above (NaN-aware). Slow; small inputs only."""
    a, window, mc = _check(a, window, min_count)
    n = a.shape[0]
    y = np.empty(n, dtype=np.float64)
    asum = 0.0
    count = 0
    for i in range(n):
        ai = a[i]
        if i < window:
            if ai == ai:
                asum += ai
                count += 1
        else:
            aold = a[i - window]
            if ai == ai:
                if aold == aold:
                    asum += ai - aold
                else:
                    asum += ai
                    count += 1
            elif aold == aold:
                asum -= aold
                count -= 1
        y[i] = asum if count >= mc else np.nan
    return y


def move_sum(a, window: int, min_count=None, axis: int = -1) -> np.ndarray:
    """Vectorised, bit-identical form of `move_sum_loop` for NaN-free input.

    The recurrence is one sequential accumulation over d = [a[0:w], a[w:] - a[:-w]]; NumPy's
    1-D `cumsum` accumulates left to right in float64, so `cumsum(d)` reproduces every
    intermediate rounding of the loop. Inputs containing NaN take the scalar path.
    """
    a, window, mc = _check(a, window, min_count)
    if np.isnan(a).any():
        return move_sum_loop(a, window, mc)
    d = a.copy()
    if window < a.shape[0]:
        d[window:] = a[window:] - a[:-window]
    y = np.cumsum(d)
    if mc > 1:
        y[: mc - 1] = np.nan
    return y
