"""Stated tolerances of the parity tests (BASELINE.json north_star: integer work bit-exact, fp64 1e-9).

* coverage counters, bucket switches, histogram counts away from bin edges: bit-exact.
* per-site scores and their 100-site bin sums: relative 1e-9 (observed ~1e-15: device log() vs NumPy log,
  and a different but fixed summation order inside a bin).
* smoothed arrays (S_mu, expected and additional benefit): relative 1e-9 PLUS an absolute floor of
  1e-12 * max|oracle array|. The floor is the oracle's, not ours: Bottleneck's move_sum carries a running
  accumulator along the whole contig (oracle/move_sum.py), so where the true window sum is ~0 (frozen
  sites score 2.2e-308) the oracle itself holds residues of order 1e-16 * (largest window sum seen).
* threshold: relative 1e-12 (it is a power of two times the maximum benefit).
* strategy masks: bit-exact except at entries whose benefit lies within MASK_REL of the threshold.
"""
SCORE_RTOL = 1e-9
SMOOTH_RTOL = 1e-9
SMOOTH_ATOL_FRAC = 1e-12
THRESHOLD_RTOL = 1e-12
MASK_REL = 1e-9
