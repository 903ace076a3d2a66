"""Stated tolerances of the parity tests (BASELINE.json north_star: integer work bit-exact, fp64 1e-9).

* coverage counters, bucket switches, histogram counts away from bin edges: bit-exact.
* per-site scores and their 100-site bin sums: relative 1e-9 (observed ~1e-15: device log() vs NumPy log,
  and a different but fixed summation order inside a bin).
* smoothed arrays (S_mu, expected and additional benefit): relative 1e-9 PLUS an absolute floor of
  1e-12 * max|oracle array|. The floor is the oracle's, not ours: Bottleneck's move_sum carries a running
  accumulator along the whole contig (oracle/move_sum.py), so where the true window sum is ~0 (frozen
  sites score 2.2e-308) the oracle itself holds residues of order 1e-16 * (largest window sum seen).
* normaliser (largest benefit): relative 1e-12 (observed: equal to the last bit or two).
* exponent histogram: counts exact for exponents < HIST_HEAD, except entries of the oracle that lie within MASK_REL of
  a bin edge (they may fall either side); F-hat mass per bin (f_grid) and ubar0 relative 1e-9. Beyond HIST_HEAD
  (benefit < 2^-40 of the maximum) the entries are rounding residue of upstream's running box sums, which the CUDA
  path does not reproduce (it sums every window from the bins): every such GPU entry must be near-zero in the oracle
  too, nothing more is asserted (note on zeros).
* device-expanded F-hat: relative 1e-12.
* threshold: relative 1e-12 (it is a power of two times the maximum benefit).
* strategy masks: bit-exact except at entries whose benefit lies within MASK_REL of the threshold.
"""
SCORE_RTOL = 1e-9
SMOOTH_RTOL = 1e-9
SMOOTH_ATOL_FRAC = 1e-12
THRESHOLD_RTOL = 1e-12
MASK_REL = 1e-9
NORM_RTOL = 1e-12
HIST_HEAD = 40
HIST_F_RTOL = 1e-9
UBAR_RTOL = 1e-9
FHAT_RTOL = 1e-12
