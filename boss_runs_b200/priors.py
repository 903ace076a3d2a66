"""Host-side constants of the scoring model: error model phi, genotype priors, score of an unobserved site.

Mirrors `boss.runs.sequences.Priors` and the constructor of `boss.runs.sequences.Scoring`
(/root/reference boss/runs/sequences.py:15-327, 335-342) with the same attribute names. These are a few
hundred floats computed once at start-up; they are handed to libbossgpu, which builds the dense score
table on the device (csrc/table.cuh). The float expressions are written exactly as upstream writes them
so the constants are bit-identical.
"""
from __future__ import annotations

import numpy as np

GENOTYPES_DIPLOID = ("AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "A-", "C-", "G-", "T-", "--")
ALPHABET = "ACGT-"


def _allele_counts() -> np.ndarray:
    """cnt[b, g]: copies of symbol b (ACGT-) in diploid genotype g."""
    return np.array([[g.count(s) for g in GENOTYPES_DIPLOID] for s in ALPHABET])


class Priors:
    """phi (len_b x len_g), priors (4 x len_g), phi_stored[i, j, k] = phi[i, j] ** k."""

    def __init__(self, ploidy: int = 1):
        if int(ploidy) == 1:
            self.diploid = False
        elif int(ploidy) == 2:
            self.diploid = True
        else:
            raise ValueError("Given ploidy is not defined")          # sequences.py:29
        self.len_b, self.len_g, self.phi = self._generate_phi(diploid=self.diploid)
        self.phi_stored = self.phi[:, :, None] ** np.arange(1000)[None, None, :]     # sequences.py:159-168
        self.priors = self._diploid_priors() if self.diploid else self._haploid_priors()
        self.prior_dist = np.array([self.priors[0]])

    @staticmethod
    def _generate_phi(diploid: bool = False, deletion_error: float = 0.03, err_missed_deletion: float = 0.1,
                      substitution_error: float = 0.04):
        """Deletion-aware error model (the only variant the hot path uses; sequences.py:70-91,112-153)."""
        if not deletion_error:
            raise NotImplementedError("boss_runs_b200 implements the default deletion-aware model only")
        n = 5
        if not diploid:
            phi = np.full((n, n), substitution_error / (n - 2))
            phi[np.arange(4), np.arange(4)] = 1.0 - (substitution_error + deletion_error)
            phi[4, :4] = deletion_error
            phi[:4, 4] = err_missed_deletion / (n - 1)
            phi[4, 4] = 1.0 - err_missed_deletion
            return n, n, phi
        cnt = _allele_counts()
        ok = 1.0 - (substitution_error + deletion_error)
        phi = np.zeros((n, 15))
        base, gap = cnt[:4], cnt[4]
        # plain genotypes
        phi[:4, :10] = np.select([base[:, :10] == 2, base[:, :10] == 1],
                                 [ok, ok / 2 + substitution_error / (2 * (n - 2))],
                                 substitution_error / (n - 2))
        # one allele deleted
        phi[:4, 10:14] = np.where(base[:, 10:14] == 1,
                                  ok / 2 + err_missed_deletion / (2 * (n - 1)),
                                  substitution_error / (2 * (n - 2)) + err_missed_deletion / (2 * (n - 1)))
        phi[:4, 14] = err_missed_deletion / (n - 1)
        phi[4] = np.select([gap == 2, gap == 1],
                           [1.0 - err_missed_deletion, (1.0 - err_missed_deletion) / 2 + deletion_error / 2],
                           deletion_error)
        return n, 15, phi

    @staticmethod
    def _haploid_priors(theta: float = 0.01, del_subs_ratio: float = 0.4) -> np.ndarray:
        pri = np.full((4, 5), theta / 3)                                   # sequences.py:223-236
        pri[np.arange(4), np.arange(4)] = 1.0 - (theta * (1.0 + del_subs_ratio))
        pri[:, 4] = theta * del_subs_ratio
        return pri

    @staticmethod
    def _diploid_priors(theta: float = 0.01, del_subs_ratio: float = 0.4) -> np.ndarray:
        popsize = 1000                                                     # sequences.py:257-264
        homo = 0.0
        hetero = 0.0
        aN = np.sum(1.0 / (np.arange(1, popsize + 1)))
        for i in range(popsize):
            homo += (1.0 / ((i + 1) * aN)) * ((i + 1) * float(i + 1) / (popsize ** 2))
            hetero += (1.0 / ((i + 1) * aN)) * 2 * ((popsize - (i + 1)) * float(i + 1) / (popsize ** 2))
        p_homo = homo / (homo + hetero)
        base = _allele_counts()[:4]
        pri = np.zeros((4, 15))
        pri[:, :10] = np.select([base[:, :10] == 2, base[:, :10] == 1],
                                [1 - theta * (1 + del_subs_ratio), ((1 - p_homo) * theta) / 3],
                                (p_homo * theta) / 3)
        pri[:, 10:14] = (1 - p_homo) * del_subs_ratio * theta
        pri[:, 14] = p_homo * del_subs_ratio * theta
        return pri

    def uniform_priors(self) -> None:
        self.priors.fill(1 / self.priors.shape[1])
        self.prior_dist = np.array([self.priors[0]])


def score_of_distribution(p: np.ndarray, phi: np.ndarray) -> tuple[float, float]:
    """(score, entropy) of ONE genotype distribution: H(p) - sum_i o_i H(p*phi_i/o_i)
    (sequences.py:520-549). Used on the host only for the constant of never-observed sites
    (sequences.py:342); every other score comes from the device-built table."""
    p = np.asarray(p, dtype=np.float64)[None, :]
    logs = np.zeros_like(p)
    np.log(p, where=p > 0.0, out=logs)
    entropy = np.sum(-p * logs, axis=1)
    new_entropy = np.zeros(1)
    for i in range(phi.shape[0]):
        q = p * phi[i]
        o = np.sum(q, axis=1)
        o[o == 0] = 1e-300
        q /= o[:, None]
        np.log(q, where=q > 0.0, out=logs)
        for j in range(p.shape[1]):
            new_entropy -= o * q[:, j] * logs[:, j]
    return float((entropy - new_entropy)[0]), float(entropy[0])


class Scoring:
    """Constants half of `boss.runs.sequences.Scoring`: priors, n_ref, score0, ent0 (as 1-element arrays,
    like upstream). The table half (`score_arr`) lives on the device; see `Engine.score_table()`."""

    def __init__(self, ploidy: int = 1):
        self.priors = Priors(ploidy=ploidy)
        self.n_ref = 4
        s0, e0 = score_of_distribution(self.priors.prior_dist[0], self.priors.phi)
        self.score0, self.ent0 = np.array([s0]), np.array([e0])
