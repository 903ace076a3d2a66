"""ORACLE / TEST INFRASTRUCTURE — CPU (NumPy) restatement of BOSS-RUNS' strategy-update path.

This module is the *checker*: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it. The product (`boss_runs_b200/`) never does,
and it raises if its CUDA library is missing rather than falling back to anything in here.

It restates, with the same NumPy operations in the same order (so that floating-point results are
bit-identical to the reference wherever the reference's own arithmetic is concerned), the functions
of SURVEY.md §8(a). All citations are into /root/reference/.

  a1/a2  records -> per-read increments ........ boss/runs/sequences.py:678-739, 744-794
  a3     coverage scatter ...................... boss/runs/reference.py:122-144
  a4     error model / genotype priors ......... boss/runs/sequences.py:39-155, 186-313
  a5/a6  posterior, mutual-information score ... boss/runs/sequences.py:485-516, 520-549
  a7/a8  score table + per-site update ......... boss/runs/sequences.py:347-393, 398-455
  a9     dropout masking ....................... boss/runs/reference.py:148-179
  a10    bucket switches ....................... boss/runs/reference.py:183-211, boss/utils.py:192-226
  a11/12 binning, S_mu, staircase benefit ...... boss/runs/reference.py:215-237, 241-269
  a13    bn.move_sum ........................... oracle/move_sum.py  (PARITY UNPINNED there)
  a14    read-start distribution F-hat ......... boss/runs/readstartdist.py:43-152
  a15    read-length staircase ................. boss/readlengthdist.py:36-97
  a16-18 merge, threshold, distribute .......... boss/runs/sequences.py:553-649, boss/runs/core.py:125-198

Pinning: `oracle/make_golden.py` runs the reference's own modules (imported from /root/reference
with the import shims in oracle/shims/) on committed inputs and stores their outputs under
tests/golden/; tests/test_oracle_golden.py asserts this restatement reproduces them (bit-exact
for every array, since only `move_sum` is foreign arithmetic and both sides use the same
restatement of it). The reference's own known-answer values are asserted in
tests/test_oracle_kats.py.

Deliberate restructurings that do not change results:
  * The score table is dense over the C(34,5)=278 256 count patterns with sum <= 29 instead of the
    reference's sparse 40^5 array with fill-on-miss: `update_scores` only ever looks up patterns with
    sum < 30 (sites with sum >= 30 are frozen first, sequences.py:419-430), and table entries and
    on-miss values come from the same function (sequences.py:439 vs :385).
  * Parity quirks Q1-Q15 of SURVEY.md §8 are reproduced, not fixed.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np
from scipy.special import betaln

from oracle.move_sum import move_sum

WINDOW = 100          # reference.py:109,215 ; sequences.py:577
BUCKET = 20_000       # reference.py:83
FREEZE = 30           # sequences.py:419
RSD_WINDOW = 2000     # readstartdist.py:13
TINY = np.finfo(float).tiny


# ------------------------------------------------------------------------------------------------
# a4  error model and genotype priors (sequences.py:39-155, 186-313), deletions enabled (defaults)
# ------------------------------------------------------------------------------------------------
DIPLOID_GENOTYPES = ["AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "A-", "C-", "G-", "T-", "--"]
BASES = "ACGT-"


def phi_matrix(ploidy: int, deletion_error=0.03, err_missed_deletion=0.1, substitution_error=0.04) -> np.ndarray:
    """phi[b, g] = P(observe b | genotype g); 5x5 haploid, 5x15 diploid (sequences.py:70-91, 112-153)."""
    if int(ploidy) not in (1, 2):
        raise ValueError("Given ploidy is not defined")          # sequences.py:29
    nb_ = 5
    if int(ploidy) == 1:
        phi = np.zeros((nb_, nb_))
        for b in range(nb_ - 1):
            for g in range(nb_ - 1):
                phi[b][g] = (1.0 - (substitution_error + deletion_error)) if b == g else substitution_error / (nb_ - 2)
                phi[nb_ - 1][g] = deletion_error
            phi[b][nb_ - 1] = err_missed_deletion / (nb_ - 1)
        phi[nb_ - 1][nb_ - 1] = 1.0 - err_missed_deletion
        return phi
    gts = DIPLOID_GENOTYPES
    ng = len(gts)
    phi = np.zeros((nb_, ng))
    ok = 1.0 - (substitution_error + deletion_error)
    for b in range(nb_ - 1):
        for g in range(ng - 5):
            k = gts[g].count(BASES[b])
            if k == 2:
                phi[b][g] = ok
            elif k == 1:
                phi[b][g] = ok / 2 + substitution_error / (2 * (nb_ - 2))
            else:
                phi[b][g] = substitution_error / (nb_ - 2)
        for g in range(10, 14):
            k = gts[g].count(BASES[b])
            if k == 1:
                phi[b][g] = ok / 2 + err_missed_deletion / (2 * (nb_ - 1))
            elif k == 0:
                phi[b][g] = substitution_error / (2 * (nb_ - 2)) + err_missed_deletion / (2 * (nb_ - 1))
        phi[b][ng - 1] = err_missed_deletion / (nb_ - 1)
    for g in range(ng):
        k = gts[g].count("-")
        if k == 2:
            phi[nb_ - 1][g] = 1.0 - err_missed_deletion
        elif k == 1:
            phi[nb_ - 1][g] = (1.0 - err_missed_deletion) / 2 + deletion_error / 2
        else:
            phi[nb_ - 1][g] = deletion_error
    return phi


def genotype_priors(ploidy: int, theta=0.01, del_subs_ratio=0.4) -> np.ndarray:
    """priors[ref, g]; 4x5 haploid (sequences.py:217-237), 4x15 diploid (:255-313)."""
    if int(ploidy) == 1:
        pri = np.zeros((4, 5))
        for i in range(4):
            for j in range(4):
                pri[i][j] = (1.0 - (theta * (1.0 + del_subs_ratio))) if i == j else theta / 3
        if del_subs_ratio > 0.0001:
            pri[:, -1] = theta * del_subs_ratio
        return pri
    popsize = 1000
    homo = 0.0
    hetero = 0.0
    aN = np.sum(1.0 / (np.arange(1, popsize + 1)))
    for i in range(popsize):
        homo += (1.0 / ((i + 1) * aN)) * ((i + 1) * float(i + 1) / (popsize ** 2))
        hetero += (1.0 / ((i + 1) * aN)) * 2 * ((popsize - (i + 1)) * float(i + 1) / (popsize ** 2))
    p_homo = homo / (homo + hetero)
    gts = DIPLOID_GENOTYPES
    pri = np.zeros((4, len(gts)))
    for b in range(4):
        for g in range(10):
            k = gts[g].count(BASES[b])
            if k == 2:
                pri[b][g] = 1 - theta * (1 + del_subs_ratio)
            elif k == 1:
                pri[b][g] = ((1 - p_homo) * theta) / 3
            else:
                pri[b][g] = (p_homo * theta) / 3
        for g in range(10, 14):
            pri[b][g] = (1 - p_homo) * del_subs_ratio * theta
        pri[b][len(gts) - 1] = p_homo * del_subs_ratio * theta
    return pri


def phi_powers(phi: np.ndarray, kmax: int = 1000) -> np.ndarray:
    """phi_stored[i, j, k] = phi[i, j] ** k (sequences.py:159-168)."""
    out = np.full((phi.shape[0], phi.shape[1], kmax), 1.0)
    for i in range(phi.shape[0]):
        for j in range(phi.shape[1]):
            out[i, j, :] = phi[i, j] ** np.arange(kmax)
    return out


# ------------------------------------------------------------------------------------------------
# a5/a6  posterior and score (sequences.py:485-516, 520-549)
# ------------------------------------------------------------------------------------------------
def posterior(patterns: np.ndarray, priors: np.ndarray, phi_pow: np.ndarray) -> np.ndarray:
    """post[h, n, j] for count patterns (n, 5); normalised over j with Z floored at 1e-300."""
    cov = np.array(patterns, dtype=np.int64, copy=True)
    cov[cov > 990] = 990                                              # :493
    len_b, len_g = phi_pow.shape[0], phi_pow.shape[1]
    n = cov.shape[0]
    post = np.repeat(priors[:, np.newaxis], repeats=n, axis=1)        # (4, n, len_g)  :499
    lik = np.full(n, 1.0)
    for j in range(len_g):
        if j > 0:
            lik.fill(1.0)
        for i in range(len_b):
            lik *= phi_pow[i, j, cov[:, i]]                           # :507
        for h in range(4):
            post[h, :, j] *= lik
    for h in range(4):
        z = np.sum(post[h, :, :], axis=1)
        z[z < 1e-300] = 1e-300
        post[h, :, :] /= z[:, np.newaxis]
    return post


def score_from_posterior(post_n: np.ndarray, phi: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(score, entropy) for posteriors (n, len_g): H(p) - sum_i o_i H(p*phi_i/o_i)  (:520-549).

    `np.log(..., where=x>0)` leaves the masked slots untouched; the reference relies on those slots
    being multiplied by exactly 0.0. We start `logs` from zeros so the masked slots are 0 rather than
    uninitialised memory — identical whenever the reference's result is well defined.
    """
    n, len_g = post_n.shape
    logs = np.zeros_like(post_n)
    np.log(post_n, where=post_n > 0.0, out=logs)
    entropy = np.sum(-post_n * logs, axis=1)
    new_entropy = np.zeros(n)
    obs = np.zeros(n)
    new_post = np.zeros((n, len_g))
    for i in range(phi.shape[0]):
        np.multiply(post_n, phi[i], out=new_post)
        np.sum(new_post, axis=1, out=obs)
        obs[obs == 0] = 1e-300
        new_post /= obs[:, np.newaxis]
        np.log(new_post, where=new_post > 0.0, out=logs)
        for j in range(len_g):
            new_entropy -= obs * new_post[:, j] * logs[:, j]
    return entropy - new_entropy, entropy


def pattern_scores(patterns: np.ndarray, priors: np.ndarray, phi: np.ndarray, phi_pow: np.ndarray):
    """(entropy[4, n], score[4, n]) for every reference base (sequences.py:460-481)."""
    post = posterior(patterns, priors, phi_pow)
    n = len(patterns)
    sc = np.zeros((4, n))
    en = np.zeros((4, n))
    for h in range(4):
        sc[h], en[h] = score_from_posterior(post[h, :, :], phi)
    return en, sc


# ---- dense pattern table over {c in N^5 : sum(c) <= 29} ------------------------------------------
SMAX = FREEZE - 1
N_PATTERNS = 278_256  # C(34, 5)


def pattern_rank(c: np.ndarray) -> np.ndarray:
    """Rank of count patterns (n,5) with sum<=29 in the combinatorial number system.

    q_k = (c_0+..+c_{k-1}) + (k-1) is strictly increasing in k and bounded by 33, so
    rank = C(q1,1)+C(q2,2)+C(q3,3)+C(q4,4)+C(q5,5) is a bijection onto [0, C(34,5)).
    """
    c = np.asarray(c, dtype=np.int64)
    p = np.cumsum(c, axis=1)
    q1, q2, q3, q4, q5 = p[:, 0], p[:, 1] + 1, p[:, 2] + 2, p[:, 3] + 3, p[:, 4] + 4
    return (q1 + q2 * (q2 - 1) // 2 + q3 * (q3 - 1) * (q3 - 2) // 6
            + q4 * (q4 - 1) * (q4 - 2) * (q4 - 3) // 24
            + q5 * (q5 - 1) * (q5 - 2) * (q5 - 3) * (q5 - 4) // 120)


def all_patterns() -> np.ndarray:
    """Every pattern with sum <= 29, ordered by rank."""
    out = np.empty((N_PATTERNS, 5), dtype=np.int64)
    k = 0
    # enumerate by brute force over prefix sums; cheap enough (278k rows) and obviously correct
    rng = np.arange(SMAX + 1)
    g = np.stack(np.meshgrid(rng, rng, rng, rng, rng, indexing="ij"), axis=-1).reshape(-1, 5)
    g = g[g.sum(axis=1) <= SMAX]
    r = pattern_rank(g)
    out[r] = g
    k = len(g)
    assert k == N_PATTERNS
    return out


@dataclass
class ScoreModel:
    """Constants + dense score table for one ploidy (restates `Scoring`, sequences.py:333-393)."""
    ploidy: int = 1
    phi: np.ndarray = field(init=False)
    priors: np.ndarray = field(init=False)
    phi_pow: np.ndarray = field(init=False)
    score0: float = field(init=False)
    ent0: float = field(init=False)
    score_table: np.ndarray | None = None      # (N_PATTERNS, 4)
    entropy_table: np.ndarray | None = None

    def __post_init__(self):
        self.phi = phi_matrix(self.ploidy)
        self.priors = genotype_priors(self.ploidy)
        self.phi_pow = phi_powers(self.phi)
        # sequences.py:342 — score of the *un-normalised* first prior row
        s0, e0 = score_from_posterior(np.array([self.priors[0]]), self.phi)
        self.score0, self.ent0 = float(s0[0]), float(e0[0])

    def build_table(self, chunk: int = 70_000) -> None:
        pats = all_patterns()
        self.score_table = np.empty((N_PATTERNS, 4))
        self.entropy_table = np.empty((N_PATTERNS, 4))
        for s in range(0, N_PATTERNS, chunk):
            en, sc = pattern_scores(pats[s:s + chunk], self.priors, self.phi, self.phi_pow)
            self.score_table[s:s + chunk] = sc.T
            self.entropy_table[s:s + chunk] = en.T

    def lookup(self, patterns: np.ndarray, ref_bases: np.ndarray):
        if self.score_table is None:
            self.build_table()
        r = pattern_rank(patterns)
        return self.score_table[r, ref_bases], self.entropy_table[r, ref_bases]


# ------------------------------------------------------------------------------------------------
# contig state (reference.py:18-119)
# ------------------------------------------------------------------------------------------------
def seq_to_int(seq: str) -> np.ndarray:
    """ACGT -> 0..3, every other ASCII letter -> 0 (reference.py:46-68)."""
    lut = np.zeros(256, dtype=np.uint8)
    for ch, v in zip("ACGT", range(4)):
        lut[ord(ch)] = v
    return lut[np.frombuffer(seq.upper().encode(), dtype=np.uint8)]


@dataclass
class ContigState:
    name: str
    seq_int: np.ndarray
    nb: int = 1
    rej: bool = False
    score0: float = 0.0       # Q5: always the HAPLOID score0 (reference.py:319,334)
    ent0: float = 0.0

    def __post_init__(self):
        L = self.length = int(self.seq_int.shape[0])
        self.coverage = np.zeros((L, 5, self.nb), dtype=np.uint16)
        self.change_mask = np.zeros((L, self.nb), dtype=bool)
        self.bucket_switches = np.zeros((L // BUCKET + 1, self.nb), dtype=bool)
        self.switched_on = np.zeros(self.nb, dtype=bool)
        self.scores = np.full((L, self.nb), self.score0)
        self.entropy = np.full((L, self.nb), self.ent0)
        self.strat = np.zeros(1, dtype=bool) if self.rej else np.ones((L // WINDOW, 2, self.nb), dtype=bool)


def make_contigs(records, reject_refs=(), barcodes=None, min_len=int(1e5)) -> dict[str, ContigState]:
    """`Reference._load_contigs` (reference.py:319-338): drop < 1e5, 4-bp placeholder for reject refs."""
    hap = ScoreModel(1)
    nb = len(barcodes) if barcodes else 1
    out = {}
    for name, seq in records:
        if len(seq) < min_len:
            continue
        key = name
        cname = name.strip().split(" ")[0]
        if name not in reject_refs:
            out[key] = ContigState(cname, seq_to_int(seq), nb=nb, score0=hap.score0, ent0=hap.ent0)
        else:
            out[key] = ContigState(cname, seq_to_int("ACGT"), nb=1, rej=True, score0=hap.score0, ent0=hap.ent0)
    return out


# ------------------------------------------------------------------------------------------------
# a1/a2  records -> increments (sequences.py:678-794)
# ------------------------------------------------------------------------------------------------
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=XB])")
_COMP = str.maketrans("ATGC", "TACG")
_OP_CODE = {"M": 6, "D": 7, "I": 8, "S": 9}


def _chars_to_int(s: str) -> np.ndarray:
    """ACGT -> 0..3; anything else -> ord(c) - 48 as uint8 (sequences.py:762-763)."""
    raw = np.frombuffer(s.encode(), dtype=np.uint8)
    lut = (np.arange(256) - 48).astype(np.uint8)
    for ch, v in zip("ACGT", range(4)):
        lut[ord(ch)] = v
    return lut[raw]


def expand_cigar(cigar: str, read: str, start: int, end: int) -> np.ndarray:
    """query_arr over the reference span: base 0..3 per aligned column, 4 for deletions,
    insertion columns dropped (sequences.py:768-794). Quality handling is omitted because the
    reference's threshold is fixed at qt=0 so every `addition` is 1.0 (sequences.py:659,735-736)."""
    parts = _CIGAR_RE.findall(cigar)
    lens = np.array([p[0] for p in parts], dtype=np.uint32)
    ops = np.array([_OP_CODE.get(p[1], (ord(p[1]) - 48) & 0xFF) for p in parts], dtype=np.uint8)
    rep = np.repeat(ops, lens)
    notins = np.where(rep != 8)
    notdel = np.where(rep != 7)
    rep[notdel] = _chars_to_int(read)[start:end]       # ValueError on length mismatch, as upstream
    q = rep[notins]
    q[q == 7] = 4
    return q


def best_record(recs):
    """`Paf.choose_best_mapper` (paf.py:710-722): last element of a stable argsort by (mapq, AS)."""
    if len(recs) == 1:
        return recs[0]
    keys = np.array([(r.mapq, r.align_score) for r in recs], dtype=[("q", int), ("dp", int)])
    return recs[np.argsort(keys, order=["q", "dp"])[-1]]


def convert_records(paf_dict, seqs) -> dict[str, list]:
    """{tname: [(start, end, query_arr, barcode)]} (sequences.py:678-739)."""
    out: dict[str, list] = {}
    for rid in list(paf_dict.keys()):
        rec = best_record(paf_dict[rid])
        if rec.rev:
            seq = seqs[rec.qname].translate(_COMP)[::-1]
            qs, qe = rec.qlen - rec.qend, rec.qlen - rec.qstart
        else:
            seq = seqs[rec.qname]
            qs, qe = rec.qstart, rec.qend
        q = expand_cigar(rec.cigar, seq, qs, qe)
        start, end = min(rec.tstart, rec.tend), max(rec.tstart, rec.tend)
        assert (end - start) == q.shape[0]
        out.setdefault(rec.tname, []).append((start, end, q, rec.barcode))
    return out


# ------------------------------------------------------------------------------------------------
# a3  coverage scatter (reference.py:122-144)
# ------------------------------------------------------------------------------------------------
def increment_coverage(c: ContigState, increments: list) -> None:
    c.change_mask.fill(0)
    tmp = np.zeros(c.coverage.shape, dtype=np.uint16)
    for (start, end, q, barcode) in increments:
        idx = np.arange(q.shape[0])
        np.add.at(tmp[start:end], (idx, q, 0 if barcode is None else barcode), np.ones(q.shape[0]))
    c.change_mask[np.where(tmp)[0]] = 1        # Q6: whole row, every barcode
    c.coverage += tmp


# ------------------------------------------------------------------------------------------------
# a8/a9  per-site scores + dropout (sequences.py:398-455, reference.py:148-179)
# ------------------------------------------------------------------------------------------------
def update_scores(c: ContigState, model: ScoreModel) -> None:
    for b in range(c.nb):
        scores = c.scores[:, b]
        entropy = c.entropy[:, b]
        cov = c.coverage[:, :, b]
        cm = c.change_mask[:, b]
        maxed = np.where(cov.sum(axis=1) >= FREEZE)[0]
        cm[maxed] = False
        pos = np.nonzero(cm)[0]
        pats = cov[pos]
        s_new, e_new = model.lookup(pats, c.seq_int[pos])
        scores[pos] = s_new
        scores[maxed] = TINY
        missing = np.argwhere(scores == 0.0).flatten()        # Q8: zeroed dropouts get re-scored
        if missing.shape[0]:
            # on-miss values == table values (same function); frozen sites were set to TINY above,
            # so every pattern here has sum < 30
            s_m, e_m = model.lookup(cov[missing], c.seq_int[missing])
            scores[missing] = s_m
            entropy[missing] = e_m
        entropy[pos] = e_new
        c.scores[:, b] = scores
        c.entropy[:, b] = entropy


def modify_scores(c: ContigState, mod: int = 8) -> int:
    covsum = np.sum(c.coverage, axis=1)
    if np.mean(covsum) > 5:
        thr = int(np.mean(covsum) / mod)
        rows = np.where(covsum <= thr)[0]
        c.scores[rows] = 0
        return int(rows.shape[0])
    return 0


# ------------------------------------------------------------------------------------------------
# a10  buckets (reference.py:183-211; utils.py:192-226)
# ------------------------------------------------------------------------------------------------
def adjust_length(original_size: int, expanded: np.ndarray) -> np.ndarray:
    d = original_size - expanded.shape[0]
    if d > 0:
        out = np.append(expanded, expanded[-d:], axis=0)
    elif d < 0:
        out = expanded[:-abs(d)]
    else:
        out = expanded
    assert out.shape[0] == original_size
    return out


def check_buckets(c: ContigState, threshold: float) -> None:
    for b in range(c.nb):
        sw = c.bucket_switches[:, b]
        csum = np.sum(c.coverage[:, :, b], axis=1)
        wsum = np.sum(csum[: (len(csum) // BUCKET) * BUCKET].reshape(-1, BUCKET), axis=1)
        mean = adjust_length(sw.shape[0], np.divide(wsum, BUCKET))      # Q9: last bucket = copy of previous
        sw[np.where(mean >= threshold)] = 1
        if len(np.bincount(sw)) == 2 and not all(c.switched_on):
            c.switched_on[np.logical_not(c.switched_on)] = True
        c.bucket_switches[:, b] = sw


# ------------------------------------------------------------------------------------------------
# a11/a12  binning, S_mu, staircase benefit (reference.py:215-269)
# ------------------------------------------------------------------------------------------------
def calc_smu(c: ContigState, mu: int = 400) -> None:
    n = c.length // WINDOW + 1
    c.smu = np.zeros((n, 2, c.nb))
    c.scores_ds = np.zeros((n, c.nb))
    for b in range(c.nb):
        np.add.at(c.scores_ds[:, b], np.arange(c.length) // WINDOW, c.scores[:, b])   # sequential order
        c.smu[:, 0, b] = move_sum(c.scores_ds[::-1, b], window=mu // WINDOW, min_count=1)[::-1]
        c.smu[:, 1, b] = move_sum(c.scores_ds[:, b], window=mu // WINDOW, min_count=1)


def calc_u(c: ContigState, approx_ccl: np.ndarray) -> None:
    w = approx_ccl // WINDOW
    mult = np.arange(0.05, 1, 0.1)[::-1]
    c.expected_benefit = np.zeros((c.scores_ds.shape[0], 2, c.nb))
    for b in range(c.nb):
        acc = np.zeros((c.scores_ds.shape[0], 2))
        for i in range(10):
            fwd = move_sum(c.scores_ds[::-1, b], window=int(w[i]), min_count=1)[::-1]
            rev = move_sum(c.scores_ds[:, b], window=int(w[i]), min_count=1)
            acc[:, 0] += fwd * mult[i]
            acc[:, 1] += rev * mult[i]
        c.expected_benefit[:, :, b] = acc
    c.additional_benefit = c.expected_benefit - c.smu
    c.additional_benefit[c.additional_benefit < 0] = 0


# ------------------------------------------------------------------------------------------------
# a14  read-start distribution (readstartdist.py:11-152)
# ------------------------------------------------------------------------------------------------
class ReadStarts:
    def __init__(self, contigs: dict[str, ContigState], alpha: float = 1.0, p0: float = 0.1):
        self.alpha, self.p0 = alpha, p0
        self.counts = {k: np.zeros((int(c.length / RSD_WINDOW), 2)) for k, c in contigs.items()}
        self.target_size = int(np.sum([c.length for c in contigs.values()]) // 100)
        # upstream asserts that the expanded array is less than one window short of the target (readstartdist.py:131); that
        # fails for references of more than ~150 contigs (int(L/2000)*20 drifts from L//100 by up to 19 rows per contig).
        # strict = False evaluates the same expressions without the assertion (what the GPU path supports).
        self.strict = True

    def count(self, paf_dict) -> None:
        fwd: dict[str, list] = {}
        rev: dict[str, list] = {}
        for rid in paf_dict.keys():
            rec = best_record(paf_dict[rid])
            (rev if rec.rev else fwd).setdefault(rec.tname, []).append(rec.tend if rec.rev else rec.tstart)
        for k, arr in self.counts.items():
            nw = int(arr.shape[0])
            for col, src in ((0, fwd), (1, rev)):
                arr[:, col] += np.histogram(src.get(k, []), bins=nw, range=(0, RSD_WINDOW * nw))[0].astype(float)

    def fhat_windows(self) -> np.ndarray:
        """F-hat per 2 kb window and strand, before expansion (readstartdist.py:86-115)."""
        merged = np.concatenate(list(self.counts.values()))
        nw = merged.shape[0]
        fhat = np.zeros(merged.shape)
        nzi = np.nonzero(merged)
        nz = merged[nzi]
        csum = np.sum(nz)
        fhat[nzi] = np.divide(np.add(self.alpha, nz), 2 * nw * self.alpha + csum)
        rhs = self.alpha / (2 * nw * self.alpha + csum)
        bnum = np.exp(betaln(self.alpha, ((2 * nw - 1) * self.alpha + csum)))
        bden = np.exp(betaln(self.alpha, ((2 * nw - 1) * self.alpha))) or 1e-20
        p0_bit = self.p0 / (self.p0 + (1 - self.p0))
        zero = np.ones(fhat.shape, dtype=bool)
        zero[nzi] = 0
        fhat[zero] = (1 - p0_bit * (bnum / bden)) * rhs
        return fhat

    def fhat(self) -> np.ndarray:
        """Expanded x20, tail-fixed to target_size, normalised to sum 1 (readstartdist.py:121-152)."""
        f = np.repeat(self.fhat_windows(), RSD_WINDOW // WINDOW, axis=0)
        d = self.target_size - f.shape[0]
        assert d < RSD_WINDOW or not self.strict
        if d > 0:
            f = np.append(f, f[-d:], axis=0)
        elif d < 0:
            f = f[:-abs(d)]
        s = np.sum(f)
        if s != 0:
            f = np.multiply(f, 1 / s)
        return f


# ------------------------------------------------------------------------------------------------
# a15  read-length distribution (readlengthdist.py:7-97)
# ------------------------------------------------------------------------------------------------
class ReadLengths:
    def __init__(self, mu: int = 400, sd: int = 4000, lam: int = 6000, eta: int = 11):
        self.mu, self.eta, self.lam = mu, eta, lam
        self.hist = np.zeros(int(1e6), dtype=np.uint16)
        x = np.arange(int(lam + 10 * sd), dtype=int)
        L = np.exp(-((x - lam + 1) ** 2) / (2 * (sd ** 2))) / (sd * np.sqrt(2 * np.pi))
        L /= sum(L)
        self.L = L
        self.approx_ccl = self._staircase()

    def update(self, read_lengths: dict) -> None:
        for _, n in read_lengths.items():
            if n > self.mu * 2:
                self.hist[min(int(n), int(1e6) - 1)] += 1
        seen = np.nonzero(self.hist)
        if len(seen[0]) == 0:
            return
        self.lam = np.sum(seen * self.hist[seen]) / np.sum(self.hist[seen])
        self.longest_read = np.max(np.where(self.hist))
        self.L = np.copy(self.hist[: self.longest_read + 1]).astype("float64")
        self.L /= sum(self.L)
        self.approx_ccl = self._staircase()
        self.time_cost = self.lam - 400 - 300          # Q14: only exists after a successful update

    def _staircase(self) -> np.ndarray:
        ccl = np.zeros(len(self.L) + 1)
        ccl[0] = 1
        ccl[1:] = 1 - np.concatenate((self.L[1:].cumsum(), np.ones(1)))
        ccl[ccl < 1e-6] = 0
        ccl = np.concatenate((np.trim_zeros(ccl, trim="b"), np.zeros(1)))
        out = np.zeros(self.eta - 1, dtype="int32")
        i = 0
        for part in range(self.eta - 1):
            prob = 1 - (part + 0.5) / (self.eta - 1)
            while (ccl[i] > prob) and (len(ccl) > i):
                i += 1
            out[part] = i
        return out


# ------------------------------------------------------------------------------------------------
# a17  threshold (sequences.py:566-649) — threads replaced by the same 12-way split, summed in order
# ------------------------------------------------------------------------------------------------
def find_strategy(benefit: np.ndarray, smu: np.ndarray, fhat: np.ndarray, time_cost: float):
    alpha = rho = 300 // WINDOW
    mu = 400 // WINDOW
    tc = time_cost // WINDOW
    flat = benefit.flatten("F")
    nzi = np.nonzero(flat)
    nz = flat[nzi]
    norm = np.max(nz)                                   # Q15: ValueError on all-zero benefit
    _, ex = np.frexp(nz / norm)
    ex = np.abs(ex)                                     # Q3
    parts = [np.bincount(a) for a in np.array_split(ex, 12)]
    counts_full = np.zeros(max(p.shape[0] for p in parts), dtype="int")
    for p in parts:
        counts_full[: p.shape[0]] += p
    uniq = np.nonzero(counts_full)[0]
    counts = counts_full[uniq]
    fparts = [np.bincount(a, weights=w) for a, w in
              zip(np.array_split(ex, 12), np.array_split(fhat.flatten("F")[nzi], 12))]
    f_grid = np.zeros(max(p.shape[0] for p in fparts), dtype="float")
    for p in fparts:
        f_grid[: p.shape[0]] += p
    f_grid = f_grid[uniq]
    f_mean = f_grid / counts
    bins = np.power(2.0, -uniq) * norm
    ubar0 = np.sum(fhat * smu)
    cs_u = np.cumsum(bins * f_mean * counts) + ubar0
    cs_t = np.cumsum(tc * counts * f_mean) + (alpha + rho + mu)
    k = int(np.argmax(cs_u / cs_t)) + 1
    thr = bins[k] if k < bins.shape[0] else bins[-1]    # Q4
    diag = dict(normaliser=norm, exponents=uniq, counts=counts, f_grid=f_grid, ubar0=ubar0,
                cs_u=cs_u, cs_t=cs_t, k=k)
    return np.where(benefit >= thr, True, False), float(thr), diag


# ------------------------------------------------------------------------------------------------
# a19  the update proper (core.py:77-198)
# ------------------------------------------------------------------------------------------------
class OracleRun:
    """Sequencing of a3, a8-a12, a14, a16-a18 exactly as `BossRuns` (core.py:23-55, 160-224)."""

    def __init__(self, records, ploidy=1, reject_refs=(), barcodes=None, bucket_threshold=5):
        self.barcodes = barcodes
        self.nb = len(barcodes) if barcodes else 1
        self.contigs = make_contigs(records, reject_refs=set(reject_refs), barcodes=barcodes)
        self.contigs_filt = {k: c for k, c in self.contigs.items() if not c.rej}
        self.n_sites = int(np.sum([c.length for c in self.contigs.values()]))      # counts 4 bp per reject ref
        self.model = ScoreModel(ploidy)
        self.model.build_table()
        self.read_starts = ReadStarts(self.contigs_filt)
        self.rl = ReadLengths()
        self.bucket_threshold = bucket_threshold
        self.threshold = None
        self.diag = None
        self.timing: dict[str, float] = {}

    # -- coverage ---------------------------------------------------------------------------------
    def ingest(self, paf_dict, seqs) -> None:
        inc = convert_records(paf_dict, seqs)
        for k, c in self.contigs_filt.items():
            increment_coverage(c, inc.get(k, []))

    # -- update_wrapper (core.py:160-198) -----------------------------------------------------------
    def update(self) -> bool:
        for c in self.contigs_filt.values():
            update_scores(c, self.model)
            modify_scores(c)
        for c in self.contigs_filt.values():
            check_buckets(c, self.bucket_threshold)
        if not any(any(c.switched_on) for c in self.contigs.values()):
            return False
        fhat = self.read_starts.fhat()
        fhat = np.repeat(fhat[:, :, np.newaxis], self.nb, axis=2)
        for c in self.contigs_filt.values():
            calc_smu(c)
            calc_u(c, self.rl.approx_ccl)
        benefit = np.concatenate([c.additional_benefit for c in self.contigs_filt.values()])
        target = self.n_sites // 100
        ben = adjust_length(target, benefit)
        smu = adjust_length(target, benefit)                 # Q1: upstream passes benefit twice
        fh = adjust_length(target, fhat)
        assert fh.shape == ben.shape == smu.shape
        strat, thr, diag = find_strategy(ben, smu, fh, self.rl.time_cost)
        self.threshold, self.diag, self.merged_strat = thr, diag, strat
        self.benefit_adj, self.fhat_adj = ben, fh
        self._distribute(strat)
        return True

    def _distribute(self, strat: np.ndarray) -> None:
        """core.py:125-155, incl. Q2: consumes L//100 rows per contig from an array that holds L//100+1."""
        i = 0
        for c in self.contigs_filt.values():
            rows = c.strat.shape[0]
            gate = adjust_length(rows, np.repeat(c.bucket_switches, BUCKET // WINDOW, axis=0))
            cs = strat[i: i + c.length // WINDOW, :]
            assert cs.shape == c.strat.shape
            for b in range(self.nb):
                c.strat[gate[:, b], :, b] = cs[gate[:, b], :, b]
            i += c.length // WINDOW

    def strategy_dict(self) -> dict[str, np.ndarray]:
        return {k: c.strat for k, c in self.contigs.items()}
