"""Known-answer tests the upstream reference holds for this path, asserted against the oracle
AND against the product's host mirrors. Literal numbers are the ones upstream's own tests assert
(tests/base/test_runs_sequences.py, test_readlengthdist.py); tests/golden/kats.npz holds values re-derived
by running the reference here (oracle/make_golden.py:make_kats)."""
import hashlib
import io

from pathlib import Path

import numpy as np
import pytest

from oracle import boss_oracle as bo
from boss_runs_b200 import hostmodel, priors
from boss_runs_b200.runs import CoverageConverter

GOLDEN = __import__("golden_io").GOLDEN


@pytest.fixture(scope="module")
def kats():
    return dict(np.load(GOLDEN / "kats.npz", allow_pickle=False))


@pytest.mark.parametrize("ploidy", [1, 2])
def test_priors_and_phi(kats, ploidy):
    m = bo.ScoreModel(ploidy)
    p = priors.Priors(ploidy=ploidy)
    len_g = 5 if ploidy == 1 else 15
    assert m.phi.shape == p.phi.shape == (5, len_g)                 # test_runs_sequences.py:9-61
    assert m.priors.shape == p.priors.shape == (4, len_g)
    for mine in (m.phi, p.phi):
        assert np.array_equal(mine, kats[f"p{ploidy}_phi"])
    for mine in (m.priors, p.priors):
        assert np.array_equal(mine, kats[f"p{ploidy}_priors"])
    assert np.array_equal(m.phi_pow[:, :, :30], kats[f"p{ploidy}_phi_pow30"])
    assert np.array_equal(p.phi_stored[:, :, :30], kats[f"p{ploidy}_phi_pow30"])
    s = priors.Scoring(ploidy=ploidy)
    assert m.score0 == float(kats[f"p{ploidy}_score0"][0]) == float(s.score0[0])
    assert m.ent0 == float(kats[f"p{ploidy}_ent0"][0]) == float(s.ent0[0])


def test_score0_literals():
    m = bo.ScoreModel(1)
    assert np.allclose(m.score0, 0.04969294)                         # test_runs_sequences.py:114-115
    assert np.allclose(m.ent0, 0.09302521)
    with pytest.raises(ValueError):
        bo.phi_matrix(3)                                             # sequences.py:29
    with pytest.raises(ValueError):
        priors.Scoring(ploidy=3)


@pytest.mark.parametrize("ploidy", [1, 2])
def test_pattern_scores(kats, ploidy):
    m = bo.ScoreModel(ploidy)
    pats = kats[f"p{ploidy}_patterns"]
    en, sc = bo.pattern_scores(pats, m.priors, m.phi, m.phi_pow)
    assert np.array_equal(sc, kats[f"p{ploidy}_pattern_scores"])
    assert np.array_equal(en, kats[f"p{ploidy}_pattern_entropies"])


def test_score_table_literals(kats):
    m = bo.ScoreModel(1)
    m.build_table()
    r = int(bo.pattern_rank(np.array([[2, 0, 0, 0, 0]]))[0])
    assert np.allclose(m.score_table[r, 3], 0.17253973305650225)     # test_runs_sequences.py:120-125
    assert np.allclose(m.entropy_table[r, 3], 0.22957118271635163)
    assert m.score_table[r, 3] == float(kats["score_arr_2_0_0_0_0_3"])
    assert m.entropy_table[r, 3] == float(kats["entropy_arr_2_0_0_0_0_3"])
    assert int(kats["score_arr_n_prefilled"]) == 136_982             # SURVEY §8 a7 [measured]
    # rank is a bijection onto [0, C(34,5))
    pats = bo.all_patterns()
    assert pats.shape == (bo.N_PATTERNS, 5) and pats.sum(axis=1).max() == 29
    assert np.array_equal(bo.pattern_rank(pats), np.arange(bo.N_PATTERNS))
    # the all-zero pattern under the diploid model is NOT the contig's score0 (Q5)
    d = bo.ScoreModel(2)
    en, sc = bo.pattern_scores(np.zeros((1, 5), dtype=np.int64), d.priors, d.phi, d.phi_pow)
    assert abs(sc[0, 0] - 0.0597927) < 1e-6 and abs(m.score0 - 0.04969294) < 1e-8


def test_default_staircase(kats):
    want = [1167, 2729, 3903, 4918, 5866, 6808, 7797, 8912, 10321, 12713]   # test_readlengthdist.py:28-31
    assert list(kats["default_approx_ccl"]) == want
    assert list(bo.ReadLengths().approx_ccl) == want
    assert list(hostmodel.ReadlengthDist().approx_ccl) == want
    assert not hasattr(hostmodel.ReadlengthDist(), "time_cost")      # Q14


def test_readlength_update_matches_oracle():
    rng = np.random.default_rng(3)
    lens = {f"r{i}": int(x) for i, x in enumerate(np.clip(rng.gamma(4, 2500, size=3000), 100, 2_000_000))}
    a, b = bo.ReadLengths(), hostmodel.ReadlengthDist()
    for chunk in (dict(list(lens.items())[:1000]), dict(list(lens.items())[1000:])):
        a.update(chunk)
        b.update(chunk)
        assert np.array_equal(a.approx_ccl, b.approx_ccl)
        assert a.lam == b.lam and a.time_cost == b.time_cost
        assert np.array_equal(a.L, b.L) and a.longest_read == b.longest_read          # bit for bit
    empty = hostmodel.ReadlengthDist()
    empty.update({"r": 500})                                         # below 2*mu: ignored, no time_cost yet
    assert not hasattr(empty, "time_cost")


def test_readlength_model_equals_upstream_class():
    """The array-shaped `hostmodel.ReadlengthDist` against upstream's own class (imported when the reference is mounted;
    boss/readlengthdist.py needs nothing but NumPy): prior, every attribute after each update, whales, short reads,
    single-length and bimodal histograms."""
    import sys
    ref = Path("/root/reference")
    if not (ref / "boss" / "readlengthdist.py").is_file():
        pytest.skip("reference checkout not mounted (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_upstream_rld", ref / "boss" / "readlengthdist.py")
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode, old = True, sys.dont_write_bytecode
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = old
    rng = np.random.default_rng(11)
    for kw in ({}, dict(sd=1000, lam=3000), dict(eta=6)):
        up, mine = mod.ReadlengthDist(**kw), hostmodel.ReadlengthDist(**kw)
        assert np.array_equal(up.L, mine.L) and np.array_equal(up.ccl, mine.ccl) and np.array_equal(up.approx_ccl, mine.approx_ccl)
        assert mine.approx_ccl.dtype == up.approx_ccl.dtype
        batches = [
            {f"a{i}": int(x) for i, x in enumerate(np.clip(rng.gamma(4, 2500, size=500), 50, 3_000_000))},
            {"w1": 1_000_000, "w2": 5_000_000, "s": 799, "t": 800, "u": 801},
            {f"b{i}": 12_345 for i in range(70_000)},                                    # uint16 counter wraps
            {f"c{i}": int(x) for i, x in enumerate(np.concatenate([rng.normal(1500, 50, 400), rng.normal(60_000, 3000, 400)]))},
        ]
        for bt in batches:
            up.update(bt)
            mine.update(bt)
            for attr in ("lam", "longest_read", "time_cost"):
                assert getattr(up, attr) == getattr(mine, attr), attr
            for attr in ("read_lengths", "L", "ccl", "approx_ccl"):
                assert np.array_equal(getattr(up, attr), getattr(mine, attr)), attr


def _real_batch(kats):
    paf = kats["real_paf"].tobytes().decode()
    rids = [str(x) for x in kats["real_rids"]]
    reads = dict(zip(rids, kats["real_reads"].tobytes().decode().split("\n")))
    return paf, reads


def test_convert_records_real_data(kats):
    """A slice of upstream's own PAF/FASTQ fixture through `convert_records`: interval, order and the expanded
    query array of every increment, as the reference produced them."""
    paf, reads = _real_batch(kats)
    pd = hostmodel.parse_PAF(io.StringIO(paf))
    inc = bo.convert_records(pd, reads)
    rows = [(t, s, e, q) for t, lst in inc.items() for (s, e, q, bc) in lst]
    assert [r[0] for r in rows] == [str(x) for x in kats["real_inc_tname"]]
    assert np.array_equal([r[1] for r in rows], kats["real_inc_start"])
    assert np.array_equal([r[2] for r in rows], kats["real_inc_end"])
    assert np.array_equal(np.concatenate([r[3] for r in rows]), kats["real_inc_query_concat"])
    for r, want in zip(rows, kats["real_inc_sha"]):
        assert hashlib.sha256(np.ascontiguousarray(r[3]).tobytes()).hexdigest() == str(want)


def test_product_host_half_matches_oracle_on_real_data(kats, lib):
    """The product's host half (record choice, slice bounds, C++ CIGAR tokenizer, reverse complement) expands to
    exactly the oracle's query arrays. No device call is made: bossgpu_tokenize_cigar is plain host code."""
    import ctypes as C
    paf, reads = _real_batch(kats)
    pd = hostmodel.parse_PAF(io.StringIO(paf))
    names = []
    for recs in pd.values():
        t = hostmodel.best_record(recs).tname
        if t not in names:
            names.append(t)
    cc = CoverageConverter({n: i for i, n in enumerate(names)})
    batch = cc.convert_records(paf_dict=pd, seqs=reads)
    assert len(batch) == len(pd) and batch.n_skipped == 0
    inc = bo.convert_records(pd, reads)
    want = {t: list(lst) for t, lst in inc.items()}
    comp = str.maketrans("ATGC", "TACG")
    cig_off, cig_text, seq_off, seq_text = batch.texts()
    for i in range(len(batch)):
        t = names[batch.contig[i]]
        s, e, q, _ = want[t].pop(0)
        assert (min(batch.tstart[i], batch.tend[i]), max(batch.tstart[i], batch.tend[i])) == (s, e)
        text = cig_text[cig_off[i]: cig_off[i + 1]]
        ops = np.empty(len(text) // 2 + 1, dtype=np.uint32)
        r, qs = C.c_int64(), C.c_int64()
        k = lib.bossgpu_tokenize_cigar(text, len(text), ops.ctypes.data, len(ops), C.byref(r), C.byref(qs))
        assert k > 0 and r.value == e - s
        sl = seq_text[seq_off[i]: seq_off[i + 1]].decode()
        assert qs.value == len(sl)
        if batch.rev[i]:
            sl = sl.translate(comp)[::-1]
        codes = np.frombuffer(sl.encode(), dtype=np.uint8)
        lut = np.full(256, 255, dtype=np.uint8)
        lut[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4)
        codes = lut[codes]
        out, qi = [], 0
        for op in ops[:k]:
            n, cls = int(op >> 4), int(op & 15)
            if cls == 0:
                out.append(codes[qi: qi + n]); qi += n
            elif cls == 1:
                qi += n
            else:
                out.append(np.full(n, 4, dtype=np.uint8))
        assert np.array_equal(np.concatenate(out), q), f"read {i}"


def test_read_start_distribution_matches_oracle(kats):
    paf, _ = _real_batch(kats)
    pd = hostmodel.parse_PAF(io.StringIO(paf))

    class C_:
        def __init__(self, n):
            self.length = n
    lens = {}
    for recs in pd.values():
        r = hostmodel.best_record(recs)
        lens[r.tname] = r.tlen
    contigs = {n: C_(L) for n, L in lens.items()}
    orc = bo.ReadStarts(contigs)
    orc.count(pd)
    mine = hostmodel.ReadStartDist(contigs=contigs, strict=False)
    wins, strands = mine.count_read_starts(pd)
    assert 0 < len(wins) <= len(pd)                                  # starts behind the last whole window are dropped, like np.histogram
    assert len(wins) == int(np.concatenate(list(orc.counts.values())).sum())
    assert np.array_equal(mine.merge(), np.concatenate(list(orc.counts.values())))
    fw = mine.update_f_pointmass()
    assert np.array_equal(fw, orc.fhat_windows())
    a, denom, zero = mine.pointmass_scalars()
    cnt = mine.merge()
    assert np.array_equal(np.where(cnt > 0, (a + cnt) / denom, zero), fw)      # what k_fhat_from_counts evaluates
    f = orc.fhat()
    assert np.array_equal(mine.expand(fw), f)
    assert abs(f.sum() - 1.0) < 1e-12 and f.shape == (orc.target_size, 2)


def test_move_sum_restatement():
    from oracle.move_sum import move_sum, move_sum_loop
    rng = np.random.default_rng(0)
    a = rng.random(500) * np.where(rng.random(500) < 0.2, 0.0, 1.0)
    for w in (1, 4, 17, 500):
        assert np.array_equal(move_sum(a, w, min_count=1), move_sum_loop(a, w, min_count=1))
        direct = np.array([a[max(0, i - w + 1): i + 1].sum() for i in range(len(a))])
        assert np.allclose(move_sum(a, w, min_count=1), direct, rtol=1e-12, atol=1e-12)
    with pytest.raises(ValueError):
        move_sum(a, 0, min_count=1)                                  # Bottleneck rejects window < 1 (a12)
    with pytest.raises(ValueError):
        move_sum(a, 501, min_count=1)
