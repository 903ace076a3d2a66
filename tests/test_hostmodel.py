"""Host half of the coverage update: the compiled `_fastconv` walk and the Python loop produce identical batches,
and both follow upstream's rules (record choice, slice bounds incl. the truncated-reverse-read quirk Q12, contigs
nobody tracks, error types). CPU only."""
import io

import numpy as np
import pytest

from boss_runs_b200 import build, synth
from boss_runs_b200.hostmodel import PafLine, parse_PAF, best_record
from boss_runs_b200.runs import CoverageConverter, _fastconv

FIELDS = ("contig", "tstart", "tend", "barcode", "rev", "cigar_ptr", "cigar_len", "seq_ptr", "seq_from", "seq_to")


@pytest.fixture(scope="module")
def fc():
    build.build_fastconv()
    mod = _fastconv()
    assert mod is not None, "_fastconv did not build"
    return mod


def same(a, b):
    for f in FIELDS:
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.keep == b.keep and a.n_skipped == b.n_skipped and len(a) == len(b)


def test_fastconv_equals_python_loop(fc):
    contigs = synth.random_contigs({"a": 300_000, "b": 200_000, "c": 120_000}, seed=2)
    rb = synth.read_batch(contigs, n_reads=600, seed=9, mean_len=3000.0, n_barcodes=4)
    pd = parse_PAF(io.StringIO(rb.paf_text))
    for rid, recs in pd.items():
        for r in recs:
            r.barcode = rb.barcodes[rid]
    # a read with two records (the better one wins), one on a contig that is not tracked, one truncated reverse read
    rid0 = next(r for r, recs in pd.items() if recs[0].tname != "c")
    extra = PafLine(pd[rid0][0].line)
    extra.mapq, extra.tstart, extra.barcode = 10, extra.tstart + 7, 1
    pd[rid0].append(extra)
    cc = CoverageConverter({"a": 0, "b": 1})                 # "c" is unknown to the converter
    seqs = dict(rb.seqs)
    rev_rid = next(r for r, recs in pd.items() if recs[0].rev and recs[0].tname != "c")
    seqs[rev_rid] = seqs[rev_rid][:400]                        # Q12: rejected reads are truncated to 400 bases
    a = cc.convert_records(pd, seqs)
    b = cc._convert_records_py(pd, seqs)
    same(a, b)
    assert a.n_skipped == sum(1 for recs in pd.values() if best_record(recs).tname == "c") > 0
    i = list(r for r, recs in pd.items() if best_record(recs).tname != "c").index(rid0)
    assert a.tstart[i] == pd[rid0][0].tstart and a.barcode[i] == pd[rid0][0].barcode   # mapq 60 beats mapq 10
    # slice bounds against upstream's literal expression: reverse_complement(read)[qlen-qend : qlen-qstart]
    comp = str.maketrans("ATGC", "TACG")
    k = 0
    for rid, recs in pd.items():
        rec = best_record(recs)
        if rec.tname == "c":
            continue
        s = seqs[rid]
        want = s.translate(comp)[::-1][rec.qlen - rec.qend: rec.qlen - rec.qstart] if rec.rev else s[rec.qstart: rec.qend]
        got = s[a.seq_from[k]: a.seq_to[k]]
        assert (got.translate(comp)[::-1] if rec.rev else got) == want
        k += 1


def test_routed_conversion_is_the_filtered_conversion(fc):
    """One process per GPU: `convert_records(..., ranges=...)` converts the reads overlapping this process' range of their
    contig — exactly the sub-batch the library would have selected from the whole batch — and lists every tracked read in
    the `all_*` arrays (depth totals, read starts). C helper and Python loop agree."""
    contigs = synth.random_contigs({"a": 300_000, "b": 200_000, "c": 120_000}, seed=2)
    rb = synth.read_batch(contigs, n_reads=500, seed=19, mean_len=3000.0, n_barcodes=2)
    pd = parse_PAF(io.StringIO(rb.paf_text))
    for rid, recs in pd.items():
        for r in recs:
            r.barcode = rb.barcodes[rid]
    cc = CoverageConverter({"a": 0, "b": 1})                 # "c" is unknown to the converter
    whole = cc.convert_records(pd, rb.seqs)
    for ranges in ([[0, 140_000], [0, 0]], [[140_000, 300_000], [0, 60_000]], [[0, 0], [60_000, 200_000]], [[0, 0], [0, 0]]):
        ranges = np.array(ranges, dtype=np.int64)
        a = cc.convert_records(pd, rb.seqs, ranges=ranges)
        b = cc._convert_records_py(pd, rb.seqs, ranges)
        same(a, b)
        t0, t1 = np.minimum(whole.tstart, whole.tend), np.maximum(whole.tstart, whole.tend)
        lo, hi = ranges[whole.contig, 0], ranges[whole.contig, 1]
        mine = (t1 > lo) & (t0 < hi)
        for f in FIELDS:
            assert np.array_equal(getattr(a, f), getattr(whole, f)[mine]), f
        for got, got_py, want in ((a.all_contig, b.all_contig, whole.contig), (a.all_tstart, b.all_tstart, whole.tstart),
                                  (a.all_tend, b.all_tend, whole.tend), (a.all_rev, b.all_rev, whole.rev)):
            assert np.array_equal(got, want) and np.array_equal(got_py, want)
        assert all(np.array_equal(x, y) for x, y in zip(a.whole_batch(), whole.whole_batch()))
    # the three ranges above partition the genome: every read is converted by at least one process, spanning reads by two
    assert whole.all_contig is None


def test_fastconv_errors(fc):
    contigs = synth.random_contigs({"a": 150_000}, seed=3)
    rb = synth.read_batch(contigs, n_reads=5, seed=1, mean_len=2000.0)
    pd = parse_PAF(io.StringIO(rb.paf_text))
    cc = CoverageConverter({"a": 0})
    seqs = dict(rb.seqs)
    victim = next(iter(pd))
    del seqs[victim]
    for fn in (cc.convert_records, cc._convert_records_py):
        with pytest.raises(KeyError):
            fn(pd, seqs)                                       # seqs[rec.qname] upstream
    pd[victim][0].cigar = None
    for fn in (cc.convert_records, cc._convert_records_py):
        with pytest.raises(AssertionError):
            fn(pd, rb.seqs)                                    # sequences.py:718
    assert len(cc.convert_records({}, {})) == 0


def test_paf_parsing_follows_upstream():
    line = "r1\t1000\t10\t990\t-\tctg\t200000\t5000\t5980\t900\t980\t60\tAS:i:880\ttp:A:P\ts1:i:400\tcg:Z:500M2D478M\n"
    sec = line.replace("tp:A:P", "tp:A:S").replace("r1", "r2")
    short = line.replace("\t980\t60", "\t0\t60").replace("r1", "r3")
    pd = parse_PAF(io.StringIO(line + sec + short))
    assert list(pd) == ["r1"]                                  # primary only, block length >= 1 (paf.py:666-669)
    r = pd["r1"][0]
    assert (r.qlen, r.qstart, r.qend, r.rev, r.tname, r.tstart, r.tend, r.mapq, r.align_score, r.cigar, r.primary) == \
        (1000, 10, 990, 1, "ctg", 5000, 5980, 60, 880, "500M2D478M", 1)
    assert r.barcode is None
    assert parse_PAF(12345) == {}


# ---- the text path: PAF text -> batch in one C pass (fastconv.c convert_text) -----------------------------------
def _via_objects(cc, text, seqs, min_len, barcodes=None):
    pd = parse_PAF(io.StringIO(text), min_len=min_len)
    if barcodes:
        for rid, recs in pd.items():
            for r in recs:
                r.barcode = barcodes.get(rid)
    return pd, cc._convert_records_py(pd, seqs)


def same_text(a, b):
    """Same batch; the pointers differ (CIGARs live inside the text on one side, in PafLine attributes on the other)."""
    for f in ("contig", "tstart", "tend", "barcode", "rev", "cigar_len", "seq_ptr", "seq_from", "seq_to"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.n_skipped == b.n_skipped and len(a) == len(b)
    for i in range(len(a)):
        assert a.cigar_bytes(i) == b.cigar_bytes(i) and a.slice_bytes(i) == b.slice_bytes(i)


def test_text_path_equals_object_path(fc):
    contigs = synth.random_contigs({"a": 300_000, "b": 200_000, "7": 120_000, "c": 110_000}, seed=4)
    rb = synth.read_batch(contigs, n_reads=800, seed=21, mean_len=3000.0, min_len=150, n_barcodes=3)
    lines = rb.paf_text.splitlines()
    rng = np.random.default_rng(0)
    extra = []
    for i in rng.choice(len(lines), size=200, replace=False):
        f = lines[i].split("\t")
        kind = int(rng.integers(0, 5))
        if kind == 0:                                            # a second record that loses on mapq
            f[11] = "10"; f[7] = str(int(f[7]) + 3)
        elif kind == 1:                                          # ties on (mapq, AS): the LATER record wins
            f[7] = str(int(f[7]) + 5)
        elif kind == 2:                                          # wins on AS
            f[12] = "AS:i:999999"; f[5] = "b"; f[6] = "200000"
        elif kind == 3:                                          # secondary alignment: filtered
            f[13] = "tp:A:S"; f[11] = "61"
        else:                                                    # too short: filtered by min_len
            f[10] = "150"; f[11] = "61"
        extra.append("\t".join(f))
    # one read with 20 tied records (> 16: NumPy's sort is no longer an insertion sort; the helper asks NumPy)
    f = lines[0].split("\t")
    big = []
    for k in range(20):
        g = list(f); g[0] = "whale"; g[7] = str(int(f[7]) + k); g[11] = str(30 + (k * 7) % 3); g[12] = f"AS:i:{100 + (k * 5) % 4}"
        big.append("\t".join(g))
    text = "\n".join(lines[:400] + extra[:100] + big + lines[400:] + extra[100:])      # no trailing newline, like mapper.py:87
    seqs = dict(rb.seqs)
    seqs["whale"] = seqs[f[0]]
    cc = CoverageConverter({"a": 0, "b": 1, "7": 2})            # "c" is not tracked; "7" is an all-digit name
    bcs = dict(rb.barcodes); bcs["whale"] = 2
    for min_len in (1, 200):
        for barcodes in (None, bcs):
            pd, want = _via_objects(cc, text, seqs, min_len, barcodes)
            got = cc.convert_text(text, seqs, min_len=min_len, barcodes=barcodes)
            same_text(got, want)
            assert got.n_skipped > 0 and len(got) > 600
    # read starts from the arrays == read starts from the objects
    from boss_runs_b200.hostmodel import ReadStartDist

    class C:
        def __init__(self, n): self.length = n
    tracked = {n: C(len(contigs[n])) for n in ("a", "b", "7")}
    r1, r2 = ReadStartDist(tracked), ReadStartDist(tracked)
    pd, want = _via_objects(cc, text, seqs, 200)
    w1, s1 = r1.count_read_starts(pd)
    w2, s2 = r2.count_read_starts_arrays(want.contig, want.tstart, want.tend, want.rev)
    assert np.array_equal(w1, w2) and np.array_equal(s1, s2) and np.array_equal(r1.merge(), r2.merge()) and len(w1) > 600


def test_text_path_names_and_errors(fc):
    contigs = synth.random_contigs({"a": 150_000}, seed=3)
    read = contigs["a"][1000:1400]
    base = "\t".join(["007", "400", "0", "400", "+", "a", "150000", "1000", "1400", "400", "400", "60", "AS:i:400", "tp:A:P", "cg:Z:400M"])
    cc = CoverageConverter({"a": 0})
    # an all-digit read id is an int upstream and comes back as "7" (paf.py:54-56): the lookup uses the canonical text
    got = cc.convert_text(base + "\n", {"7": read})
    assert len(got) == 1 and got.cigar_bytes(0) == b"400M" and got.slice_bytes(0) == read.encode()
    with pytest.raises(KeyError):
        cc.convert_text(base, {"007": read})
    pd, want = _via_objects(cc, base, {"7": read}, 1)
    same_text(got, want)
    assert len(cc.convert_text("", {})) == 0
    for bad, exc in ((base.replace("\tcg:Z:400M", ""), AssertionError),           # no CIGAR (sequences.py:718)
                     ("r\t400\t0\t400\t+\ta", IndexError),                         # fewer than 12 columns
                     (base + "\n\n" + base, IndexError),                           # an empty line inside the text
                     (base + "\tzz:Z:a:b", ValueError),                            # tag with an extra ':'
                     (base + "\tzz:Q:1", KeyError),                                # unknown tag type
                     (base.replace("AS:i:400", "AS:i:x"), ValueError),
                     (base.replace("\t1000\t", "\tx\t"), ValueError)):
        with pytest.raises(exc):
            cc.convert_text(bad, {"7": read})
        if exc is not AssertionError and exc is not ValueError:
            with pytest.raises(exc):
                parse_PAF(io.StringIO(bad))
    # later duplicates of a tag win; '\r\n' line ends are stripped; the strand test is `!= '+'`
    odd = base.replace("tp:A:P", "tp:A:S") + "\ttp:A:P\r\n" + base.replace("\t+\t", "\t-\t").replace("007", "r2") + "\n"
    got = cc.convert_text(odd, {"7": read, "r2": read})
    pd, want = _via_objects(cc, odd, {"7": read, "r2": read}, 1)
    same_text(got, want)
    assert list(got.rev) == [0, 1]


def test_switch_pull_flags_whole_contigs_from_the_flat_image():
    """`BossRuns._pull_switches` with thousands of contigs: one segmented reduction over the library's flat switch image
    (reference.py:203-207: any switch of any barcode flags the whole contig), contigs flagged once stay flagged, and the
    image is read again on every call (the library rewrites it in place)."""
    from types import SimpleNamespace
    from boss_runs_b200.runs import BossRuns
    nb, sizes = 2, [3, 1, 5, 2]
    flat = np.zeros((sum(sizes), nb), dtype=bool)
    views, row = [], 0
    for k in sizes:
        views.append(flat[row: row + k])
        row += k

    class Eng:
        buckets_flat = flat
        def buckets_host(self):
            return views
    run = BossRuns.__new__(BossRuns)
    run.engine = Eng()
    run._switch_views = None
    run.contigs_filt = {f"c{i}": SimpleNamespace(switched_on=np.zeros(nb, dtype=bool), bucket_switches=None) for i in range(len(sizes))}
    run._pull_switches()
    assert not any(c.switched_on.any() for c in run.contigs_filt.values())
    assert all(c.bucket_switches is v for c, v in zip(run.contigs_filt.values(), views))
    flat[3, 1] = True                       # contig 1 (one row), barcode 1
    flat[8, 0] = True                       # contig 2, last row
    run._pull_switches()
    assert [bool(c.switched_on.all()) for c in run.contigs_filt.values()] == [False, True, True, False]
    flat[0, 0] = True
    flat[10, 1] = True
    run._pull_switches()
    assert all(c.switched_on.all() for c in run.contigs_filt.values())
    run._pull_switches()                    # everything flagged: nothing left to do
    # without a flat image (another engine type) the per-contig test gives the same answer
    run2 = BossRuns.__new__(BossRuns)
    run2.engine = SimpleNamespace(buckets_host=lambda: views)
    run2._switch_views = None
    run2.contigs_filt = {f"c{i}": SimpleNamespace(switched_on=np.zeros(nb, dtype=bool), bucket_switches=None) for i in range(len(sizes))}
    flat[:] = False
    flat[4, 0] = True
    run2._pull_switches()
    assert [bool(c.switched_on.all()) for c in run2.contigs_filt.values()] == [False, False, True, False]
