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
