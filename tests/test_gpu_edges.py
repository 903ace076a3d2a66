"""GPU edge cases of the coverage update and the strategy update, each against the oracle or against upstream's
documented behaviour: empty batches, characters outside ACGT (aligned -> IndexError, inside an insertion -> ignored),
CIGAR/interval mismatches, untracked contigs, barcode fallback (Q11), uint16 wrap (Q13), staircase windows < 1,
a missing time_cost (Q14), the packed and the text ingest agreeing, and the mirror following the device state."""
import io

import numpy as np
import pytest

import helpers as H
from oracle import boss_oracle as bo
from boss_runs_b200 import synth
from boss_runs_b200.hostmodel import parse_PAF

pytestmark = pytest.mark.gpu


def make_run(lengths, seed=5, **kw):
    from boss_runs_b200.runs import BossRuns
    contigs = synth.random_contigs(lengths, seed=seed)
    return contigs, BossRuns(contigs=contigs, write_debug=True, **kw)


def paf_line(rid, read, qs, qe, strand, tname, tlen, ts, te, cigar):
    return "\t".join(map(str, (rid, len(read), qs, qe, strand, tname, tlen, ts, te, te - ts, te - ts, 60, "AS:i:100", "tp:A:P", "s1:i:50",
                               f"cg:Z:{cigar}"))) + "\n"


def test_empty_batch_and_untracked_contigs(lib):
    contigs, run = make_run({"a": 120_000, "b": 100_000}, bucket_threshold=0)
    run.rl_dist.update({"x": 5000})
    run.process_batch_runs({}, {})                                   # nothing mapped: still a full update
    assert run.last.switched_on and (run.contigs["a"].coverage == 0).all()
    read = contigs["a"][1000:1300]
    pd = parse_PAF(io.StringIO(paf_line("r1", read, 0, 300, "+", "elsewhere", 500_000, 1000, 1300, "300M")))
    run.process_batch_runs(pd, {"r1": read})                         # reads on contigs nobody tracks are dropped (core.py:83-86)
    assert (run.contigs["a"].coverage == 0).all() and (run.contigs["b"].coverage == 0).all()


def test_non_acgt_characters(lib):
    contigs, run = make_run({"a": 120_000}, bucket_threshold=0)
    ref = contigs["a"]
    # an N inside an INSERTION is never looked at upstream (the column is dropped): the counts must match the oracle
    read = ref[5000:5100] + "NN" + ref[5100:5200]
    pd = parse_PAF(io.StringIO(paf_line("ins", read, 0, 202, "+", "a", 120_000, 5000, 5200, "100M2I100M")))
    orc = H.oracle_run(list(contigs.items()), 1, [], None, 0)
    orc.ingest(pd, {"ins": read})
    run._effect_increments(run.cc.convert_records(pd, {"ins": read}))
    assert np.array_equal(run.contigs["a"].coverage, orc.contigs["a"].coverage)
    # the same on the reverse strand (walked backwards, complemented)
    rc = read.translate(str.maketrans("ATGC", "TACG"))[::-1]
    pd = parse_PAF(io.StringIO(paf_line("insr", rc, 0, 202, "-", "a", 120_000, 5000, 5200, "100M2I100M")))
    orc.ingest(pd, {"insr": rc})
    run._effect_increments(run.cc.convert_records(pd, {"insr": rc}))
    assert np.array_equal(run.contigs["a"].coverage, orc.contigs["a"].coverage)
    # an N in an ALIGNED column indexes outside the five counters upstream: IndexError (reference.py:138-140)
    bad = ref[7000:7050] + "N" + ref[7051:7100]
    pd = parse_PAF(io.StringIO(paf_line("bad", bad, 0, 100, "+", "a", 120_000, 7000, 7100, "100M")))
    with pytest.raises(IndexError):
        run._effect_increments(run.cc.convert_records(pd, {"bad": bad}))
    # characters '0'..'4' translate to codes 0..4 upstream (ord - 48): '4' lands in the deletion column
    odd = ref[9000:9010] + "4" + ref[9011:9020]
    pd = parse_PAF(io.StringIO(paf_line("odd", odd, 0, 20, "+", "a", 120_000, 9000, 9020, "20M")))
    before = run.contigs["a"].coverage[9010].copy()
    run._effect_increments(run.cc.convert_records(pd, {"odd": odd}))
    after = run.contigs["a"].coverage[9010]
    assert after[4, 0] == before[4, 0] + 1


def test_shape_mismatches_raise_like_upstream(lib):
    contigs, run = make_run({"a": 120_000}, bucket_threshold=0)
    read = contigs["a"][100:300]
    for cigar, qe, te in (("150M", 200, 300),        # CIGAR consumes fewer read bases than the slice holds
                          ("200M", 200, 290),        # CIGAR spans more reference than tend - tstart
                          ("100M5D100M", 200, 300)): # deletion makes the reference span too long
        pd = parse_PAF(io.StringIO(paf_line("r", read, 0, qe, "+", "a", 120_000, 100, te, cigar)))
        with pytest.raises(AssertionError):
            run._effect_increments(run.cc.convert_records(pd, {"r": read}))
    assert (run.contigs["a"].coverage == 0).all(), "a rejected batch must not touch the counters"


def test_barcode_fallback_and_u16_wrap(lib):
    contigs, run = make_run({"a": 100_000}, bucket_threshold=0, barcodes=["barcode01", "barcode02"])
    read = contigs["a"][2000:2100]
    pd = parse_PAF(io.StringIO(paf_line("r", read, 0, 100, "+", "a", 100_000, 2000, 2100, "100M")))
    for rec in pd["r"]:
        rec.barcode = 99                                             # unclassified: index 0 (Q11)
    run._effect_increments(run.cc.convert_records(pd, {"r": read}))
    cov = run.contigs["a"].coverage
    assert cov[2000:2100, :, 0].sum() == 100 and cov[:, :, 1].sum() == 0
    # counters are uint16 and wrap (Q13)
    full = np.zeros((100_000, 5, 2), dtype=np.uint16)
    full[2000:2100, :, 0] = 65535
    run.engine.set_coverage(0, full)
    run._effect_increments(run.cc.convert_records(pd, {"r": read}))
    cov = run.contigs["a"].coverage
    codes = np.frombuffer(read.encode(), dtype=np.uint8)
    lut = np.zeros(256, np.uint8); lut[np.frombuffer(b"ACGT", np.uint8)] = np.arange(4)
    want = full.copy()
    want[np.arange(2000, 2100), lut[codes], 0] += np.uint16(1)       # NumPy wraps the same way
    assert np.array_equal(cov, want)


def test_update_argument_errors(lib):
    contigs, run = make_run({"a": 120_000}, bucket_threshold=0)
    with pytest.raises(ValueError):                                  # bn.move_sum rejects windows < 1 (reference.py:259-260)
        run.engine.update(approx_ccl=np.array([50, 200, 300, 400, 500, 600, 700, 800, 900, 1000]), time_cost=5000.0, bucket_threshold=0,
                          fhat_scalars=(1.0, 100.0, 0.01))
    # Q14: no read-length update yet -> no time_cost -> AttributeError once a bucket is on
    with pytest.raises(AttributeError):
        run.update_wrapper()


def test_packed_and_text_ingest_agree(lib):
    contigs, run_text = make_run({"a": 150_000, "b": 110_000}, bucket_threshold=0)
    _, run_packed = make_run({"a": 150_000, "b": 110_000}, bucket_threshold=0)
    rb = synth.read_batch(contigs, n_reads=700, seed=77, mean_len=2500.0, min_len=300, max_len=9000)
    pd = parse_PAF(io.StringIO(rb.paf_text))
    inc = run_text.cc.convert_records(pd, rb.seqs)
    run_text._effect_increments(inc)
    d = run_packed.pack_for_device(inc)
    run_packed.engine.ingest_packed(d["seg"], d["tstart"], d["barcode"], d["cig_off"], d["cigar"], d["base_off"], d["bases"],
                                    ascii_bases=True)
    orc = H.oracle_run(list(contigs.items()), 1, [], None, 0)
    orc.ingest(pd, rb.seqs)
    for name in contigs:
        assert np.array_equal(run_text.contigs[name].coverage, orc.contigs[name].coverage)
        assert np.array_equal(run_packed.contigs[name].coverage, orc.contigs[name].coverage)
    assert 0 < run_text.engine.ingest_bytes() < sum(len(s) for s in rb.seqs.values())    # 2-bit bases + 4-byte ops < the text


def test_mirror_follows_device_state(lib):
    contigs, run = make_run({"a": 180_000, "b": 120_000}, bucket_threshold=0)
    moved = []
    for b in range(4):
        rb = synth.read_batch(contigs, n_reads=400, seed=300 + b, mean_len=2500.0, min_len=300, max_len=9000)
        pd = parse_PAF(io.StringIO(rb.paf_text))
        run.rl_dist.update({rid: recs[0].qlen for rid, recs in pd.items()})
        run.process_batch_runs(pd, rb.seqs)
        dev = run.engine.strat_all()                                 # explicit device->host copy
        row = 0
        for c in run.contigs_filt.values():
            n = c.length // 100
            assert np.array_equal(c.strat, dev[row: row + n]), f"batch {b}: mirror of {c.name} is stale"
            row += n
        moved.append(run.last.mirror_bytes)
        packed = np.unpackbits(run.engine.strat_packed(), bitorder="little")[: dev.size].astype(bool)
        assert np.array_equal(packed, dev.reshape(-1))
    assert moved[0] > 0 and all(m <= dev.size + 1024 for m in moved)


def test_text_batches_equal_object_batches(lib):
    """`process_batch_text` (PAF text tokenised in C, read starts from arrays) leaves exactly the state
    `process_batch_runs` leaves on the parsed objects of the same text — and that equals the oracle's."""
    contigs, run_obj = make_run({"a": 180_000, "b": 120_000}, bucket_threshold=0, barcodes=["barcode01", "barcode02", "barcode03"])
    _, run_txt = make_run({"a": 180_000, "b": 120_000}, bucket_threshold=0, barcodes=["barcode01", "barcode02", "barcode03"])
    orc = H.oracle_run(list(contigs.items()), 1, [], ["barcode01", "barcode02", "barcode03"], 0)
    for b in range(3):
        rb = synth.read_batch(contigs, n_reads=500, seed=700 + b, mean_len=2500.0, min_len=300, max_len=9000, n_barcodes=3)
        lens = {rid: len(s) for rid, s in rb.seqs.items()}
        pd = H.parse_batch(rb.paf_text, rb.barcodes, True)
        pd = {rid: recs for rid, recs in pd.items() if recs[0].alignment_block_length >= 200}     # mapper.py:64
        for run in (run_obj, run_txt):
            run.rl_dist.update(lens)
        orc.rl.update(lens)
        run_obj.process_batch_runs(pd, rb.seqs)
        run_txt.process_batch_text(rb.paf_text, rb.seqs, barcodes=rb.barcodes)
        orc.ingest(pd, rb.seqs); orc.read_starts.count(pd); orc.update()
        assert run_obj.threshold == run_txt.threshold and run_obj.last.ubar0 == run_txt.last.ubar0
        assert np.array_equal(run_obj.engine.read_starts(), run_txt.engine.read_starts())
        for name in contigs:
            assert np.array_equal(run_obj.contigs[name].coverage, run_txt.contigs[name].coverage)
            assert np.array_equal(run_obj.contigs[name].strat, run_txt.contigs[name].strat)
        H.compare_state(run_txt, orc, True, f"text/b{b}")


def test_split_score_pass_is_invisible(lib):
    """`process_batch_*` announce the batch before ingesting it, so the tiles it will not touch are scored while the
    host packs the reads (bossgpu_prescore). Every number must equal the run that scores all tiles after the scatter:
    bins, dropout counts, switches, thresholds, masks — also after a batch was rejected half-way and after updates
    that came without an announcement."""
    lens = {"a": 260_000, "b": 141_000, "c": 100_000}
    contigs, split = make_run(lens, bucket_threshold=4)
    _, whole = make_run(lens, bucket_threshold=4)
    whole.use_prescore = False
    ref = contigs["a"]
    bad_read = ref[7000:7050] + "N" + ref[7051:7100]
    bad = parse_PAF(io.StringIO(paf_line("bad", bad_read, 0, 100, "+", "a", 260_000, 7000, 7100, "100M")))
    for b in range(5):
        rb = synth.read_batch(contigs, n_reads=500, seed=900 + b, mean_len=2500.0, min_len=300, max_len=9000,
                              focus=("a", 30_000, 33_000, 0.15))
        pd = parse_PAF(io.StringIO(rb.paf_text))
        lens_b = {rid: recs[0].qlen for rid, recs in pd.items()}
        if b == 2:                                                   # announced, then rejected by the scatter (IndexError upstream)
            for run in (split, whole):
                with pytest.raises(IndexError):
                    run.process_batch_runs(bad, {"bad": bad_read})
        for run in (split, whole):
            run.rl_dist.update(lens_b)
            if b == 3:                                               # the text entry point announces too
                run.process_batch_text(rb.paf_text, rb.seqs, min_len=1)
            else:
                run.process_batch_runs(pd, rb.seqs)
        if b == 1:                                                   # an update nobody announced
            for run in (split, whole):
                run.update_wrapper()
        assert split.last.switched_on == whole.last.switched_on
        assert (split.threshold, split.last.ubar0, split.last.n_nonzero, split.last.n_dropout, split.last.n_accept) == \
               (whole.threshold, whole.last.ubar0, whole.last.n_nonzero, whole.last.n_dropout, whole.last.n_accept), f"batch {b}"
        for name in lens:
            s, w = split.contigs[name], whole.contigs[name]
            assert np.array_equal(s.coverage, w.coverage) and np.array_equal(s.bucket_switches, w.bucket_switches)
            assert np.array_equal(s.scores_ds, w.scores_ds), f"batch {b}/{name}: bins differ"
            assert np.array_equal(s.strat, w.strat), f"batch {b}/{name}: masks differ"
    assert split.last.n_dropout > 0 and (split.contigs["a"].coverage.sum(axis=1) >= 30).any()


def test_rejected_batch_leaves_no_trace(lib):
    """A batch that upstream rejects with IndexError (a character outside ACGT in an aligned column) must not reach the
    state at all: upstream discards `tmp_cov` of that contig (reference.py:128-146). The aligned-column check runs
    before any counter or depth total is touched, so coverage, dropout thresholds and every later update equal the
    oracle that never saw the batch — through the text ingest and through the pre-tokenised one."""
    lens = {"a": 200_000}
    contigs, run = make_run(lens, bucket_threshold=0)
    orc = H.oracle_run(list(contigs.items()), 1, [], None, 0)
    ref = contigs["a"]
    for b in range(3):
        rb = synth.read_batch(contigs, n_reads=900, seed=1200 + b, mean_len=3000.0, min_len=300, max_len=9000)
        pd = parse_PAF(io.StringIO(rb.paf_text))
        if b >= 1:
            # the same batch plus one reverse-strand read with an N in an aligned column (the walk is mirrored)
            seqs = dict(rb.seqs)
            good = ref[7000:7200]
            bad_fwd = good[:120] + "N" + good[121:]
            seqs["bad"] = bad_fwd.translate(str.maketrans("ATGC", "TACG"))[::-1]
            pd_bad = parse_PAF(io.StringIO(rb.paf_text + paf_line("bad", seqs["bad"], 0, 200, "-", "a", 200_000, 7000, 7200, "100M2I98M2D")))
            before = run.contigs["a"].coverage
            with pytest.raises(IndexError):
                orc.ingest(pd_bad, seqs)
            if b == 1:      # text ingest: 2-bit bases + the list of other characters
                with pytest.raises(IndexError):
                    run._effect_increments(run.cc.convert_records(pd_bad, seqs))
            else:           # pre-tokenised ingest: bases as bytes
                d = run.pack_for_device(run.cc.convert_records(pd_bad, seqs))
                with pytest.raises(IndexError):
                    run.engine.ingest_packed(d["seg"], d["tstart"], d["barcode"], d["cig_off"], d["cigar"], d["base_off"], d["bases"],
                                             ascii_bases=True, contig_cov_add=d["cov_add"])
            assert np.array_equal(run.contigs["a"].coverage, before), "a rejected batch must not touch the counters"
            assert np.array_equal(orc.contigs["a"].coverage, before)
        H.oracle_step(orc, pd, rb.seqs)
        H.product_step(run, pd, rb.seqs)
        H.compare_state(run, orc, True, f"after-reject/b{b}")
    assert run.last.n_dropout > 0           # the depth rule is active: a leaked depth total would have moved its threshold


def test_second_ingest_after_announced_batch(lib):
    """Announce A, ingest A, ingest B, update (e.g. a caller that runs `_effect_increments` twice per batch): the tiles B
    wrote to carry no marks, so the update has to score every tile again — results equal the run without the split pass."""
    lens = {"a": 300_000, "b": 160_000}
    contigs, split = make_run(lens, bucket_threshold=0)
    _, whole = make_run(lens, bucket_threshold=0)
    whole.use_prescore = False
    for b in range(3):
        rba = synth.read_batch(contigs, n_reads=300, seed=1500 + 2 * b, mean_len=2500.0, min_len=300, max_len=9000)
        rbb = synth.read_batch(contigs, n_reads=300, seed=1501 + 2 * b, mean_len=2500.0, min_len=300, max_len=9000)
        pda, pdb = parse_PAF(io.StringIO(rba.paf_text)), parse_PAF(io.StringIO(rbb.paf_text))
        for run in (split, whole):
            run.rl_dist.update({rid: recs[0].qlen for rid, recs in pda.items()})
            run._prescore_begin()
            inc_a = run.cc.convert_records(pda, rba.seqs)
            run._prescore(inc_a)
            run._effect_increments(inc_a)
            run._effect_increments(run.cc.convert_records(pdb, rbb.seqs))
            run.count_read_starts(pda)
            run.update_wrapper()
        assert (split.threshold, split.last.ubar0, split.last.n_nonzero, split.last.n_dropout) == \
               (whole.threshold, whole.last.ubar0, whole.last.n_nonzero, whole.last.n_dropout), f"batch {b}"
        for name in lens:
            s, w = split.contigs[name], whole.contigs[name]
            assert np.array_equal(s.coverage, w.coverage)
            assert np.array_equal(s.scores_ds, w.scores_ds), f"batch {b}/{name}: bins differ"
            assert np.array_equal(s.strat, w.strat), f"batch {b}/{name}: masks differ"


def test_packed_strategy_file_matches_npz(lib, tmp_path):
    """`strategy_format="both"`: boss.bits (packed on the GPU) and upstream's boss.npz describe the same masks, for a
    reference with a reject ref and barcodes, before the first update and after every update."""
    from boss_runs_b200 import stratfile
    from boss_runs_b200.runs import BossRuns
    contigs = synth.random_contigs({"a": 180_000, "rej": 120_000, "b": 130_000}, seed=8)
    barcodes = ["barcode01", "barcode02"]
    run = BossRuns(contigs=contigs, bucket_threshold=0, barcodes=barcodes, reject_refs="rej", out_dir=str(tmp_path),
                   strategy_format="both")
    sb = stratfile.StrategyBits(tmp_path / "masks", barcodes=barcodes)

    def same():
        assert sb.reload() in (0, 1)
        npz = np.load(tmp_path / "masks" / "boss.npz")
        bits = sb.as_dict()
        assert set(npz.files) == set(bits) == {"a", "rej", "b"}
        for name in npz.files:
            assert np.array_equal(npz[name], bits[name]), name
        assert sb.check_coord("rej", 500, False, "barcode01") == 0
    same()
    tracked = {n: s for n, s in contigs.items() if n != "rej"}
    for b in range(3):
        rb = synth.read_batch(tracked, n_reads=400, seed=40 + b, mean_len=2500.0, min_len=300, max_len=9000, n_barcodes=2)
        pd = H.parse_batch(rb.paf_text, rb.barcodes, True)
        run.rl_dist.update({rid: recs[0].qlen for rid, recs in pd.items()})
        run.process_batch_runs(pd, rb.seqs)
        os_mtime = (tmp_path / "masks" / "boss.bits").stat().st_mtime
        sb.last_mask_mtime = 0.0                                    # force the reload whatever the clock's granularity
        same()
        assert not run.contigs["a"].strat.all(), "the update should have rejected something"
        for pos, rev, bc in ((0, False, "barcode01"), (90_050, True, "barcode02"), (179_999, False, "barcode02")):
            assert sb.check_coord("a", pos, rev, bc) == int(run.contigs["a"].strat[pos // 100, int(rev), int(bc[-1]) - 1])
