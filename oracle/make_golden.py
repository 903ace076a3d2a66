"""ORACLE / TEST INFRASTRUCTURE — generate tests/golden/*.npz by running the UPSTREAM reference itself.

Runs only in the build container (needs /root/reference). The reference's own modules are imported
unmodified from /root/reference with the import shims in oracle/shims/ (Bottleneck, mappy, minknow_api are
not installed; only `bn.move_sum` carries arithmetic and is restated in oracle/move_sum.py). Each case
drives `boss.runs.core.BossRuns` exactly like `process_batch_runs` (core.py:202-224) does after the
mapping call, batch by batch, and records the reference's state after every update.

    python -m oracle.make_golden            # writes tests/golden/case_*.npz and tests/golden/kats.npz

The committed .npz files hold the inputs too (contig sequences, PAF text, reads), so the tests never need
the reference or this script again.
"""
from __future__ import annotations

import hashlib
import io
import os
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REFERENCE = Path(os.environ.get("BOSS_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path[:0] = [str(REPO), str(REPO / "oracle" / "shims"), str(REFERENCE)]

import numpy as np  # noqa: E402

from boss_runs_b200 import synth  # noqa: E402  (input generator only)

GOLDEN = REPO / "tests" / "golden"

CASES = {
    # name: ploidy, barcodes, bucket_threshold, tracked contig lengths, rejected contig lengths, dropped, batches
    "hap_nb1": dict(ploidy=1, nb=0, bucket_threshold=12, tracked=[130_050, 150_000], rejected=[104_000], dropped=[60_000],
                    n_batches=5, reads=230, seed=101),
    "dip_nb1": dict(ploidy=2, nb=0, bucket_threshold=0, tracked=[120_030, 141_999, 100_000], rejected=[], dropped=[],
                    n_batches=4, reads=260, seed=202),
    "hap_nb3": dict(ploidy=1, nb=3, bucket_threshold=2, tracked=[125_000, 110_007], rejected=[100_000], dropped=[],
                    n_batches=4, reads=330, seed=303),
    "dip_nb2": dict(ploidy=2, nb=2, bucket_threshold=0, tracked=[101_234, 100_000, 118_400, 100_099], rejected=[100_001, 100_002],
                    dropped=[99_999], n_batches=3, reads=300, seed=404),
    # 60 reject refs add 240 phantom sites => target rows exceed the merged rows (adjust_length pads, utils.py:215-217)
    "hap_pad": dict(ploidy=1, nb=0, bucket_threshold=0, tracked=[130_050], rejected=[100_000] * 60, dropped=[],
                    n_batches=3, reads=150, seed=505),
}


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build_inputs(spec):
    """Contig table in FASTA order: tracked and rejected interleaved, dropped last."""
    names, seqs, kinds = [], {}, {}
    t = synth.random_contigs({f"trk{i + 1}": n for i, n in enumerate(spec["tracked"])}, seed=spec["seed"])
    order = []
    ti = ri = 0
    rej = spec["rejected"]
    # interleave: tracked, rejected, tracked, ... (so reject refs sit between tracked contigs)
    while ti < len(spec["tracked"]) or ri < len(rej):
        if ti < len(spec["tracked"]):
            order.append(("trk", ti)); ti += 1
        if ri < len(rej):
            order.append(("rej", ri)); ri += 1
    for kind, i in order:
        if kind == "trk":
            name = f"trk{i + 1}"
            seqs[name] = t[name]
        else:
            name = f"rej{i + 1}"
            seqs[name] = "ACGT" * (rej[i] // 4) + "A" * (rej[i] % 4)
        names.append(name)
        kinds[name] = kind
    for i, n in enumerate(spec["dropped"]):
        name = f"short{i + 1}"
        names.append(name)
        seqs[name] = "ACGT" * (n // 4) + "A" * (n % 4)
        kinds[name] = "short"
    return names, seqs, kinds


def make_batches(spec, names, seqs, kinds):
    tracked = {n: seqs[n] for n in names if kinds[n] == "trk"}
    first = next(iter(tracked))
    batches = []
    for b in range(spec["n_batches"]):
        # a pile-up window on the first contig pushes local depth past the freeze threshold (30)
        rb = synth.read_batch(tracked, n_reads=spec["reads"], seed=spec["seed"] * 10 + b, mean_len=2000.0, min_len=400,
                              max_len=9000, n_barcodes=spec["nb"], focus=(first, 30_000, 33_000, 0.18))
        batches.append(rb)
    return batches


def run_reference(spec, names, seqs, kinds, batches, workdir: Path):
    from boss.config import Config
    from boss.paf import Paf
    from boss.runs.core import BossRuns

    fa = workdir / "ref.fa"
    with open(fa, "w") as fh:
        for n in names:
            fh.write(f">{n}\n{seqs[n]}\n")
    (workdir / "ref.mmi").touch()
    cwd = os.getcwd()
    os.chdir(workdir)
    try:
        conf = Config()
        args = conf.args
        args.general.name = "golden"
        args.general.ref = str(fa)
        args.general.mmi = str(workdir / "ref.mmi")
        if spec["nb"]:
            args.general.barcodes = [f"barcode{i + 1:02d}" for i in range(spec["nb"])]
        args.optional.ploidy = spec["ploidy"]
        args.optional.bucket_threshold = spec["bucket_threshold"]
        rej_names = [n for n in names if kinds[n] == "rej"]
        args.optional.reject_refs = ",".join(rej_names) if rej_names else None
        exp = BossRuns(args)
        exp.init()
        captured = {}
        orig = exp.scoring.find_strat_thread

        def spy(benefit, smu, fhat, time_cost):
            strat, thr = orig(benefit=benefit, smu=smu, fhat=fhat, time_cost=time_cost)
            captured.update(benefit=benefit.copy(), fhat=fhat.copy(), threshold=float(thr), strat=strat.copy(),
                            time_cost=float(time_cost))
            return strat, thr

        exp.scoring.find_strat_thread = spy
        out = {}
        for bi, rb in enumerate(batches):
            paf_dict = Paf.parse_PAF(io.StringIO(rb.paf_text))
            for rid, recs in paf_dict.items():
                for r in recs:
                    r.barcode = rb.barcodes.get(rid) if spec["nb"] else None
            exp.rl_dist.update(read_lengths={rid: recs[0].qlen for rid, recs in paf_dict.items()})
            quals = {rid: "5" * len(s) for rid, s in rb.seqs.items()}
            inc = exp.cc.convert_records(paf_dict=paf_dict, seqs=rb.seqs, quals=quals)
            exp._effect_increments(increments=inc)
            exp.read_starts.count_read_starts(paf_dict=paf_dict)
            captured.clear()
            exp.update_wrapper()
            p = f"b{bi}_"
            out[p + "approx_ccl"] = exp.rl_dist.approx_ccl.copy()
            out[p + "time_cost"] = np.float64(getattr(exp.rl_dist, "time_cost", np.nan))
            out[p + "updated"] = np.bool_(bool(captured))
            if captured:
                out[p + "threshold"] = np.float64(captured["threshold"])
                if bi == len(batches) - 1 or spec.get("full_dump", True):
                    out[p + "benefit_adj"] = captured["benefit"]
                    out[p + "fhat_adj"] = captured["fhat"]
                out[p + "merged_strat"] = np.packbits(captured["strat"].ravel())
                out[p + "merged_strat_shape"] = np.array(captured["strat"].shape)
            last = bi == len(batches) - 1
            for cname, c in exp.contigs.items():
                q = f"{p}{cname}_"
                out[q + "strat"] = c.strat.copy()
                if c.rej:
                    continue
                out[q + "coverage_sha"] = np.array(sha(c.coverage))
                out[q + "scores_sha"] = np.array(sha(c.scores))
                out[q + "bucket_switches"] = c.bucket_switches.copy()
                out[q + "switched_on"] = c.switched_on.copy()
                out[q + "scores_sum"] = np.float64(c.scores.sum())
                # a strided sample of the per-site arrays at every batch, the full arrays at the last one
                out[q + "scores_sample"] = c.scores[::97].copy()
                out[q + "coverage_sample"] = c.coverage[::97].copy()
                if last and cname == next(iter(exp.contigs_filt)) and spec.get("full_dump", True):
                    out[q + "coverage"] = c.coverage.copy()
                    out[q + "scores"] = c.scores.copy()
                if captured:
                    out[q + "scores_ds"] = c.scores_ds.copy()
                    out[q + "additional_benefit"] = c.additional_benefit.copy()
                    if last or spec.get("full_dump", True):
                        out[q + "smu"] = c.smu.copy()
                        out[q + "expected_benefit"] = c.expected_benefit.copy()
        out["n_sites"] = np.int64(exp.ref.n_sites)
        out["score0"] = np.float64(exp.scoring.score0[0])
        out["contig_score0"] = np.float64(next(iter(exp.contigs_filt.values())).score0[0])
        return out
    finally:
        os.chdir(cwd)


def pack2bit(codes: np.ndarray) -> np.ndarray:
    """uint8 codes 0..3 -> 4 per byte (big-endian within the byte, np.packbits order)."""
    return np.packbits(np.unpackbits(codes[:, None], axis=1)[:, 6:].ravel())


def pack_inputs(spec, names, seqs, kinds, batches):
    d = {"names": np.array(names), "kinds": np.array([kinds[n] for n in names]),
         "lengths": np.array([len(seqs[n]) for n in names], dtype=np.int64),
         "ploidy": np.int64(spec["ploidy"]), "nb": np.int64(spec["nb"]), "bucket_threshold": np.float64(spec["bucket_threshold"]),
         "n_batches": np.int64(len(batches))}
    lut = np.zeros(256, dtype=np.uint8)
    lut[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
    for n in names:
        if kinds[n] == "trk":
            d[f"seq_{n}"] = pack2bit(lut[np.frombuffer(seqs[n].encode(), dtype=np.uint8)])
    for bi, rb in enumerate(batches):
        d[f"in{bi}_paf"] = np.frombuffer(rb.paf_text.encode(), dtype=np.uint8)
        rids = list(rb.seqs.keys())
        d[f"in{bi}_rids"] = np.array(rids)
        # reads are pure ACGT here: 2 bits per base + lengths
        d[f"in{bi}_read_len"] = np.array([len(rb.seqs[r]) for r in rids], dtype=np.int64)
        d[f"in{bi}_reads2bit"] = pack2bit(lut[np.frombuffer("".join(rb.seqs[r] for r in rids).encode(), dtype=np.uint8)])
        if spec["nb"]:
            d[f"in{bi}_barcodes"] = np.array([rb.barcodes[r] for r in rids], dtype=np.int32)
    return d


def make_kats():
    """Known-answer values of the reference's own unit tests for this path, re-derived by
    running the reference here, plus a slice of its real-data fixture pushed through `convert_records`."""
    import boss.runs.sequences as brs
    from boss.paf import Paf
    from boss.readlengthdist import ReadlengthDist

    out = {}
    for ploidy in (1, 2):
        s = brs.Scoring(ploidy=ploidy)
        out[f"p{ploidy}_score0"] = s.score0.copy()
        out[f"p{ploidy}_ent0"] = s.ent0.copy()
        out[f"p{ploidy}_phi"] = s.priors.phi.copy()
        out[f"p{ploidy}_priors"] = s.priors.priors.copy()
        out[f"p{ploidy}_phi_pow30"] = s.priors.phi_stored[:, :, :30].copy()
        # a spread of count patterns (sum <= 29), scored by the reference for every reference base
        rng = np.random.default_rng(7 + ploidy)
        pats = [np.zeros(5, dtype=np.uint16)]
        for tot in list(range(1, 30)) * 12:
            pats.append(rng.multinomial(tot, [0.8, 0.05, 0.05, 0.05, 0.05]).astype(np.uint16)[rng.permutation(5)])
        pats.append(np.array([2, 0, 0, 0, 0], dtype=np.uint16))
        pats.append(np.array([28, 0, 0, 0, 0], dtype=np.uint16))
        pats = np.array(pats, dtype=np.uint16)
        en, sc = s.calc_posterior_and_scores(cov_patterns=pats.copy())
        out[f"p{ploidy}_patterns"] = pats
        out[f"p{ploidy}_pattern_scores"] = sc
        out[f"p{ploidy}_pattern_entropies"] = en
    # test_runs_sequences.py:118-125 — values of the pre-filled table
    s = brs.Scoring(ploidy=1)
    s.init_score_array()
    out["score_arr_2_0_0_0_0_3"] = np.float64(s.score_arr[2, 0, 0, 0, 0, 3])
    out["entropy_arr_2_0_0_0_0_3"] = np.float64(s.entropy_arr[2, 0, 0, 0, 0, 3])
    out["score_arr_n_prefilled"] = np.int64(np.count_nonzero(s.score_arr[..., 0]))
    # read-length staircase defaults (test_readlengthdist.py:28-31)
    out["default_approx_ccl"] = ReadlengthDist().approx_ccl.copy()
    # real-data slice: first 120 primary records of the reference's PAF fixture through convert_records
    data = REFERENCE / "data" / "BOSS_test_data"
    paf_lines = []
    want = set()
    with open(data / "ERR3152366_10k.paf") as fh:
        for line in fh:
            if "tp:A:P" in line and len(want) < 120:
                want.add(line.split("\t", 1)[0])
                paf_lines.append(line)
            elif line.split("\t", 1)[0] in want:
                paf_lines.append(line)
    reads = {}
    with open(data / "ERR3152366_10k.fq") as fh:
        while True:
            head = fh.readline()
            if not head:
                break
            seq = fh.readline().strip()
            fh.readline(); fh.readline()
            rid = head[1:].split()[0]
            if rid in want:
                reads[rid] = seq
    paf_text = "".join(paf_lines)
    paf_dict = Paf.parse_PAF(io.StringIO(paf_text))
    cc = brs.CoverageConverter()
    inc = cc.convert_records(paf_dict=paf_dict, seqs=reads, quals={r: "5" * len(s_) for r, s_ in reads.items()})
    rows = []
    qcat = []
    for tname, lst in inc.items():
        for (start, end, q, add, bc) in lst:
            rows.append((tname, start, end, sha(q)))
            qcat.append(q)
    out["real_paf"] = np.frombuffer(paf_text.encode(), dtype=np.uint8)
    rids = list(reads.keys())
    out["real_rids"] = np.array(rids)
    out["real_reads"] = np.frombuffer("\n".join(reads[r] for r in rids).encode(), dtype=np.uint8)
    out["real_inc_tname"] = np.array([r[0] for r in rows])
    out["real_inc_start"] = np.array([r[1] for r in rows], dtype=np.int64)
    out["real_inc_end"] = np.array([r[2] for r in rows], dtype=np.int64)
    out["real_inc_sha"] = np.array([r[3] for r in rows])
    out["real_inc_query_concat"] = np.concatenate(qcat).astype(np.uint8)
    np.savez_compressed(GOLDEN / "kats.npz", **out)
    print("kats.npz", {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:6]})


def real_case_inputs(n_batches=3, reads_per_batch=700):
    """BASELINE config 1 on the reference's own data (data/BOSS_test_data): the two zymo contigs most of the
    ERR3152366 reads map to (+ one contig under 100 kb that the loader drops), and the real reads with their real
    minimap2 records (PAF file order, primary and secondary lines alike; some map to contigs outside this reduced
    reference, as in a real run with `reject`-free references)."""
    data = REFERENCE / "data" / "BOSS_test_data"
    want = {"NZ_CP041014.1": "trk", "NZ_VFAG01000001.1": "trk", "NZ_VFAF01000001.1": "short"}
    seqs, name, buf = {}, None, []
    for line in open(data / "zymo.fa"):
        if line.startswith(">"):
            if name in want:
                seqs[name] = "".join(buf)
            name, buf = line[1:].split()[0], []
        else:
            buf.append(line.strip())
    if name in want:
        seqs[name] = "".join(buf)
    names = [n for n in want if n in seqs]
    kinds = {n: want[n] for n in names}
    order, lines = [], {}
    with open(data / "ERR3152366_10k.paf") as fh:
        for line in fh:
            rid = line.split("\t", 1)[0]
            if rid not in lines:
                if len(order) >= n_batches * reads_per_batch:
                    continue
                order.append(rid)
                lines[rid] = []
            lines[rid].append(line)
    reads = {}
    need = set(order)
    with open(data / "ERR3152366_10k.fq") as fh:
        while True:
            head = fh.readline()
            if not head:
                break
            seq = fh.readline().strip()
            fh.readline(); fh.readline()
            rid = head[1:].split()[0]
            if rid in need:
                reads[rid] = seq
    batches = []
    for b in range(n_batches):
        rids = order[b * reads_per_batch: (b + 1) * reads_per_batch]
        batches.append(synth.ReadBatch("".join("".join(lines[r]) for r in rids), {r: reads[r] for r in rids}, {}, 0))
    spec = dict(ploidy=1, nb=0, bucket_threshold=0, tracked=[len(seqs[n]) for n in names if kinds[n] == "trk"], rejected=[],
                dropped=[len(seqs[n]) for n in names if kinds[n] == "short"], n_batches=n_batches, reads=reads_per_batch, seed=0,
                full_dump=False)
    return spec, names, seqs, kinds, batches


def main(which=None):
    GOLDEN.mkdir(parents=True, exist_ok=True)
    import logging
    logging.disable(logging.CRITICAL)
    if which is None or "kats" in which:
        make_kats()
    cases = dict(CASES)
    cases["real_zymo"] = None
    for name, spec in cases.items():
        if which is not None and name not in which:
            continue
        if name == "real_zymo":
            spec, names, seqs, kinds, batches = real_case_inputs()
        else:
            names, seqs, kinds = build_inputs(spec)
            batches = make_batches(spec, names, seqs, kinds)
        with tempfile.TemporaryDirectory() as td:
            ref_out = run_reference(spec, names, seqs, kinds, batches, Path(td))
        d = pack_inputs(spec, names, seqs, kinds, batches)
        d.update({f"ref_{k}": v for k, v in ref_out.items()})
        path = GOLDEN / f"case_{name}.npz"
        np.savez_compressed(path, **d)
        upd = [bool(ref_out[f"b{i}_updated"]) for i in range(spec["n_batches"])]
        print(f"{path.name}: {path.stat().st_size / 1e6:.2f} MB, strategy updated per batch: {upd}")


if __name__ == "__main__":
    main(sys.argv[1:] or None)
