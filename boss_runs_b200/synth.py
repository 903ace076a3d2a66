"""Synthetic references and read batches of the shapes BASELINE.json names (tests and bench.py).

The generator emits what the hot path receives upstream: PAF text with `cg:Z:` CIGARs (as
`minimap2 -x map-ont --secondary=no -c` writes them) plus the read strings, so the same inputs can be fed
to this package, to the oracle and — in the build container — to the reference itself.

Read model: length ~ clip(Gamma(4, 2500), 1000, 60000), uniform start, 50 %
reverse strand, 4 % substitutions, an indel (50/50 insertion/deletion, 1-2 bp) after every U[5,40) aligned
bases, alignments begin and end with a match run, unaligned flanks of 0-30 bases on the read.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP_CODE = np.array([3, 2, 1, 0], dtype=np.uint8)


def random_contigs(lengths: dict[str, int] | list[int], seed: int = 7) -> dict[str, str]:
    """{name: sequence} with iid uniform bases."""
    rng = np.random.default_rng(seed)
    if not isinstance(lengths, dict):
        lengths = {f"ctg{i + 1}": int(n) for i, n in enumerate(lengths)}
    return {name: ACGT[rng.integers(0, 4, size=int(n), dtype=np.uint8)].tobytes().decode() for name, n in lengths.items()}


def grch38_like_lengths(total: int = 3_100_000_000, n: int = 25) -> list[int]:
    """`n` contig lengths with GRCh38-like proportions (chr1..22, X, Y + one small), summing to ~total."""
    mb = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28, 114.36,
          107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23, 0.5]
    mb = np.array((mb * ((n + len(mb) - 1) // len(mb)))[:n])
    out = np.maximum((mb / mb.sum() * total).astype(np.int64), 100_000)
    return [int(x) for x in out]


@dataclass
class ReadBatch:
    """One batch in the three forms the tests need."""
    paf_text: str
    seqs: dict[str, str]
    barcodes: dict[str, int] = field(default_factory=dict)
    n_ref_positions: int = 0


def _one_alignment(rng, ref_codes: np.ndarray, t0: int, span: int, sub_rate: float, indels: bool = True):
    """Alignment of `span` reference positions starting at t0 -> (ops [(len, 'M'|'I'|'D')], read codes in
    reference orientation)."""
    if not indels:
        q = ref_codes[t0:t0 + span].copy()
        sub = rng.random(span) < sub_rate
        q[sub] = (q[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
        return [(span, "M")], q
    k = span // 5 + 2
    m_len = rng.integers(5, 40, size=k)
    is_del = rng.random(k) < 0.5
    gap = rng.integers(1, 3, size=k)
    ref_use = m_len + np.where(is_del, gap, 0)
    cum = np.cumsum(ref_use)
    last = int(np.searchsorted(cum, span - 5, side="left"))      # block whose match run closes the alignment
    before = int(cum[last - 1]) if last > 0 else 0
    m_len = m_len[: last + 1].copy()
    m_len[last] = span - before                                   # final run: matches only, no trailing indel
    is_del, gap = is_del[:last], gap[:last]
    ops = []
    for i in range(last):
        ops.append((int(m_len[i]), "M"))
        ops.append((int(gap[i]), "D" if is_del[i] else "I"))
    ops.append((int(m_len[last]), "M"))
    # read bases: per block `m_len` matches followed by `gap` inserted bases (insertions only)
    ins_len = np.where(is_del, 0, gap)
    del_len = np.where(is_del, gap, 0)
    q_block = m_len.copy()
    q_block[:last] += ins_len
    q_total = int(q_block.sum())
    blk = np.repeat(np.arange(last + 1), q_block)
    q_start_of_blk = np.concatenate(([0], np.cumsum(q_block)[:-1]))
    within = np.arange(q_total) - q_start_of_blk[blk]
    is_ins = within >= m_len[blk]
    ref_start_of_blk = t0 + np.concatenate(([0], np.cumsum(m_len[:last] + del_len)))
    ref_pos = ref_start_of_blk[blk] + np.minimum(within, m_len[blk] - 1)
    q = ref_codes[ref_pos].copy()
    sub = (rng.random(q_total) < sub_rate) & ~is_ins
    q[sub] = (q[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    q[is_ins] = rng.integers(0, 4, size=int(is_ins.sum()), dtype=np.uint8)
    return ops, q


def read_batch(contigs: dict[str, str], n_reads: int, seed: int, mean_len: float = 10_000.0, min_len: int = 1000,
               max_len: int = 60_000, sub_rate: float = 0.04, n_barcodes: int = 0, focus: tuple | None = None,
               prefix: str = "r", codes: dict[str, np.ndarray] | None = None, weird_base_every: int = 0) -> ReadBatch:
    """`n_reads` reads over `contigs` (chosen proportionally to length).

    :param n_barcodes: > 0 tags every read with a barcode index in [0, n_barcodes) (a few fall to 0, like
                       upstream's "unclassified", Q11)
    :param focus: optional (contig name, start, end, fraction): that fraction of reads starts inside the
                  window (drives local depth past the freeze threshold of 30 in small tests)
    :param weird_base_every: > 0 injects a non-ACGT character into every k-th read (error-path tests)
    """
    rng = np.random.default_rng(seed)
    names = list(contigs.keys())
    if codes is None:
        lut = np.zeros(256, dtype=np.uint8)
        lut[ACGT] = np.arange(4, dtype=np.uint8)
        # contigs may be strings or already-integerised uint8 arrays
        codes = {n: (lut[np.frombuffer(contigs[n].encode(), dtype=np.uint8)] if isinstance(contigs[n], str)
                     else np.asarray(contigs[n], dtype=np.uint8)) for n in names}
    lens = np.array([len(contigs[n]) for n in names], dtype=np.float64)
    which = rng.choice(len(names), size=n_reads, p=lens / lens.sum())
    lines, seqs, bcs = [], {}, {}
    total_ref = 0
    for i in range(n_reads):
        name = names[int(which[i])]
        L = len(contigs[name])
        span = int(np.clip(rng.gamma(4.0, mean_len / 4.0), min_len, min(max_len, L - 1)))
        if focus is not None and rng.random() < focus[3]:
            name = focus[0]
            L = len(contigs[name])
            span = min(span, L - 1)
            t0 = int(rng.integers(max(focus[1] - span // 2, 0), max(min(focus[2], L - span), 1)))
        else:
            t0 = int(rng.integers(0, L - span))
        rev = bool(rng.random() < 0.5)
        ops, q = _one_alignment(rng, codes[name], t0, span, sub_rate)
        if rev:
            q = _COMP_CODE[q[::-1]]                       # the read as sequenced
        lf, rf = int(rng.integers(0, 31)), int(rng.integers(0, 31))
        flank_l = rng.integers(0, 4, size=lf, dtype=np.uint8)
        flank_r = rng.integers(0, 4, size=rf, dtype=np.uint8)
        read_codes = np.concatenate((flank_l, q, flank_r))
        read = ACGT[read_codes].tobytes().decode()
        if weird_base_every and i % weird_base_every == weird_base_every - 1:
            mid = lf + len(q) // 2
            read = read[:mid] + "N" + read[mid + 1:]
        qlen = len(read)
        rid = f"{prefix}{seed}_{i}"
        cigar = "".join(f"{n}{o}" for n, o in ops)
        matches = sum(n for n, o in ops if o == "M")
        block = sum(n for n, _ in ops)
        lines.append("\t".join(map(str, (rid, qlen, lf, lf + len(q), "-" if rev else "+", name, L, t0, t0 + span, matches,
                                         block, 60, f"AS:i:{matches}", "tp:A:P", f"s1:i:{matches // 2}", f"cg:Z:{cigar}"))))
        seqs[rid] = read
        total_ref += span
        if n_barcodes > 0:
            bcs[rid] = int(rng.integers(0, n_barcodes)) if rng.random() > 0.05 else 0
    return ReadBatch("\n".join(lines) + "\n", seqs, bcs, total_ref)


def packed_batch(contig_lengths, n_reads: int, seed: int, mean_len: float = 10_000.0, n_barcodes: int = 1,
                 indel_every: int = 22) -> dict:
    """A large batch directly in libbossgpu's packed form (no text), fully vectorised: every read is
    [M(indel_every) D(1) M(indel_every) I(1)]* M(rest) with random bases. For scatter-throughput runs where
    text generation would dominate; the parity tests use `read_batch`.

    Returns dict(seg, tstart, barcode, cig_off, cigar, base_off, bases, n_ref_positions)."""
    rng = np.random.default_rng(seed)
    lens = np.asarray(contig_lengths, dtype=np.int64)
    seg = rng.choice(len(lens), size=n_reads, p=lens / lens.sum()).astype(np.int32)
    span = np.clip(rng.gamma(4.0, mean_len / 4.0, size=n_reads), 1000, 60_000).astype(np.int64)
    span = np.minimum(span, lens[seg] - 1)
    tstart = (rng.random(n_reads) * (lens[seg] - span)).astype(np.int64)
    n_blk = span // (indel_every + 1)                 # blocks of indel_every matches + one indel
    n_del = (n_blk + 1) // 2                          # even blocks delete, odd blocks insert
    n_ins = n_blk // 2
    last_m = span - (n_blk * indel_every + n_del)     # closing match run keeps the reference span exact
    n_ops = 2 * n_blk + 1
    cig_off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(n_ops, out=cig_off[1:])
    total_ops = int(cig_off[-1])
    rd = np.repeat(np.arange(n_reads), n_ops)
    pos = np.arange(total_ops) - cig_off[:-1][rd]
    is_last = pos == n_ops[rd] - 1
    is_m = pos % 2 == 0
    cls = np.where(is_m, 0, np.where((pos // 2) % 2 == 0, 2, 1)).astype(np.uint32)
    ln = np.where(is_last, last_m[rd], np.where(is_m, indel_every, 1)).astype(np.uint32)
    cigar = (ln << 4) | cls
    q_len = n_blk * indel_every + n_ins + last_m
    base_off = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(q_len, out=base_off[1:])
    bases = rng.integers(0, 4, size=int(base_off[-1]), dtype=np.uint8)
    barcode = (rng.integers(0, n_barcodes, size=n_reads).astype(np.int32) if n_barcodes > 1
               else np.zeros(n_reads, dtype=np.int32))
    return dict(seg=seg, tstart=tstart, barcode=barcode, cig_off=cig_off, cigar=cigar, base_off=base_off, bases=bases,
                n_ref_positions=int(span.sum()))
