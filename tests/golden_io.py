"""Decode tests/golden/case_*.npz (written by oracle/make_golden.py): inputs + the reference's outputs."""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ("hap_nb1", "dip_nb1", "hap_nb3", "dip_nb2", "hap_pad", "real_zymo")
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def unpack2bit(packed: np.ndarray, n: int) -> np.ndarray:
    bits = np.unpackbits(packed)[: 2 * n].reshape(n, 2)
    return (bits[:, 0] << 1 | bits[:, 1]).astype(np.uint8)


@dataclass
class GoldenCase:
    name: str
    z: dict
    records: list            # [(name, seq)] in FASTA order, incl. rejected + dropped contigs
    reject_refs: list
    barcodes: list | None
    ploidy: int
    bucket_threshold: float
    batches: list            # [(paf_text, {rid: read}, {rid: barcode index})]

    def ref(self, key):
        return self.z["ref_" + key]

    def has(self, key) -> bool:
        return ("ref_" + key) in self.z


def load_case(name: str) -> GoldenCase:
    z = dict(np.load(GOLDEN / f"case_{name}.npz", allow_pickle=False))
    names = [str(x) for x in z["names"]]
    kinds = [str(x) for x in z["kinds"]]
    lengths = z["lengths"]
    records = []
    for n, k, L in zip(names, kinds, lengths):
        if k == "trk":
            seq = _ACGT[unpack2bit(z[f"seq_{n}"], int(L))].tobytes().decode()
        else:
            seq = "ACGT" * (int(L) // 4) + "A" * (int(L) % 4)
        records.append((n, seq))
    nb = int(z["nb"])
    batches = []
    for bi in range(int(z["n_batches"])):
        paf = z[f"in{bi}_paf"].tobytes().decode()
        rids = [str(x) for x in z[f"in{bi}_rids"]]
        lens = z[f"in{bi}_read_len"]
        allb = _ACGT[unpack2bit(z[f"in{bi}_reads2bit"], int(lens.sum()))].tobytes().decode()
        off = np.concatenate(([0], np.cumsum(lens)))
        seqs = {r: allb[off[i]: off[i + 1]] for i, r in enumerate(rids)}
        bcs = {r: int(b) for r, b in zip(rids, z[f"in{bi}_barcodes"])} if nb else {}
        batches.append((paf, seqs, bcs))
    return GoldenCase(name, z, records, [n for n, k in zip(names, kinds) if k == "rej"],
                      [f"barcode{i + 1:02d}" for i in range(nb)] if nb else None, int(z["ploidy"]),
                      float(z["bucket_threshold"]), batches)
