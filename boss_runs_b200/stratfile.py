"""Packed strategy file and its consumer-side lookup — optional, behind `strategy_format`.

Upstream hands strategies to readfish as `out_<name>/masks/boss.npz`: one bool array `(L//100, 2, nb)` per contig,
`zeros(1)` for reject refs, written with `np.savez` + rename (boss/runs/core.py:59-69) and re-read whole by
`BossBits._reload_masks` / `_reload_npz` whenever the file's mtime moves (boss/dynamic_readfish.py:71-110); decisions
are single-element lookups `arr[:, reverse(, b)][start // 100]` (`_check_coord`, dynamic_readfish.py:169-210).
At 3.1 Gb that is 62 MB of bytes per update for 62 Mbit of information.

`boss.bits` holds the same masks as bits, in the order the GPU keeps them (`bossgpu_get_strat_packed`: bit
`(row * 2 + strand) * nb + barcode` of every tracked contig back to back, LSB first), behind a one-line JSON header.
`StrategyBits` is the reader: `reload()` follows `_reload_masks` (mtime test, "exception" state on a bad file),
`check_coord()` follows `_check_coord` decision for decision — including its accept-on-error paths and NumPy's
wrap-around for negative rows. The npz stays the default and the contract; this file only exists when asked for.
"""
from __future__ import annotations

import json
import logging
import mmap
import os
from pathlib import Path

import numpy as np

MAGIC = b"BOSSBITS1\n"
ALIGN = 64


def write_bits(path: str | os.PathLike, contigs: list[tuple[str, int]], nb: int, packed: np.ndarray,
               rejected: list[str] | tuple = ()) -> None:
    """`contigs`: (name, strategy rows) of every tracked contig in mask order; `packed`: uint8, their bits back to back
    (LSB first), as `Engine.strat_packed()` / `np.packbits(..., bitorder="little")` give them. Written to a temporary
    name and renamed, like upstream's npz (core.py:66-69)."""
    path = Path(path)
    off, table = 0, []
    for name, rows in contigs:
        table.append({"name": name, "rows": int(rows), "bit_off": off})
        off += int(rows) * 2 * int(nb)
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    if packed.size * 8 < off:
        raise ValueError(f"{packed.size} bytes cannot hold {off} strategy bits")
    head = json.dumps({"nb": int(nb), "n_bits": off, "contigs": table, "rejected": list(rejected)}).encode() + b"\n"
    pad = (-(len(MAGIC) + len(head))) % ALIGN
    tmp = path.with_name(path.name + ".tmp")
    with open(tmp, "wb") as fh:
        fh.write(MAGIC)
        fh.write(head[:-1] + b" " * pad + b"\n")
        fh.write(packed[: (off + 7) // 8].tobytes())
    tmp.rename(path)


class StrategyBits:
    """Consumer side: what `BossBits` keeps in `self.masks`, served from the packed file (memory-mapped)."""

    def __init__(self, mask_path: str | os.PathLike, barcodes: list[str] | None = None, scale_factor: int = 100,
                 name: str = "boss.bits"):
        self.file = Path(mask_path) / name
        self.scale_factor = scale_factor
        # dynamic_readfish builds the same index from the TOML's barcode list
        self.barcodes_index = {int(bc.split("barcode")[1]): i for i, bc in enumerate(barcodes)} if barcodes else None
        self.last_mask_mtime = 0.0
        self.exception = False
        self._contigs: dict[str, tuple[int, int]] = {}
        self._rejected: set[str] = set()
        self._nb = 1
        self._bits: np.ndarray | None = None
        self._map = None

    # -- _reload_masks (dynamic_readfish.py:86-110) -----------------------------------------------------------------
    def reload(self) -> int:
        if not self.file.is_file():
            raise FileNotFoundError("No mask files present")
        mtime = self.file.stat().st_mtime
        if not mtime > self.last_mask_mtime:
            return 0
        try:
            self._load()
            self.exception = False
        except Exception as e:  # noqa: BLE001 - upstream logs and accepts everything until the next good file
            logging.error(f"Error reading strategy bits ->>> {e!r}")
            self.exception = True
        self.last_mask_mtime = mtime
        return 1

    def _load(self) -> None:
        fh = open(self.file, "rb")
        try:
            if fh.read(len(MAGIC)) != MAGIC:
                raise ValueError("not a boss.bits file")
            head = json.loads(fh.readline())
            start = fh.tell()
            m = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
        finally:
            fh.close()
        n_bytes = (int(head["n_bits"]) + 7) // 8
        if len(m) - start < n_bytes:
            raise ValueError("truncated boss.bits file")
        self._map = m
        self._bits = np.frombuffer(m, dtype=np.uint8, count=n_bytes, offset=start)
        self._nb = int(head["nb"])
        self._contigs = {c["name"]: (int(c["rows"]), int(c["bit_off"])) for c in head["contigs"]}
        self._rejected = set(head.get("rejected", ()))

    def __len__(self) -> int:
        return len(self._contigs) + len(self._rejected)

    def __contains__(self, contig: str) -> bool:
        return contig in self._contigs or contig in self._rejected

    # -- _check_coord (dynamic_readfish.py:169-210) -----------------------------------------------------------------
    def check_coord(self, contig: str, start_pos: int, reverse: bool, barcode: str | None = None) -> int:
        if self.exception:
            return 1
        if contig not in self:
            logging.warning(f"{contig} is not in mask dict")
            return 1
        if contig in self._rejected:
            return 0
        rows, off = self._contigs[contig]
        if rows == 1:
            return 0                       # upstream: any array with shape[0] == 1 is a reject ref
        try:
            if barcode is None:
                if self._nb != 1:
                    # arr[:, reverse][row] on a (rows, 2, nb) array is a row of nb values; `if d:` raises for nb > 1
                    # inside readfish; within _check_coord itself nothing is caught, the row is returned. Serve strand
                    # bit of barcode 0 only when there is a single plane.
                    raise ValueError("barcoded masks need a barcode")
                b = 0
            else:
                b = self.barcodes_index[int(barcode.split("barcode")[1])]      # KeyError / ValueError -> accept
            row = start_pos // self.scale_factor
            if row < 0:
                row += rows                # NumPy wraps negative indices
            if not 0 <= row < rows or not 0 <= b < self._nb:
                raise IndexError(row)
            bit = off + (row * 2 + int(reverse)) * self._nb + b
            return int((self._bits[bit >> 3] >> (bit & 7)) & 1)
        except Exception:  # noqa: BLE001 - upstream: `except Exception: return 1`
            return 1

    # -- the npz view, for tests and tools ----------------------------------------------------------------------------
    def as_dict(self) -> dict[str, np.ndarray]:
        out = {}
        for name, (rows, off) in self._contigs.items():
            n = rows * 2 * self._nb
            lo, hi = off >> 3, (off + n + 7) >> 3
            bits = np.unpackbits(self._bits[lo:hi], bitorder="little")[off & 7: (off & 7) + n]
            out[name] = bits.astype(bool).reshape(rows, 2, self._nb)
        for name in self._rejected:
            out[name] = np.zeros(1, dtype=bool)
        return out


def pack_strategies(strats: list[np.ndarray]) -> np.ndarray:
    """Host-side packing of bool `(rows, 2, nb)` arrays in mask order (sharded runs, tests); one GPU packs on the device
    (`Engine.strat_packed`)."""
    if not strats:
        return np.zeros(0, dtype=np.uint8)
    flat = np.concatenate([np.asarray(s, dtype=np.uint8).reshape(-1) for s in strats])
    return np.packbits(flat, bitorder="little")
