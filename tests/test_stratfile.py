"""Packed strategy file: `boss.bits` holds exactly what `boss.npz` holds, and `StrategyBits.check_coord`
takes every decision `BossBits._check_coord` (boss/dynamic_readfish.py:169-210) takes on the npz arrays — restated
below on NumPy arrays the way upstream indexes them. CPU only."""
import logging
import os
import time

import numpy as np
import pytest

from boss_runs_b200 import stratfile


def upstream_check_coord(masks, contig, start_pos, reverse, barcode, barcodes_index, scale_factor=100):
    """dynamic_readfish.py:190-210, verbatim logic."""
    if masks.get("exception", False):
        return 1
    if contig not in masks:
        return 1
    arr = masks.get(contig)
    if arr.shape[0] == 1:
        return 0
    try:
        if barcode is None:
            d = arr[:, int(reverse)][start_pos // scale_factor]
        else:
            b = barcodes_index[int(barcode.split("barcode")[1])]
            d = arr[:, int(reverse), b][start_pos // scale_factor]
        return d
    except Exception:  # noqa
        return 1


def make_masks(rng, nb):
    rows = {"ctgA": 1300, "ctgB": 1001, "ctgC": 7, "tiny": 1}
    masks = {n: rng.random((r, 2, nb)) < 0.4 for n, r in rows.items()}
    masks["rejected1"] = np.zeros(1, dtype=bool)
    return masks


@pytest.mark.parametrize("nb", [1, 3])
def test_bits_file_equals_npz_content_and_decisions(tmp_path, nb):
    rng = np.random.default_rng(nb)
    masks = make_masks(rng, nb)
    tracked = [(n, a.shape[0]) for n, a in masks.items() if a.ndim == 3]
    packed = stratfile.pack_strategies([masks[n] for n, _ in tracked])
    stratfile.write_bits(tmp_path / "boss.bits", tracked, nb, packed, rejected=["rejected1"])
    barcodes = [f"barcode{i + 1:02d}" for i in range(nb)] if nb > 1 else None
    sb = stratfile.StrategyBits(tmp_path, barcodes=barcodes)
    assert sb.reload() == 1 and sb.reload() == 0                    # second call: mtime unchanged (dynamic_readfish.py:97-98)
    got = sb.as_dict()
    assert set(got) == set(masks)
    for n, a in masks.items():
        assert np.array_equal(got[n], a), n
    bidx = sb.barcodes_index
    names = list(masks) + ["unknown_contig"]
    bcs = [None] if nb == 1 else ["barcode01", "barcode03", "barcode07", "unclassified"]
    for _ in range(4000):
        contig = names[int(rng.integers(len(names)))]
        pos = int(rng.integers(-140_000, 140_000))
        rev = bool(rng.integers(2))
        bc = bcs[int(rng.integers(len(bcs)))]
        want = upstream_check_coord(masks, contig, pos, rev, bc, bidx)
        have = sb.check_coord(contig, pos, rev, bc)
        assert int(np.asarray(want).reshape(-1)[0]) == have, (contig, pos, rev, bc)


def test_bits_reload_follows_upstream_states(tmp_path, caplog):
    sb = stratfile.StrategyBits(tmp_path)
    with pytest.raises(FileNotFoundError):
        sb.reload()                                                 # "No mask files present"
    (tmp_path / "boss.bits").write_bytes(b"garbage")
    with caplog.at_level(logging.ERROR):
        assert sb.reload() == 1
    assert sb.exception and sb.check_coord("anything", 5, False) == 1   # masks = {"exception": True}: accept everything
    m = {"c": np.ones((12, 2, 1), dtype=bool)}
    m["c"][3, 1, 0] = False
    stratfile.write_bits(tmp_path / "boss.bits", [("c", 12)], 1, stratfile.pack_strategies([m["c"]]))
    os.utime(tmp_path / "boss.bits", (time.time() + 5, time.time() + 5))
    assert sb.reload() == 1 and not sb.exception
    assert sb.check_coord("c", 399, True) == 0 and sb.check_coord("c", 399, False) == 1 and sb.check_coord("c", 1200, True) == 1
    assert not (tmp_path / "boss.bits.tmp").exists()
    with pytest.raises(ValueError):
        stratfile.write_bits(tmp_path / "x.bits", [("c", 12)], 1, np.zeros(1, np.uint8))
