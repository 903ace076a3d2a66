"""ORACLE / TEST INFRASTRUCTURE — golden vectors for BOSS-AEONS' benefit / threshold step (second half).

Runs only in the build container (needs /root/reference). Calls the UPSTREAM functions themselves — `Benefit.init_scoring_vec`,
`Benefit.score_array`, `Benefit.calc_fragment_benefit`, `ContigPool.find_threshold` (unbound, on a stub that carries
`.sequences`), `Sequence.find_strat_m0` — on seeded pools of contigs: node coverages with assembled-looking depth profiles,
random end markers, the default read-length staircase and the one of the reference's zymo reads.

    python -m oracle.make_golden_aeons       # writes tests/golden/aeons.npz
"""
from __future__ import annotations

import os
import sys
from pathlib import Path
from types import SimpleNamespace

REPO = Path(__file__).resolve().parent.parent
REFERENCE = Path(os.environ.get("BOSS_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path[:0] = [str(REPO), str(REPO / "oracle" / "shims"), str(REFERENCE)]

import numpy as np  # noqa: E402

GOLDEN = REPO / "tests" / "golden"
CCLS = {"prior": [1167, 2729, 3903, 4918, 5866, 6808, 7797, 8912, 10321, 12713],      # test_readlengthdist.py:28-31
        "zymo": [1647, 2280, 2810, 3305, 3840, 4379, 5045, 5867, 7015, 9768]}         # test_readlengthdist.py:12-15
POOLS = {"small": dict(n_seq=12, lo=30, hi=400, seed=1, ccl="zymo", lam=4500.0, lowcov=10),
         "mixed": dict(n_seq=60, lo=30, hi=6000, seed=2, ccl="prior", lam=6000.0, lowcov=10),
         "deep": dict(n_seq=25, lo=140, hi=3000, seed=3, ccl="zymo", lam=5200.5, lowcov=4)}


def make_pool(spec):
    """Node coverages (sum of per-base depth over 100-bp nodes, as Sequence.chunk_up_coverage leaves them)."""
    rng = np.random.default_rng(spec["seed"])
    covs, ends = [], []
    for _ in range(spec["n_seq"]):
        n = int(np.exp(rng.uniform(np.log(spec["lo"]), np.log(spec["hi"]))))
        level = rng.gamma(2.0, 4.0)
        walk = np.clip(level + np.cumsum(rng.normal(0, 0.4, size=n)), 0, None)
        walk[: min(n, 12)] *= np.linspace(0.1, 1, min(n, 12))        # coverage tails off at contig ends
        walk[-min(n, 12):] *= np.linspace(1, 0.1, min(n, 12))
        if rng.random() < 0.3:
            a = int(rng.integers(0, n))
            walk[a: a + int(rng.integers(3, 40))] = 0                 # a gap
        covs.append(np.floor(walk * 100).astype(np.float64))
        ends.append((bool(rng.random() < 0.6), bool(rng.random() < 0.6)))
    return covs, ends


def main():
    from boss.aeons.sequences import Benefit, ContigPool, Sequence
    out = {}
    for name, spec in POOLS.items():
        covs, ends = make_pool(spec)
        ccl = np.array(CCLS[spec["ccl"]])
        vec = Benefit.init_scoring_vec(lowcov=spec["lowcov"])
        seqs = {}
        for i, (cov, (e1, e2)) in enumerate(zip(covs, ends)):
            sc = Benefit.score_array(score_vec=vec, cov_arr=np.minimum(cov, 100 * 100), node_size=100)
            if e1:
                sc[0] = 1                                            # set_contig_ends (sequences.py:384-395)
            if e2:
                sc[-1] = 1
            b, ss = Benefit.calc_fragment_benefit(scores=sc, mu=400, approx_ccl=ccl, e1=e1, e2=e2, node_size=100)
            seqs[f"s{i}"] = SimpleNamespace(benefit=b, smu_sum=ss)
            out[f"{name}_{i}_scores"] = sc
            out[f"{name}_{i}_benefit"] = b
            out[f"{name}_{i}_smu_sum"] = np.float64(ss)
        stub = SimpleNamespace(sequences=seqs)
        thr = ContigPool.find_threshold(stub, mu=400, lam=spec["lam"], node_size=100)
        flat = np.column_stack([s.benefit for s in seqs.values()]).ravel()
        bins, counts = Benefit.benefit_bins(flat)
        out[f"{name}_ends"] = np.array(ends, dtype=np.uint8)
        out[f"{name}_ccl"] = ccl
        out[f"{name}_lam"] = np.float64(spec["lam"])
        out[f"{name}_threshold"] = np.float64(thr)
        out[f"{name}_bins"], out[f"{name}_counts"] = bins, counts
        for i, s in enumerate(seqs.values()):
            out[f"{name}_{i}_strat"] = np.packbits(Sequence.find_strat_m0(s, threshold=thr).ravel())
        acc = np.mean([np.mean(s.benefit >= thr) for s in seqs.values()])
        print(name, "contigs", len(seqs), "nodes", int(sum(len(c) for c in covs)), "threshold", thr, "bins", len(bins), "accept fraction", acc)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    path = GOLDEN / "aeons.npz"
    np.savez_compressed(path, **out)
    print(path.name, f"{path.stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
