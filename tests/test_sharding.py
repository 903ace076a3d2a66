"""Host side of the multi-GPU path on CPU: shard planning, and the whole sharded update protocol
(boss_runs_b200/sharding.py) under torch.distributed's gloo backend with world_size 2, each rank holding a
NumPy model of its shard (tests/shard_model.py), checked against the unsharded oracle."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from boss_runs_b200 import synth
from boss_runs_b200._lib import BIN, BUCKET
from boss_runs_b200.sharding import merged_row_starts, plan_shards

REPO = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("n", [1, 2, 3, 4, 8])
def test_plan_covers_the_genome_once(n):
    for lens in (synth.grch38_like_lengths(), [130_050, 150_000, 100_000, 2_345_678], [20_000 * 9 + 17]):
        plan = plan_shards(lens, n)
        assert len(plan) == n and all(plan)
        pos = {k: 0 for k in range(len(lens))}
        order = []
        for segs in plan:
            for s in segs:
                assert s.start == pos[s.contig] and s.length > 0          # contiguous, in order, no overlap
                assert s.start % BUCKET == 0
                end = s.start + s.length
                assert end == lens[s.contig] or end % BUCKET == 0
                if end == lens[s.contig]:
                    assert lens[s.contig] // BUCKET - s.start // BUCKET >= 1   # tail keeps a complete bucket
                pos[s.contig] = end
                order.append(s.contig)
            assert len({s.contig for s in segs}) == len(segs)              # one segment per contig in a shard
        assert order == sorted(order) and pos == {k: L for k, L in enumerate(lens)}
        sizes = [sum(s.length for s in segs) for segs in plan]
        if sum(lens) > 50 * n * BUCKET:
            assert max(sizes) - min(sizes) <= max(lens) // 2 + 2 * BUCKET or max(sizes) / (sum(lens) / n) < 1.02
        rs = merged_row_starts(lens, plan)
        assert rs[0] == 0 and rs[-1] == sum(L // BIN + 1 for L in lens) and np.all(np.diff(rs) > 0)


def test_plan_balance_at_human_scale():
    lens = synth.grch38_like_lengths()
    for n in (2, 4, 8):
        sizes = [sum(s.length for s in segs) for segs in plan_shards(lens, n)]
        assert max(sizes) / (sum(lens) / n) < 1.001


def test_plan_rejects_the_impossible():
    with pytest.raises(ValueError):
        plan_shards([100_000], 6)            # five buckets cannot feed six shards
    with pytest.raises(ValueError):
        plan_shards([10_000], 1)
    with pytest.raises(ValueError):
        plan_shards([100_000], 0)


def _torchrun(nproc, *args, timeout=600):
    port = 29500 + os.getpid() % 2000
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(REPO / "tests" / "dist_worker.py"), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=str(REPO))


@pytest.mark.parametrize("case", ["hap_nb1", "dip_nb2"])
def test_sharded_update_gloo_world2(case):
    r = _torchrun(2, "--backend", "gloo", "--case", case)
    assert r.returncode == 0 and "SHARDED-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]


def test_sharded_text_batches_gloo_world2():
    """The raw-PAF-text entry point (`process_batch_text`: C tokeniser, read starts from arrays) through the sharded run."""
    r = _torchrun(2, "--backend", "gloo", "--case", "hap_nb3", "--text")
    assert r.returncode == 0 and "SHARDED-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]


def test_virtual_shards_numpy_model():
    """The same protocol with every shard in one process (LocalGroup), three shards: exercises a shard that lies
    between two cuts and contigs split in the middle."""
    sys.path.insert(0, str(REPO / "tests"))
    import helpers as H
    import tolerances as tol
    from golden_io import load_case
    from shard_model import NumpyShardEngine
    from boss_runs_b200.sharding import ShardedRun
    g = load_case("hap_nb3")
    run = ShardedRun(contigs=g.records, ploidy=g.ploidy, barcodes=g.barcodes, reject_refs=",".join(g.reject_refs),
                     bucket_threshold=g.bucket_threshold, n_virtual=3, halo_bins=160, engine_factory=NumpyShardEngine)
    assert any(s.start > 0 for segs in run.plan for s in segs), "the plan should split a contig"
    orc = H.oracle_run(g.records, g.ploidy, g.reject_refs, g.barcodes, g.bucket_threshold)
    for bi, (paf, seqs, bcs) in enumerate(g.batches):
        pd = H.parse_batch(paf, bcs, True)
        assert H.product_step(run, pd, seqs) == H.oracle_step(orc, pd, seqs)
        for (name, pc), oc in zip(run.contigs_filt.items(), orc.contigs_filt.values()):
            assert np.array_equal(pc.coverage, oc.coverage), f"b{bi}/{name}: coverage"
            assert np.array_equal(pc.bucket_switches, oc.bucket_switches)
        if run.last.switched_on:
            assert abs(run.threshold - orc.threshold) <= tol.THRESHOLD_RTOL * orc.threshold
            i = 0
            for (name, pc), oc in zip(run.contigs_filt.items(), orc.contigs_filt.values()):
                n = oc.length // 100
                H.assert_masks_match(pc.strat, oc.strat, orc.benefit_adj[i: i + n], orc.threshold, f"b{bi}/{name}")
                i += n
