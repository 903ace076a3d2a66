"""TEST INFRASTRUCTURE — one rank of a sharded run (launched by torch.distributed.run from the tests).

    torchrun --nproc-per-node 2 tests/dist_worker.py --backend gloo --case hap_nb1      # CPU: NumPy shard model
    torchrun --nproc-per-node 2 tests/dist_worker.py --backend nccl --case hap_nb1      # GPU box: libbossgpu

Every rank builds the same `ShardedRun` and feeds it the golden case's batches; rank 0 also runs the unsharded
oracle and checks threshold and every contig's strategy after every batch, then prints `SHARDED-OK`.
"""
import argparse
import os
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
for p in (str(REPO), str(REPO / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import helpers as H  # noqa: E402
import tolerances as tol  # noqa: E402
from golden_io import load_case  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--case", default="hap_nb1")
    ap.add_argument("--halo", type=int, default=160)
    ap.add_argument("--exchange", default="auto")
    ap.add_argument("--text", action="store_true", help="feed the raw PAF text (process_batch_text) instead of parsed records")
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from boss_runs_b200.engine import Engine as factory
    else:
        dist.init_process_group("gloo")
        from shard_model import NumpyShardEngine as factory
    from boss_runs_b200.sharding import ShardedRun
    rank = dist.get_rank()
    g = load_case(a.case)
    run = ShardedRun(contigs=g.records, ploidy=g.ploidy, barcodes=g.barcodes,
                     reject_refs=",".join(g.reject_refs) if g.reject_refs else None, bucket_threshold=g.bucket_threshold,
                     halo_bins=a.halo, engine_factory=factory, device=local, exchange=a.exchange, fabric_timeout_s=20.0)
    if a.exchange != "auto":
        assert run.exchange_mode == a.exchange, run.exchange_mode
    orc = H.oracle_run(g.records, g.ploidy, g.reject_refs, g.barcodes, g.bucket_threshold) if rank == 0 else None
    n_upd = 0
    for bi, (paf, seqs, bcs) in enumerate(g.batches):
        pd = H.parse_batch(paf, bcs, g.barcodes is not None)
        if a.text:
            run.rl_dist.update({rid: recs[0].qlen for rid, recs in pd.items()})
            run.process_batch_text(paf, seqs, barcodes=bcs if g.barcodes is not None else None, min_len=1)
            updated = bool(run.last.switched_on)
        else:
            updated = H.product_step(run, pd, seqs)
        if rank != 0:
            continue
        assert H.oracle_step(orc, pd, seqs) == updated, f"b{bi}: switched_on differs"
        n_upd += updated
        if updated:
            assert abs(run.threshold - orc.threshold) <= tol.THRESHOLD_RTOL * abs(orc.threshold), (run.threshold, orc.threshold)
            i = 0
            for (name, pc), oc in zip(run.contigs_filt.items(), orc.contigs_filt.values()):
                n = oc.length // 100
                H.assert_masks_match(pc.strat, oc.strat, orc.benefit_adj[i: i + n], orc.threshold, f"b{bi}/{name}: strat")
                i += n
        for name, oc in orc.contigs.items():
            if oc.rej:
                assert run.contigs[name].strat.shape == (1,) and not run.contigs[name].strat.any()
    if rank == 0:
        assert n_upd >= 2
        print(f"SHARDED-OK case={a.case} world={dist.get_world_size()} exchange={run.exchange_mode} plan={[[(s.contig, s.start, s.length) for s in segs] for segs in run.plan]}")
    dist.barrier()
    run.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
