"""GPU parity of the sharded path: the same CUDA kernels split into phases with halos and exchanges, (a) as
"virtual shards" — N handles on one GPU exchanging through `LocalGroup` — against the oracle on the golden
cases, state by state; (b) on two GPUs over NCCL when the box has them."""
import numpy as np
import pytest

import helpers as H
from golden_io import load_case
from test_sharding import _torchrun

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["phases", "fabric"])
@pytest.mark.parametrize("case,n_virtual", [("hap_nb1", 2), ("hap_nb1", 3), ("dip_nb1", 4), ("hap_nb3", 3), ("dip_nb2", 2), ("hap_pad", 2), ("real_zymo", 3)])
def test_virtual_shards_match_oracle(case, n_virtual, exchange, lib):
    """`phases`: the host runs the exchanges between the phases (copies between the handles). `fabric`: each
    virtual shard runs on its own stream and the handles exchange on the device through each other's exchange
    blocks (csrc/fabric.cuh) — the same kernels N processes use over NVLink."""
    from boss_runs_b200.sharding import ShardedRun
    g = load_case(case)
    run = ShardedRun(contigs=g.records, ploidy=g.ploidy, barcodes=g.barcodes,
                     reject_refs=",".join(g.reject_refs) if g.reject_refs else None, bucket_threshold=g.bucket_threshold,
                     n_virtual=n_virtual, halo_bins=160, write_debug=True, exchange=exchange, fabric_timeout_s=20.0)
    assert run.exchange_mode == exchange
    if n_virtual > len(run.contigs_filt):
        assert any(s.start > 0 for segs in run.plan for s in segs)
    orc = H.oracle_run(g.records, g.ploidy, g.reject_refs, g.barcodes, g.bucket_threshold)
    for bi, (paf, seqs, bcs) in enumerate(g.batches):
        pd = H.parse_batch(paf, bcs, g.barcodes is not None)
        upd_o = H.oracle_step(orc, pd, seqs)
        upd_p = H.product_step(run, pd, seqs)
        assert upd_o == upd_p
        H.compare_state(run, orc, upd_p, f"{case}/v{n_virtual}/b{bi}")


def test_sharded_equals_unsharded_bitwise(lib):
    """Integer-limb histogram + direct window sums: the sharded result is bit-identical to the one-handle result
    (threshold, benefit, masks), not merely within tolerance."""
    from boss_runs_b200.sharding import ShardedRun
    g = load_case("dip_nb1")
    kw = dict(contigs=g.records, ploidy=g.ploidy, barcodes=g.barcodes, reject_refs=None, bucket_threshold=g.bucket_threshold)
    one = H.product_run(g.records, g.ploidy, g.reject_refs, g.barcodes, g.bucket_threshold)
    runs = [ShardedRun(n_virtual=3, halo_bins=160, write_debug=True, **kw),
            ShardedRun(n_virtual=4, halo_bins=160, write_debug=True, exchange="fabric", fabric_timeout_s=20.0, **kw)]
    for bi, (paf, seqs, bcs) in enumerate(g.batches):
        pd = H.parse_batch(paf, bcs, False)
        H.product_step(one, pd, seqs)
        for many in runs:
            H.product_step(many, pd, seqs)
            assert one.last.switched_on == many.last.switched_on
            if one.last.switched_on:
                assert one.threshold == many.threshold and one.last.ubar0 == many.last.ubar0
                assert one.last.n_nonzero == many.last.n_nonzero and one.last.n_dropout == many.last.n_dropout
            for (name, a), b in zip(one.contigs_filt.items(), many.contigs_filt.values()):
                assert np.array_equal(a.coverage, b.coverage)
                assert np.array_equal(a.strat, b.strat), f"b{bi}/{name}/{many.exchange_mode}"
                if one.last.switched_on:
                    assert np.array_equal(a.additional_benefit, b.additional_benefit), f"b{bi}/{name}/{many.exchange_mode}"


def test_fabric_peer_timeout_is_an_error_not_a_hang(lib):
    """A shard whose peer never arrives gives up after the timeout and reports BOSSGPU_EPEER."""
    from boss_runs_b200._lib import PeerTimeout
    from boss_runs_b200.sharding import ShardedRun
    g = load_case("hap_nb1")
    run = ShardedRun(contigs=g.records, ploidy=g.ploidy, barcodes=g.barcodes, reject_refs=None,
                     bucket_threshold=g.bucket_threshold, n_virtual=2, halo_bins=160, exchange="fabric", fabric_timeout_s=0.2)
    e0 = run.engines[0]
    p = e0.params(run.rl_dist.approx_ccl, 500.0, run.bucket_threshold, fhat_scalars=run.read_starts.pointmass_scalars())
    e0.update_fused_begin(p)                 # shard 1 never starts its update
    with pytest.raises(PeerTimeout):
        e0.update_fused_end()


@pytest.mark.parametrize("exchange", ["phases", "fabric"])
@pytest.mark.parametrize("case", ["hap_nb1", "dip_nb2"])
def test_two_gpus_nccl(case, exchange, lib):
    """Two processes, one GPU each: `phases` exchanges with NCCL collectives, `fabric` over CUDA-IPC-mapped peer
    memory inside the update's own kernels."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = _torchrun(2, "--backend", "nccl", "--case", case, "--exchange", exchange)
    assert r.returncode == 0 and "SHARDED-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]
