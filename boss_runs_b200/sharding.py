"""Multi-GPU strategy update: the genome axis split into contiguous shards, one per B200.

The reference is a single process; per-contig work is independent upstream
(boss/runs/core.py:83-86,95-99,107-108,119-121 are per-contig loops) and only three things couple positions:
the box-filter support of the smoothing (reference.py:233-260), the single global threshold
(sequences.py:566-649) and the row shift of `_distribute_strategy` (core.py:141,155; quirk Q2). So:

  * `plan_shards` cuts the concatenated genome into N ranges balanced by sites. Cuts fall on contig ends or,
    inside a contig, on multiples of the 20 kb bucket (so bucket sums, 100-site bins and 2 kb windows never
    straddle a cut).
  * every shard is handed the whole (small) batch; the library keeps the reads overlapping its range.
  * one update = the same kernels as on one GPU, interleaved with five small exchanges
    (include/bossgpu.h, bossgpu_update_phase):
        halo of scores_ds bins to both neighbours (only meaningful where a contig is split),
        allreduce(max) of the bucket switch, allreduce(max) of the normaliser bits,
        allreduce(sum) of the integer-limb histogram (exact, so the threshold does not depend on N),
        allgather of the packed merged mask (strategy row d reads merged row d, which may live next door).

`ShardGroup` runs those exchanges either between the shards of ONE process (`LocalGroup`: N engines on one
GPU — "virtual shards", used by the GPU parity tests and to dry-run the sharded path without N GPUs) or between
processes with one shard each (`DistGroup`: torch.distributed, NCCL over NVLink on the GPU box, gloo in the CPU
tests). torch is plumbing here; no numerics run in it.
"""
from __future__ import annotations

import logging

import numpy as np

from ._lib import BIN, BUCKET, BUF_SWITCH, BUF_NORM, BUF_HIST, BUF_MASK, BUF_HALO_SEND, BUF_HALO_RECV, BUF_STRAT, BUF_COV_TOTAL
from .engine import Engine, SegmentSpec, UpdateOutcome
from .runs import BossRuns, PackedBatch


# ------------------------------------------------------------------------------------------------------
# planning
# ------------------------------------------------------------------------------------------------------
def plan_shards(contig_lengths, n_shards: int) -> list[list[SegmentSpec]]:
    """Split the concatenated contigs (contigs_filt order) into `n_shards` contiguous, non-empty ranges of
    near-equal size. Inside a contig a cut must be a multiple of BUCKET that leaves at least one complete bucket
    on its right (the tail segment owns the contig's extra last switch, which repeats the last complete bucket's
    mean: reference.py:198-202, utils.py:215-217)."""
    lens = [int(x) for x in contig_lengths]
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    if any(L < BUCKET for L in lens):
        raise ValueError("contigs shorter than one bucket cannot be tracked")
    starts = np.concatenate(([0], np.cumsum(lens)))
    total = int(starts[-1])

    def candidates(k):
        """allowed cut offsets inside contig k, plus its two ends"""
        L = lens[k]
        last = (L // BUCKET - 1) * BUCKET
        return 0, L, (BUCKET, last) if last >= BUCKET else None

    cuts = [0]
    for i in range(1, n_shards):
        want = total * i // n_shards
        k = int(np.searchsorted(starts, want, side="right") - 1)
        k = min(k, len(lens) - 1)
        x = want - int(starts[k])
        lo, hi, inner = candidates(k)
        opts = [lo, hi]
        if inner is not None:
            c = int(round(x / BUCKET)) * BUCKET
            opts.append(min(max(c, inner[0]), inner[1]))
        best = min(opts, key=lambda o: abs(o - x))
        g = int(starts[k]) + best
        if g <= cuts[-1]:
            # nearest allowed position is already taken: move right to the next allowed one
            g = _next_cut_after(cuts[-1], lens, starts)
            if g is None:
                raise ValueError(f"cannot cut {len(lens)} contig(s) of {total} sites into {n_shards} shards")
        cuts.append(g)
    cuts.append(total)
    if any(b <= a for a, b in zip(cuts, cuts[1:])):
        raise ValueError(f"cannot cut {len(lens)} contig(s) of {total} sites into {n_shards} shards")
    plan = []
    for a, b in zip(cuts, cuts[1:]):
        segs = []
        k = int(np.searchsorted(starts, a, side="right") - 1)
        pos = a
        while pos < b:
            c0, c1 = int(starts[k]), int(starts[k + 1])
            end = min(b, c1)
            segs.append(SegmentSpec(contig=k, start=pos - c0, length=end - pos))
            pos = end
            k += 1
        plan.append(segs)
    return plan


def _next_cut_after(g, lens, starts):
    k = int(np.searchsorted(starts, g, side="right") - 1)
    while k < len(lens):
        c0, L = int(starts[k]), lens[k]
        x = g - c0
        last = (L // BUCKET - 1) * BUCKET
        c = (x // BUCKET + 1) * BUCKET
        if x < L and last >= BUCKET and max(c, BUCKET) <= last:
            return c0 + max(c, BUCKET)
        if c0 + L > g and k + 1 < len(lens):
            return c0 + L
        k += 1
    return None


def merged_row_starts(contig_lengths, plan) -> np.ndarray:
    """Global merged-row index (rows of the concatenated per-contig benefit arrays, L//100 + 1 per contig:
    reference.py:225,254) of each shard's first row, plus the total."""
    lens = np.asarray(contig_lengths, dtype=np.int64)
    o_row = np.concatenate(([0], np.cumsum(lens // BIN + 1)))
    out = [int(o_row[segs[0].contig] + segs[0].start // BIN) for segs in plan]
    out.append(int(o_row[-1]))
    return np.asarray(out, dtype=np.int64)


# ------------------------------------------------------------------------------------------------------
# exchange groups
# ------------------------------------------------------------------------------------------------------
_NP = {"i32": np.int32, "i64": np.int64, "u8": np.uint8, "f64": np.float64}


def _view(t, kind: str):
    import torch
    return t.view({"i32": torch.int32, "i64": torch.int64, "u8": torch.uint8, "f64": torch.float64}[kind])


class LocalGroup:
    """All shards live in this process (engines on one device): exchanges are copies between their buffers."""

    def __init__(self, engines):
        self.engines = list(engines)
        self.n_shards = len(self.engines)
        self.shard_ids = list(range(self.n_shards))
        self.rank, self.world = 0, 1

    def _sync(self) -> None:
        """The copies below run on torch's current stream; virtual shards may each run on their own."""
        import torch
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def allreduce(self, which: int, op: str, kind: str):
        self._sync()
        ts = [_view(e.exchange_tensor(which), kind) for e in self.engines]
        acc = ts[0].clone()
        for t in ts[1:]:
            acc = (acc.maximum(t) if op == "max" else acc + t)
        for t in ts:
            t.copy_(acc)
        self._sync()
        return acc

    def allgather_mask(self) -> None:
        ts = [e.exchange_tensor(BUF_MASK) for e in self.engines]
        stride = ts[0].numel() // self.n_shards
        for s, src in enumerate(ts):
            for d, dst in enumerate(ts):
                if d != s:
                    dst[s * stride:(s + 1) * stride].copy_(src[s * stride:(s + 1) * stride])

    def exchange_halos(self) -> None:
        send = [_view(e.exchange_tensor(BUF_HALO_SEND), "f64") for e in self.engines]
        recv = [_view(e.exchange_tensor(BUF_HALO_RECV), "f64") for e in self.engines]
        h = send[0].numel() // 2
        for s in range(self.n_shards):
            if s > 0:
                recv[s][:h].copy_(send[s - 1][h:])          # left neighbour's last bins
            if s + 1 < self.n_shards:
                recv[s][h:].copy_(send[s + 1][:h])          # right neighbour's first bins

    def shared_mirror(self, nbytes: int) -> np.ndarray:
        """Host array every shard's distribution kernel writes its slice of (here: an anonymous mapping of this process,
        page-locked once for all engines). It owns whole pages: page-locking works on pages, and a heap array would share
        its first and last page with unrelated small arrays — a later pageable copy into one of those (half inside a
        locked page, half outside) is refused by the driver with "invalid argument"."""
        import mmap
        self._mirror_map = mmap.mmap(-1, (max(nbytes, 1) + mmap.PAGESIZE - 1) // mmap.PAGESIZE * mmap.PAGESIZE)
        self._mirror = np.frombuffer(self._mirror_map, dtype=np.uint8)[:nbytes]
        self._mirror[:] = 1
        return self._mirror

    def setup_fabric(self, timeout_s: float = 0.0) -> bool:
        """Virtual shards share one address space: every engine is handed the others' exchange blocks by
        address. Each engine must run on its OWN stream (a shard spins on the device until its peers' kernels
        have run; on one stream they would queue behind the spin)."""
        if not all(hasattr(e, "fabric_info") for e in self.engines):
            return False
        ptrs = [e.fabric_info()[0] for e in self.engines]
        for e in self.engines:
            e.fabric_attach(ptrs, timeout_s)
        return True

    def agree(self, ok: bool) -> bool:
        return ok

    def barrier(self) -> None:
        pass


class DistGroup:
    """One shard per process; exchanges are torch.distributed collectives on tensors aliasing the engine's
    buffers (CUDA memory with NCCL, host memory with gloo in the CPU tests)."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.engines = [engine]
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n_shards = self.world
        self.shard_ids = [self.rank]
        self._gather_buf = None

    def allreduce(self, which: int, op: str, kind: str):
        t = _view(self.engines[0].exchange_tensor(which), kind)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM, group=self.group)
        return t

    def allgather_mask(self) -> None:
        t = self.engines[0].exchange_tensor(BUF_MASK)
        stride = t.numel() // self.world
        self.dist.all_gather_into_tensor(t, t[self.rank * stride:(self.rank + 1) * stride].clone(), group=self.group)

    def exchange_halos(self) -> None:
        dist = self.dist
        send = _view(self.engines[0].exchange_tensor(BUF_HALO_SEND), "f64")
        recv = _view(self.engines[0].exchange_tensor(BUF_HALO_RECV), "f64")
        h = send.numel() // 2
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, send[:h].contiguous(), self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, recv[:h], self.rank - 1, self.group))
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.isend, send[h:].contiguous(), self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, recv[h:], self.rank + 1, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def setup_fabric(self, timeout_s: float = 0.0) -> bool:
        """One process per GPU: every rank exports its exchange block as a CUDA IPC handle, maps the others'
        (peer access over NVLink) and attaches. Returns False — on EVERY rank — if any rank could not (no
        peer access, not NCCL, ...); the caller then keeps the collective-based phase protocol."""
        e = self.engines[0]
        ok, mine = True, None
        try:
            if self.dist.get_backend(self.group) != "nccl" or not hasattr(e, "fabric_info"):
                raise RuntimeError("peer-memory fabric needs CUDA shards")
            mine = e.fabric_info()
        except Exception as ex:                    # noqa: BLE001 - any failure means "use the phase protocol"
            logging.info("fabric unavailable on rank %d: %s", self.rank, ex)
            ok = False
        infos = [None] * self.world
        self.dist.all_gather_object(infos, (ok, mine[2] if mine else None), group=self.group)
        ok = all(i[0] for i in infos)
        self._ipc = []
        if ok:
            try:
                ptrs = []
                for r, (_, handle) in enumerate(infos):
                    if r == self.rank:
                        ptrs.append(mine[0])
                    else:
                        ptrs.append(e.ipc_open(handle))
                        self._ipc.append(ptrs[-1])
                e.fabric_attach(ptrs, timeout_s)
            except Exception as ex:                # noqa: BLE001
                logging.warning("fabric attach failed on rank %d: %s", self.rank, ex)
                ok = False
        flags = [None] * self.world
        self.dist.all_gather_object(flags, ok, group=self.group)
        self.dist.barrier(group=self.group)        # every block has been zeroed before anybody's first push
        return all(flags)

    _CTL_SLOTS = 64

    def shared_mirror(self, nbytes: int) -> np.ndarray:
        """ONE POSIX shared-memory array for the whole node: rank 0 creates it, every rank maps it, and each
        rank's GPU writes its own slice (bossgpu_set_strat_mirror). No gather of masks is ever needed: after
        the barrier that ends an update every process sees every contig's strategy. The segment's tail holds
        2 x 64 control words: the per-update host barrier / agreement between the ranks is a spin on those
        (about a microsecond) instead of a collective."""
        from multiprocessing import resource_tracker, shared_memory
        name = [None]
        ctl_off = (max(nbytes, 1) + 63) // 64 * 64
        total = ctl_off + 2 * self._CTL_SLOTS * 8
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=total)
            np.frombuffer(self._shm.buf, dtype=np.int64, offset=ctl_off, count=2 * self._CTL_SLOTS)[:] = 0
            np.frombuffer(self._shm.buf, dtype=np.uint8)[:nbytes] = 1          # Contig.strat starts all-accept
            name[0] = self._shm.name
        self.dist.broadcast_object_list(name, src=0, group=self.group)
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
            try:                                             # the creator owns the segment's lifetime
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:
                pass
        self.dist.barrier(group=self.group)
        if self.world <= self._CTL_SLOTS:
            self._ctl = np.frombuffer(self._shm.buf, dtype=np.int64, offset=ctl_off, count=2 * self._CTL_SLOTS).reshape(2, -1)
            self._ctl_seq = 0
        return np.frombuffer(self._shm.buf, dtype=np.uint8)[:nbytes]

    def _ctl_round(self, ok: bool, timeout_s: float = 60.0) -> bool:
        """All ranks post (sequence number, ok) and wait for each other; slots alternate by sequence parity so
        a rank that is already one round ahead cannot overwrite a word somebody still has to read."""
        import time
        self._ctl_seq += 1
        seq, row = self._ctl_seq, self._ctl[self._ctl_seq & 1]
        row[self.rank] = seq * 2 + (1 if ok else 0)
        t0 = time.perf_counter()
        while True:
            vals = row[: self.world]
            if (vals >= seq * 2).all():
                return bool((vals & 1).all())
            if time.perf_counter() - t0 > timeout_s:
                raise TimeoutError(f"rank {self.rank}: peers did not reach host round {seq}")

    def close(self) -> None:
        e = self.engines[0] if self.engines else None
        for p in getattr(self, "_ipc", []):
            try:
                e.ipc_close(p)
            except Exception:
                pass
        self._ipc = []
        self._ctl = None
        shm = getattr(self, "_shm", None)
        if shm is not None:
            self._shm = None
            try:
                shm.close()
            except Exception:                  # a view of the segment is still alive somewhere: leave the mapping to the OS
                pass
            if self.rank == 0:
                try:
                    shm.unlink()
                except Exception:
                    pass

    def agree(self, ok: bool) -> bool:
        if getattr(self, "_ctl", None) is not None:
            return self._ctl_round(ok)
        import torch
        dev = self.engines[0].exchange_tensor(BUF_SWITCH).device
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def barrier(self) -> None:
        if getattr(self, "_ctl", None) is not None:
            self._ctl_round(True)
            return
        self.dist.barrier(group=self.group)


# ------------------------------------------------------------------------------------------------------
# the sharded run
# ------------------------------------------------------------------------------------------------------
class _Pieces:
    """Engine-shaped accessor for a contig whose sites may be spread over several local segments: what
    `runs.Contig` fetches through (`coverage`, `scores`, `scores_ds`, `benefit`, `buckets`)."""

    def __init__(self, run: "ShardedRun"):
        self.run = run

    def _cat(self, contig: int, fn):
        pieces = self.run._pieces[contig]
        if not pieces or not self.run._complete[contig]:
            raise RuntimeError("this contig has segments on other ranks; per-site state is only assembled for local contigs")
        parts = [fn(e, i) for e, i in pieces]
        return parts[0] if len(parts) == 1 else tuple(np.concatenate(x) for x in zip(*parts)) if isinstance(parts[0], tuple) \
            else np.concatenate(parts)

    def coverage(self, contig):
        return self._cat(contig, lambda e, i: e.coverage(i))

    def scores(self, contig, entropy=False):
        return self._cat(contig, lambda e, i: e.scores(i, entropy=entropy))

    def scores_ds(self, contig):
        return self._cat(contig, lambda e, i: e.scores_ds(i))

    def benefit(self, contig, debug=False):
        return self._cat(contig, lambda e, i: e.benefit(i, debug=debug))

    def buckets(self, contig):
        pieces = self.run._pieces[contig]
        sw = np.concatenate([e.buckets(i)[0] for e, i in pieces])
        return sw, np.full(self.run.nbarcodes, bool(sw.any()))


class ShardedRun(BossRuns):
    """`BossRuns` over N shards. In a torch.distributed job (one process per GPU) every rank constructs it with
    the same arguments and calls the same methods with the same batch; rank 0 ends up with every contig's
    `strat` (and writes boss.npz). With `n_virtual` it instead drives that many shards on one GPU."""

    def __init__(self, *args, n_virtual: int | None = None, halo_bins: int = 2048, group=None,
                 engine_factory=Engine, exchange: str = "auto", fabric_timeout_s: float = 0.0, **kw):
        """`exchange`: "fabric" = the GPUs exchange over peer memory inside the update (csrc/fabric.cuh; no host
        round trip), "phases" = the host runs a collective after each phase (NCCL / gloo / copies between
        virtual shards), "auto" = fabric between processes when every rank can map its peers, else phases;
        virtual shards default to phases (they share one stream unless fabric is asked for)."""
        if exchange not in ("auto", "fabric", "phases"):
            raise ValueError("exchange must be 'auto', 'fabric' or 'phases'")
        self._exchange = exchange
        self._fabric_timeout_s = float(fabric_timeout_s)
        self._fabric = False
        self._n_virtual = n_virtual
        self._halo_bins = int(halo_bins)
        self._dist_group = group
        self._engine_factory = engine_factory
        super().__init__(*args, **kw)

    # -- construction ----------------------------------------------------------------------------------
    def _create_engine(self, device: int, stream) -> None:
        lens = [c.length for c in self.contigs_filt.values()]
        codes = [c.seq_int for c in self.contigs_filt.values()]
        if self._n_virtual is not None:
            n, mine = int(self._n_virtual), list(range(int(self._n_virtual)))
        else:
            import torch.distributed as dist
            n, mine = dist.get_world_size(self._dist_group), [dist.get_rank(self._dist_group)]
        self.plan = plan_shards(lens, n)
        self.row_start = merged_row_starts(lens, self.plan)
        for segs in self.plan:
            for s in segs:
                head, tail = s.start == 0, s.start + s.length == lens[s.contig]
                if not head and not tail and s.length // BIN < self._halo_bins:
                    raise ValueError("a shard lies strictly inside a contig and is shorter than the bin halo")
        self.engines = []
        self._streams = []
        want_fabric_local = self._n_virtual is not None and self._exchange == "fabric"
        for sid in mine:
            segs = self.plan[sid]
            st = stream
            if want_fabric_local:
                import torch
                self._streams.append(torch.cuda.Stream(device=device))     # one stream per virtual shard
                st = self._streams[-1].cuda_stream
            e = self._engine_factory(contig_lengths=lens, ref_codes=[codes[s.contig][s.start:s.start + s.length] for s in segs],
                                     n_barcodes=self.nbarcodes, ploidy=self.ploidy, n_sites_total=int(self.ref.n_sites),
                                     device=device, stream=st, segments=segs, halo_bins=self._halo_bins)
            e.set_shards(n, sid, self.row_start)
            self.engines.append(e)
        self.engine = self.engines[0]
        self.group = LocalGroup(self.engines) if self._n_virtual is not None else DistGroup(self.engines[0], self._dist_group)
        # one host array holds every contig's strategy; each shard's kernel keeps its slice current
        srow = np.concatenate(([0], np.cumsum(np.asarray(lens, dtype=np.int64) // BIN)))
        per_row = 2 * self.nbarcodes
        self._strat_global = self.group.shared_mirror(int(srow[-1]) * per_row)
        self.engine.host_register(self._strat_global)          # once per process: the shards' slices share pages
        for sid, e in zip(mine, self.engines):
            s0 = self.plan[sid][0]
            off = (int(srow[s0.contig]) + s0.start // BIN) * per_row
            e.set_strat_mirror(self._strat_global[off: off + e.seg_strat_rows(-1) * per_row], registered=True)
        flat = self._strat_global.view(np.bool_).reshape(-1, 2, self.nbarcodes)
        self._global_views = [flat[int(srow[k]): int(srow[k + 1])] for k in range(len(lens))]
        # contig -> [(engine, local segment index)] in position order, and whether that covers the contig
        self._pieces = {k: [] for k in range(len(lens))}
        for e in self.engines:
            for i, s in enumerate(e.segments):
                self._pieces[s.contig].append((e, i))
        self._complete = {k: sum(e.segments[i].length for e, i in p) == lens[k] for k, p in self._pieces.items()}
        view = _Pieces(self)
        for k, c in enumerate(self.contigs_filt.values()):
            c._bind(view, k)
        if n > 1 and (want_fabric_local or (self._n_virtual is None and self._exchange in ("auto", "fabric"))):
            self._fabric = bool(self.group.setup_fabric(self._fabric_timeout_s))
            if self._exchange == "fabric" and not self._fabric:
                raise RuntimeError("exchange='fabric' was asked for but the shards cannot map each other's memory")
        logging.info("sharded update over %d shard(s): %s", n, "peer-memory fabric" if self._fabric else "phase protocol")
        # one process per GPU: this process' range of every contig; convert_records skips the reads outside it
        self._ranges = None
        if self._n_virtual is None and n > 1 and self.route_batches:
            self._ranges = np.zeros((len(lens), 2), dtype=np.int64)
            for s in self.plan[mine[0]]:
                self._ranges[s.contig] = (s.start, s.start + s.length)

    @property
    def exchange_mode(self) -> str:
        return "fabric" if self._fabric else "phases"

    def local_sites(self) -> int:
        return int(sum(s.length for e in self.engines for s in e.segments))

    def sync_depth_totals(self) -> None:
        """After loading counters directly (set_coverage / synth_coverage): make every shard's per-contig depth
        totals global. Batches ingested later keep them global on their own."""
        self.group.allreduce(BUF_COV_TOTAL, "sum", "i64")

    def synth_coverage(self, **kw) -> None:
        for e in self.engines:
            e.synth_coverage(**kw)
        self.sync_depth_totals()

    # -- coverage ---------------------------------------------------------------------------------------
    route_batches = True        # one process per GPU: convert only the reads this process' range of the genome sees

    def _convert(self, paf_dict, seqs, quals=None, barcodes=None) -> PackedBatch:
        if self._ranges is None:
            return super()._convert(paf_dict, seqs, quals, barcodes)
        err, b = None, None
        try:
            b = self.cc.convert_records(paf_dict=paf_dict, seqs=seqs, quals=quals, barcodes=barcodes, ranges=self._ranges)
        except Exception as ex:             # noqa: BLE001 - a bad record is seen by the rank that owns it only: agree before raising
            err = ex
        if not self.group.agree(err is None):
            raise err if err is not None else RuntimeError("another shard found a bad record while converting the batch")
        return b

    def _effect_increments(self, increments: PackedBatch) -> None:
        b = increments
        err = None
        try:
            if b.all_contig is not None:
                # routed batch: own reads + the whole batch's reference span per contig (dropout rule, reference.py:157-158)
                # (spans sum far below 2^53: the float64 weights of bincount are exact)
                cov_add = np.bincount(b.all_contig, weights=np.abs(b.all_tend - b.all_tstart),
                                      minlength=len(self.contigs_filt)).astype(np.int64)
                for e in self.engines:
                    e.ingest_records_routed(b.contig, b.tstart, b.tend, b.barcode, b.rev, b.cigar_ptr, b.cigar_len,
                                            b.seq_ptr, b.seq_from, b.seq_to, cov_add)
            else:
                for e in self.engines:
                    e.ingest_records_ptr(b.contig, b.tstart, b.tend, b.barcode, b.rev, b.cigar_ptr, b.cigar_len,
                                         b.seq_ptr, b.seq_from, b.seq_to)
        except (IndexError, AssertionError, ValueError) as ex:        # what upstream raises for a bad record
            err = ex
        if not self.group.agree(err is None):
            raise err if err is not None else RuntimeError("another shard rejected the batch")

    def _read_starts_to_device(self, wins, strands) -> None:
        for e in self.engines:                      # the window counts are small and replicated on every shard
            e.read_starts_add(wins, strands)

    # -- update -----------------------------------------------------------------------------------------
    def _phases(self, approx_ccl, time_cost, bucket_threshold, fhat_windows=None, fhat_scalars=None) -> UpdateOutcome:
        g = self.group
        ps = [e.params(approx_ccl, time_cost, bucket_threshold, debug=self.write_debug, fhat_windows=fhat_windows,
                       fhat_scalars=fhat_scalars) for e in self.engines]
        w_max = max(int(np.max(np.asarray(approx_ccl) // BIN)), 4)
        if w_max - 1 > self._halo_bins and g.n_shards > 1:
            raise ValueError(f"staircase window of {w_max} bins exceeds the bin halo ({self._halo_bins}) kept on shard edges")
        if self._fabric:
            # every kernel of the update, exchanges included, is enqueued on every local shard before anybody
            # synchronises; the GPUs meet each other at the four exchange steps on their own
            for e, p in zip(self.engines, ps):
                e.update_fused_begin(p)
            outs, err = [], None
            for e in self.engines:
                try:
                    outs.append(e.update_fused_end())
                except Exception as ex:            # noqa: BLE001 - every shard must be drained before raising
                    err = err or ex
            if err is not None:
                raise err
            return self._combine(outs)
        for e, p in zip(self.engines, ps):
            e.update_phase(0, p)
            e.halo_pack()
        g.exchange_halos()
        for e in self.engines:
            e.halo_unpack()
        on = bool(g.allreduce(BUF_SWITCH, "max", "i32").max().item())
        if on:
            for e, p in zip(self.engines, ps):
                e.update_phase(1, p)
            g.allreduce(BUF_NORM, "max", "i64")          # non-negative doubles order like their bit patterns
            for e, p in zip(self.engines, ps):
                e.update_phase(2, p)
            g.allreduce(BUF_HIST, "sum", "i64")          # integer limbs: exact, order-free
            for e, p in zip(self.engines, ps):
                e.update_phase(3, p)
            g.allgather_mask()
        outs = [e.update_phase(4, p) for e, p in zip(self.engines, ps)]
        return self._combine(outs)

    @staticmethod
    def _combine(outs) -> UpdateOutcome:
        out = outs[0]
        if len(outs) > 1:
            # counters of the LOCAL shards (one process per GPU: this rank's share; virtual shards: the whole genome)
            out.n_accept = (sum(o.n_accept[0] for o in outs), sum(o.n_accept[1] for o in outs))
            out.n_dropout = sum(o.n_dropout for o in outs)
            out.mirror_bytes = sum(o.mirror_bytes for o in outs)
        return out

    def update_wrapper(self) -> None:
        scalars = self.read_starts.pointmass_scalars()
        time_cost = getattr(self.rl_dist, "time_cost", None)
        try:
            out = self._phases(self.rl_dist.approx_ccl, np.float64("nan") if time_cost is None else time_cost,
                               self.bucket_threshold, fhat_scalars=scalars)
        except AttributeError:              # Q14: a bucket is on, no time_cost — the library left every strategy as it was
            self._pull_switches()
            raise
        self.last = out
        self._pull_switches()
        if out.switched_on:
            self.threshold = out.threshold
            self._pull_strategies()
            if self.group.rank == 0:
                self._write_contig_strategies(self.ref.get_strategy_dict())

    def device_update(self, approx_ccl, time_cost, bucket_threshold, fhat_windows=None) -> UpdateOutcome:
        self.last = self._phases(approx_ccl, time_cost, bucket_threshold, fhat_windows=fhat_windows)
        return self.last

    def pack_for_device(self, batch: PackedBatch, engine=None):
        """One packed dict per local shard (reads routed and clipped like the text path does)."""
        return [super(ShardedRun, self).pack_for_device(batch, engine=e) for e in self.engines]

    def ingest_device(self, ds, engine=None) -> None:
        for e, d in zip(self.engines, ds):
            super().ingest_device(d, engine=e)

    def _pull_switches(self) -> None:
        if self._switch_views is None:
            per_engine = {id(e): e.buckets_host() for e in self.engines}
            self._switch_views = {k: [per_engine[id(e)][i] for e, i in p] for k, p in self._pieces.items()}
        for k, c in enumerate(self.contigs_filt.values()):
            if not self._complete[k]:
                continue                                 # switches of contigs with segments on other ranks stay there
            vs = self._switch_views[k]
            c.bucket_switches = vs[0] if len(vs) == 1 else np.concatenate(vs)
            if not c.switched_on.all() and c.bucket_switches.any():
                c.switched_on[...] = True

    def _pull_strategies(self) -> None:
        """Every shard's kernel has written its changed chunks into the shared host array; once all shards are
        done (barrier) each contig's strategy — split across shards or not — is a contiguous view of it."""
        self.group.barrier()
        for c, v in zip(self.contigs_filt.values(), self._global_views):
            c.strat = v

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self) -> None:
        for e in getattr(self, "engines", []):
            e.close()
            e._mirror = None
        if getattr(self, "_strat_global", None) is not None:
            try:
                self.engine.host_unregister(self._strat_global)
            except Exception:
                pass
            # every view of the shared mirror has to go before the segment can be unmapped: contigs keep copies
            for c in self.contigs_filt.values():
                if isinstance(getattr(c, "strat", None), np.ndarray):
                    c.strat = np.array(c.strat)
            self._global_views = None
            self._strat_views = None
            self._strat_global = None
        if hasattr(getattr(self, "group", None), "close"):
            self.group.close()
