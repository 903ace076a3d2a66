"""Upstream's own `BossRuns` / `BossRunsSim` with the strategy update on the GPU.

    from boss_runs_b200.dropin import BossRunsGPU, BossRunsSimGPU

    exp = BossRunsGPU(args=conf.args)         # conf = boss.config.Config(parse=True): the unchanged TOML
    exp.init()                                # upstream init (Reference, Mapper, AbundanceTracker, ...) + device state
    exp.process_batch(exp.process_batch_runs) # upstream's loop; boss.npz is written by upstream's own code

`_GpuMixin` overrides exactly the array half of `boss.runs.core.BossRuns` (core.py:20-224):

    init                 after upstream's: an `Engine` over `contigs_filt`, the GPU-side `CoverageConverter`, a
                         `ReadStartDist` whose counts also go to the device; every `Contig.strat` /
                         `Contig.bucket_switches` becomes a view of the library's pinned host mirrors; the per-site host
                         arrays of upstream's `Contig` (`coverage`, `scores`, `entropy`, ... ~47 B/site) are released
                         and read back from the GPU on access
    process_batch_runs   upstream's body, preceded by the early start of the score pass
    _effect_increments   CIGAR expansion + scatter on the GPU (core.py:77-86, reference.py:122-144)
    update_wrapper       one `bossgpu_update` (core.py:160-198), then upstream's `_write_contig_strategies`

Everything else — config/TOML, `Reference`, `Mapper`, `AbundanceTracker`, the simulator's `Sampler` / `ReadCache` /
`make_decisions`, logging, the `boss.npz` layout readfish reads — is inherited untouched. The classes are built on
first access (PEP 562) so that importing this module does not need upstream's dependencies (mappy, bottleneck,
minknow_api); `make_classes(BossRuns, BossRunsSim)` builds them over any pair of base classes with that surface.

Options come from an optional `[gpu]` table of the same TOML (upstream's pydantic models ignore unknown tables, so
existing configs keep validating): `enabled` (default true), `device` (CUDA ordinal, default 0), `prescore` (split
score pass, default true), `lean_host` (release upstream's per-site host arrays and skip the 3.3 GB host score table,
default true), `strategy_format` ("npz" | "bits" | "both", default "npz"). `main()` mirrors `boss.BOSS:main`
(BOSS.py:20-62) and picks the GPU classes unless `[gpu] enabled = false`.
"""
from __future__ import annotations

import contextlib
import logging
import sys
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from ._lib import BIN
from .engine import Engine
from .hostmodel import ReadStartDist
from .runs import CoverageConverter

__all__ = ["GpuOptions", "gpu_options", "make_classes", "BossRunsGPU", "BossRunsSimGPU", "main"]


@dataclass
class GpuOptions:
    enabled: bool = True
    device: int = 0
    prescore: bool = True
    lean_host: bool = True
    strategy_format: str = "npz"


def gpu_options(toml_path: str | None) -> GpuOptions:
    """The `[gpu]` table of a BOSS TOML (absent table / absent file -> defaults). Unknown keys are an error, like
    pydantic's validation of the other tables."""
    opts = GpuOptions()
    if not toml_path:
        return opts
    import tomllib
    with Path(toml_path).open("rb") as fh:
        table = tomllib.load(fh).get("gpu", {})
    for key, val in table.items():
        if not hasattr(opts, key):
            raise ValueError(f"[gpu] {key}: unknown option (known: {', '.join(vars(opts))})")
        want = type(getattr(opts, key))
        if not isinstance(val, want) or (want is int and isinstance(val, bool)):
            raise ValueError(f"[gpu] {key}: expected {want.__name__}, got {val!r}")
        setattr(opts, key, val)
    if opts.strategy_format not in ("npz", "bits", "both"):
        raise ValueError("[gpu] strategy_format must be 'npz', 'bits' or 'both'")
    return opts


class _DeviceReadStarts(ReadStartDist):
    """`ReadStartDist` (readstartdist.py:11-152) whose `count_read_starts` also adds the batch to the device-side
    counter the update derives F-hat from; upstream's call sites (core.py:222, simulation.py:171) stay as they are."""

    def __init__(self, contigs: dict, engines, **kw):
        super().__init__(contigs=contigs, **kw)
        self._engines = engines

    def count_read_starts(self, paf_dict: dict):
        wins, strands = super().count_read_starts(paf_dict=paf_dict)
        for e in self._engines:
            e.read_starts_add(wins, strands)
        return wins, strands


_HEAVY = ("coverage", "change_mask", "scores", "entropy", "initial_scores", "scores_ds", "smu", "expected_benefit",
          "additional_benefit")


_BACKED: dict[type, type] = {}


def _device_backed(contig_cls):
    """Subclass of upstream's `Contig` whose per-site / per-bin arrays are fetched from the GPU on access
    (tests, debugging, `AbundanceTracker`-style readers) instead of living on the host."""
    if contig_cls in _BACKED:
        return _BACKED[contig_cls]

    def fetch(name):
        def get(self):
            e, seg = self._gpu
            if name == "coverage":
                return e.coverage(seg)
            if name in ("scores", "entropy"):
                s, en = e.scores(seg, entropy=True)
                return s if name == "scores" else en
            if name == "scores_ds":
                return e.scores_ds(seg)
            if name == "additional_benefit":
                return e.benefit(seg)
            if name in ("smu", "expected_benefit"):
                return e.benefit(seg, debug=True)[1 if name == "smu" else 2]
            if name == "initial_scores":
                return np.full((self.length, self.nbarcodes), self.score0[0])
            raise AttributeError(f"{name} is not kept by the GPU path")       # change_mask: the GPU scores every site

        def put(self, value):                                                 # upstream assigns cont.scores, cont.entropy
            raise AttributeError(f"Contig.{name} lives on the GPU; it cannot be assigned")
        return property(get, put)

    _BACKED[contig_cls] = type("Contig", (contig_cls,), {n: fetch(n) for n in _HEAVY} | {"__doc__": contig_cls.__doc__})
    return _BACKED[contig_cls]


@contextlib.contextmanager
def _without_host_score_table(base_init_globals: dict):
    """Upstream's `init` fills a (40,)*5 + (4,) float64 table twice (3.3 GB each, ~4 s; sequences.py:347-393) that only
    `Scoring.update_scores` reads — which the GPU replaces with its own dense table (csrc/table.cuh)."""
    scoring = base_init_globals.get("Scoring")
    if scoring is None or not hasattr(scoring, "init_score_array"):
        yield
        return
    original = scoring.init_score_array
    scoring.init_score_array = lambda self, *a, **k: None
    try:
        yield
    finally:
        scoring.init_score_array = original


class _GpuMixin:
    """See the module docstring. Must precede the upstream class in the MRO."""

    gpu: GpuOptions = GpuOptions()
    engine_factory = Engine           # tests substitute a recording stub
    gpu_debug = False                 # also keep S_mu / expected benefit on the device for `Contig.smu` etc. (parity tests)

    # -- construction ----------------------------------------------------------------------------------
    def init(self) -> None:
        base_init = super().init
        patch = _without_host_score_table(getattr(base_init, "__globals__", {})) if self.gpu.lean_host else contextlib.nullcontext()
        with patch:
            base_init()                                   # core.py:23-55, unchanged
        self._attach_gpu()

    def _attach_gpu(self) -> None:
        filt = self.contigs_filt
        if not filt:
            raise ValueError("no contig of at least 100 kb left to track")
        self.engine = self.engine_factory(
            contig_lengths=[c.length for c in filt.values()], ref_codes=[c.seq_int for c in filt.values()],
            n_barcodes=self.nbarcodes, ploidy=self.args.optional.ploidy, n_sites_total=int(self.ref.n_sites),
            device=self.gpu.device)
        self.cc = CoverageConverter({name: i for i, name in enumerate(filt)})
        # upstream asserts the F-hat length drift (readstartdist.py:131); keep the assertion where upstream has it
        self.read_starts = _DeviceReadStarts(contigs=filt, engines=[self.engine], strict=True)
        flat, row = self.engine.strat_host(), 0          # pinned mirror the distribution kernel keeps current
        switches = self.engine.buckets_host()
        for seg, c in enumerate(filt.values()):
            n = c.length // BIN
            c.strat = flat[row: row + n]
            row += n
            c.bucket_switches = switches[seg]
            c._gpu = (self.engine, seg)
            if self.gpu.lean_host:
                for name in _HEAVY:
                    c.__dict__.pop(name, None)
                if type(c) not in _BACKED.values():
                    c.__class__ = _device_backed(type(c))
        self.threshold = None
        self.last = None

    # -- one batch (core.py:202-224) ---------------------------------------------------------------------
    def process_batch_runs(self, new_reads, new_quals) -> None:
        if self.gpu.prescore:
            self.engine.prescore_begin()                 # the GPU scores every tile while the host maps and converts
        super().process_batch_runs(new_reads, new_quals)

    def process_batch_runs_sim(self) -> None:            # simulation.py:139-190
        if self.gpu.prescore:
            self.engine.prescore_begin()
        super().process_batch_runs_sim()

    # -- coverage (core.py:77-86) --------------------------------------------------------------------------
    def _effect_increments(self, increments) -> None:
        b = increments
        if self.gpu.prescore:
            self.engine.prescore(b.contig, b.tstart, b.tend)   # the update re-scores only the tiles this batch writes to
        self.engine.ingest_records_ptr(b.contig, b.tstart, b.tend, b.barcode, b.rev, b.cigar_ptr, b.cigar_len,
                                       b.seq_ptr, b.seq_from, b.seq_to)

    # -- update (core.py:160-198) --------------------------------------------------------------------------
    def update_wrapper(self) -> None:
        time_cost = getattr(self.rl_dist, "time_cost", None)
        try:
            out = self.engine.update(approx_ccl=self.rl_dist.approx_ccl,
                                     time_cost=np.float64("nan") if time_cost is None else time_cost,
                                     bucket_threshold=self.args.optional.bucket_threshold,
                                     fhat_scalars=self.read_starts.pointmass_scalars(), debug=self.gpu_debug)
        finally:
            # reference.py:203-207: once any bucket of any barcode is on, the whole contig is flagged
            for c in self.contigs_filt.values():
                if c.bucket_switches.any():
                    c.switched_on[...] = True
        self.last = out
        if out.switched_on:
            self.threshold = out.threshold
            acc = self.engine.seg_accept()
            for i, c in enumerate(self.contigs_filt.values()):
                rows = max(c.strat.shape[0], 1)
                logging.info(f"{c.name}: {acc[i, 0] / rows}, {acc[i, 1] / rows}")          # core.py:152-154
            self._write_contig_strategies(contig_strats=self.ref.get_strategy_dict())     # core.py:59-69 (inherited)

    def _write_contig_strategies(self, contig_strats) -> None:
        fmt = self.gpu.strategy_format
        if fmt in ("npz", "both"):
            super()._write_contig_strategies(contig_strats=contig_strats)
        if fmt in ("bits", "both") and getattr(self, "engine", None) is not None:
            from . import stratfile
            tracked = [(n, c.length // BIN) for n, c in self.contigs_filt.items()]
            stratfile.write_bits(f"{self.out_dir}/masks/boss.bits", tracked, self.nbarcodes, self.engine.strat_packed(),
                                 rejected=[n for n, c in self.contigs.items() if c.rej])


def make_classes(boss_runs_cls, boss_runs_sim_cls=None, options: GpuOptions | None = None):
    """`(BossRunsGPU, BossRunsSimGPU)` over the given base classes (upstream's, or anything with their surface)."""
    ns = {"gpu": options or GpuOptions()}
    live = type("BossRunsGPU", (_GpuMixin, boss_runs_cls), dict(ns))
    sim = type("BossRunsSimGPU", (_GpuMixin, boss_runs_sim_cls), dict(ns)) if boss_runs_sim_cls is not None else None
    return live, sim


_CLASSES: dict[str, type] = {}


def __getattr__(name: str):
    if name in ("BossRunsGPU", "BossRunsSimGPU"):
        if not _CLASSES:
            try:
                from boss.runs.core import BossRuns
                from boss.runs.simulation import BossRunsSim
            except ImportError as ex:
                raise ImportError("boss_runs_b200.dropin needs the upstream package `boss` (BOSS-RUNS) and its dependencies "
                                  f"on the path: {ex}. `boss_runs_b200.runs.BossRuns` is the same engine without upstream.") from ex
            _CLASSES["BossRunsGPU"], _CLASSES["BossRunsSimGPU"] = make_classes(BossRuns, BossRunsSim)
        return _CLASSES[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def _toml_arg(argv) -> str | None:
    for i, a in enumerate(argv):
        if a == "--toml" and i + 1 < len(argv):
            return argv[i + 1]
        if a.startswith("--toml="):
            return a.split("=", 1)[1]
    return None


def main() -> None:
    """`boss.BOSS:main` (BOSS.py:20-62) with the BOSS-RUNS classes replaced by the GPU ones. BOSS-AEONS configurations
    (no `general.ref`) are handed to upstream's own entry point unchanged."""
    from time import sleep

    import boss.config
    opts = gpu_options(_toml_arg(sys.argv[1:]))
    conf = boss.config.Config(parse=True)
    if not conf.args.general.ref or not opts.enabled:
        import boss.BOSS
        return boss.BOSS.main()
    from boss.runs.core import BossRuns
    from boss.runs.simulation import BossRunsSim
    live_cls, sim_cls = make_classes(BossRuns, BossRunsSim, opts)
    if conf.args.live.device:
        exp = live_cls(args=conf.args)
        exp.init()
        exp.launch_live_components()
        try:
            while True:
                wait = exp.process_batch(exp.process_batch_runs)
                if wait > 0:
                    sleep(wait)
        except KeyboardInterrupt:
            print("exiting after keyboard interrupt.. ")
    else:
        exp = sim_cls(args=conf.args)
        exp.init_sim()
        while exp.batch < conf.args.simulation.maxb:
            exp.process_batch_sim(exp.process_batch_runs_sim)
        exp.cleanup()


if __name__ == "__main__":
    main()
