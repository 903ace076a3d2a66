"""`boss_runs_b200.dropin`: the GPU strategy update underneath upstream's `BossRuns` / `BossRunsSim`.

* CPU, reference mounted: the mixin on upstream's REAL classes, constructed from an unchanged TOML through upstream's own
  `Config`, with a recording stub in place of the `Engine` (no GPU here) — a live-style batch and two simulated batches on
  upstream's test data; checks the wiring (what reaches the engine, what upstream's own code writes to boss.npz).
* CPU, always: the `[gpu]` TOML table.
* GPU (`-m gpu`): the mixin over a stand-in for upstream's surface (tests/upstream_standin.py) with the real engine, state by
  state against the oracle on golden cases, plus boss.npz and the Q14 error."""
import io
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")


def test_gpu_options_table(tmp_path):
    from boss_runs_b200.dropin import GpuOptions, gpu_options
    assert gpu_options(None) == GpuOptions()
    t = tmp_path / "a.toml"
    t.write_text('[general]\nname = "x"\n')
    assert gpu_options(str(t)) == GpuOptions()
    t.write_text('[general]\nname = "x"\n[gpu]\ndevice = 3\nprescore = false\nstrategy_format = "both"\n')
    assert gpu_options(str(t)) == GpuOptions(device=3, prescore=False, strategy_format="both")
    for bad in ('[gpu]\nspeed = 11\n', '[gpu]\ndevice = "zero"\n', '[gpu]\ndevice = true\n', '[gpu]\nstrategy_format = "xml"\n'):
        t.write_text(bad)
        with pytest.raises(ValueError):
            gpu_options(str(t))


def test_import_without_upstream_is_explained():
    code = "import boss_runs_b200.dropin as d; d.make_classes; d.BossRunsGPU"
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", code], cwd=REPO, env=env, capture_output=True, text=True)
    assert r.returncode != 0 and "needs the upstream package" in r.stderr


_DRIVER = r'''
import io, json, sys, types
from pathlib import Path
import numpy as np
sys.dont_write_bytecode = True
from boss_runs_b200 import dropin, synth
from boss_runs_b200.engine import UpdateOutcome

class StubEngine:
    """Records what the mixin hands to the engine; the 'update' flips a recognisable pattern into the mirror."""
    log = []
    def __init__(self, contig_lengths, ref_codes, n_barcodes, ploidy, n_sites_total, device):
        self.lens, self.nb = [int(x) for x in contig_lengths], n_barcodes
        assert all(len(c) == n for c, n in zip(ref_codes, self.lens))
        assert all(set(np.unique(c)) <= {0, 1, 2, 3} for c in ref_codes)
        self.mirror = np.ones((sum(n // 100 for n in self.lens), 2, n_barcodes), dtype=np.bool_)
        self.sw = [np.zeros((n // 20000 + 1, n_barcodes), dtype=np.bool_) for n in self.lens]
        self.n_updates = 0
        StubEngine.log.append(("create", dict(n=len(self.lens), nb=n_barcodes, ploidy=ploidy, n_sites=n_sites_total, device=device)))
    def strat_host(self): return self.mirror
    def buckets_host(self): return self.sw
    def prescore_begin(self): StubEngine.log.append(("prescore_begin",))
    def prescore(self, contig, tstart, tend): StubEngine.log.append(("prescore", len(contig)))
    def ingest_records_ptr(self, contig, tstart, tend, barcode, rev, cp, cl, sp, sf, st):
        assert len({len(x) for x in (contig, tstart, tend, barcode, rev, cp, cl, sp, sf, st)}) == 1
        assert (np.asarray(contig) >= 0).all() and (np.asarray(contig) < len(self.lens)).all()
        StubEngine.log.append(("ingest", len(contig), int(np.sum(np.abs(np.asarray(tend) - np.asarray(tstart))))))
    def read_starts_add(self, wins, strands): StubEngine.log.append(("read_starts", len(wins)))
    def update(self, approx_ccl, time_cost, bucket_threshold, fhat_scalars, debug=False):
        assert len(approx_ccl) == 10 and len(fhat_scalars) == 3
        self.n_updates += 1
        for s in self.sw: s[0, :] = True
        self.mirror[self.n_updates::7, 0, :] = False
        StubEngine.log.append(("update", float(time_cost), float(bucket_threshold)))
        return UpdateOutcome(True, 0.5, 3, 1.0, 0.1, 1.0, 10, 0, (1, 2), 64)
    def seg_accept(self): return np.ones((len(self.lens), 2), dtype=np.int64)
    def coverage(self, seg): return np.zeros((self.lens[seg], 5, self.nb), dtype=np.uint16)

def model_engine():
    """The NumPy model of the library (tests/shard_model.py: oracle scoring + the kernels' phase structure) behind the engine
    interface the mixin talks to: upstream's classes + mixin + this must reproduce what upstream alone computes."""
    from shard_model import NumpyShardEngine
    from boss_runs_b200._lib import BUF_SWITCH
    from boss_runs_b200.sharding import plan_shards

    class ModelEngine(NumpyShardEngine):
        def __init__(self, contig_lengths, ref_codes, n_barcodes, ploidy, n_sites_total, device):
            plan = plan_shards([int(x) for x in contig_lengths], 1)
            super().__init__(contig_lengths, ref_codes, n_barcodes=n_barcodes, ploidy=ploidy, n_sites_total=n_sites_total,
                             segments=plan[0], halo_bins=0)
            self.set_shards(1, 0, [0, self.M_rows])
            self.thresholds = []
        def prescore_begin(self): pass
        def prescore(self, contig, tstart, tend): pass
        def update(self, approx_ccl, time_cost, bucket_threshold, fhat_scalars, debug=False):
            p = self.params(approx_ccl, time_cost, bucket_threshold, fhat_scalars=fhat_scalars)
            self.update_phase(0, p)
            if self.buf[BUF_SWITCH][0]:
                for ph in (1, 2, 3):
                    self.update_phase(ph, p)
            out = self.update_phase(4, p)
            self.thresholds.append(out.threshold)
            return out
        def seg_accept(self):
            acc, row = [], 0
            for n in self.n_srows:
                acc.append([int(self._strat[row: row + n, 0].sum()), int(self._strat[row: row + n, 1].sum())])
                row += n
            return np.asarray(acc, dtype=np.int64)
    return ModelEngine

mode, toml = sys.argv[1], sys.argv[2]
sys.argv = ["boss", "--toml", toml]
import boss.config
from boss.paf import Paf
conf = boss.config.Config(parse=True)                       # upstream's own parsing of the unchanged TOML
opts = dropin.gpu_options(toml)
out = {"device": opts.device}
if mode == "live":
    cls = dropin.BossRunsGPU
    cls.engine_factory = StubEngine
    cls.gpu = opts
    exp = cls(args=conf.args)
    exp.init()
    assert type(exp).__mro__[2].__module__ == "boss.runs.core"
    c0 = next(iter(exp.contigs_filt.values()))
    assert "coverage" not in c0.__dict__ and "scores" not in c0.__dict__       # upstream's host arrays are released
    assert c0.coverage.shape == (c0.length, 5, 1)                              # ... and read from the engine on access
    assert not hasattr(exp.scoring, "score_arr")                               # the 3.3 GB host table is never built
    contigs = {n: c.seq for n, c in exp.contigs_filt.items()}
    rb = synth.read_batch(contigs, n_reads=200, seed=5, mean_len=2000.0, min_len=400, max_len=6000)
    paf_dict = Paf.parse_PAF(io.StringIO(rb.paf_text))
    exp.mapper.map_sequences = lambda sequences: paf_dict    # minimap2 is not installed here
    exp.rl_dist.update(read_lengths={rid: len(s) for rid, s in rb.seqs.items()})
    exp.process_batch_runs(rb.seqs, {rid: "5" * len(s) for rid, s in rb.seqs.items()})
    npz = np.load(Path(exp.out_dir) / "masks" / "boss.npz")
    row = 0
    for name, c in exp.contigs.items():
        if c.rej:
            assert npz[name].shape == (1,) and not npz[name].any()
            continue
        n = c.length // 100
        assert np.array_equal(npz[name], exp.engine.mirror[row: row + n]) and np.shares_memory(c.strat, exp.engine.mirror)
        assert c.switched_on.all()
        row += n
    out["npz_false"] = int(sum((~npz[k]).sum() for k in npz.files if npz[k].ndim == 3))
elif mode == "simmodel":
    cls = dropin.BossRunsSimGPU
    cls.engine_factory = model_engine()
    cls.gpu = opts
    exp = cls(args=conf.args)
    exp.init_sim()
    counts = {}
    decisions = exp.make_decisions
    def spy(**kw):
        res = decisions(**kw)
        counts["c"] = [int(x) for x in res[2:]]
        return res
    exp.make_decisions = spy
    while exp.batch < conf.args.simulation.maxb:
        exp.process_batch_sim(exp.process_batch_runs_sim)
    out["batches"] = exp.batch
    out["counts"] = counts["c"]
    out["threshold"] = float(exp.engine.thresholds[-1]).hex()
    out["approx_ccl"] = [int(x) for x in exp.rl_dist.approx_ccl]
    npz = np.load(Path(exp.out_dir) / "masks" / "boss.npz")
    out["strat"] = {n: np.packbits(npz[n].ravel()).tobytes().hex() for n in exp.contigs_filt}
    out["shares"] = all(np.shares_memory(c.strat, exp.engine._strat) for c in exp.contigs_filt.values())
    exp.cleanup()
else:
    cls = dropin.BossRunsSimGPU
    cls.engine_factory = StubEngine
    cls.gpu = opts
    exp = cls(args=conf.args)
    exp.init_sim()
    while exp.batch < conf.args.simulation.maxb:
        exp.process_batch_sim(exp.process_batch_runs_sim)
    exp.cleanup()
    out["batches"] = exp.batch
out["log"] = StubEngine.log
print("RESULT " + json.dumps(out))
'''


def _run_driver(mode, toml, cwd):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(REPO), str(REPO / "tests"), str(REPO / "oracle" / "shims"), str(REFERENCE)])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    r = subprocess.run([sys.executable, "-c", _DRIVER, mode, str(toml)], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    import json
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])


needs_reference = pytest.mark.skipif(not (REFERENCE / "boss" / "runs" / "core.py").is_file(), reason="reference checkout not mounted (GPU box)")


@needs_reference
def test_mixin_on_upstream_bossruns_from_toml(tmp_path):
    from boss_runs_b200 import synth
    contigs = synth.random_contigs({"c1": 140_000, "rejme": 120_000, "c2": 100_100, "tiny": 50_000}, seed=4)
    fa = tmp_path / "ref.fa"
    fa.write_text("".join(f">{n}\n{s}\n" for n, s in contigs.items()))
    (tmp_path / "ref.mmi").touch()
    toml = tmp_path / "run.toml"
    toml.write_text(textwrap.dedent(f'''
        [general]
        name = "dropin"
        ref = "{fa}"
        mmi = "{tmp_path / "ref.mmi"}"
        [optional]
        reject_refs = "rejme"
        bucket_threshold = 0
        [gpu]
        device = 1
        '''))
    res = _run_driver("live", toml, tmp_path)
    kinds = [e[0] for e in res["log"]]
    assert kinds == ["create", "prescore_begin", "prescore", "ingest", "read_starts", "update"]
    create = res["log"][0][1]
    assert create == dict(n=2, nb=1, ploidy=1, n_sites=140_000 + 100_100 + 4, device=1)
    n_reads = res["log"][3][1]
    assert 150 < n_reads <= 200 and res["log"][2][1] == n_reads and res["log"][4][1] <= n_reads
    assert res["log"][5][2] == 0.0 and res["npz_false"] > 0


@needs_reference
def test_mixin_on_upstream_simulation(tmp_path):
    data = REFERENCE / "data" / "BOSS_test_data"
    for f in ("zymo.fa", "ERR3152366_10k.fq", "ERR3152366_10k.paf", "ERR3152366_10k_trunc.paf"):
        if not (data / f).exists():
            pytest.skip(f"{f} not in the reference's test data")
        os.symlink(data / f, tmp_path / f)                      # upstream writes index files next to its inputs
    (tmp_path / "zymo.mmi").touch()
    toml = tmp_path / "sim.toml"
    toml.write_text(textwrap.dedent(f'''
        [general]
        name = "dropsim"
        ref = "{tmp_path / "zymo.fa"}"
        mmi = "{tmp_path / "zymo.mmi"}"
        [optional]
        bucket_threshold = 0
        [simulation]
        fq = "{tmp_path / "ERR3152366_10k.fq"}"
        paf_full = "{tmp_path / "ERR3152366_10k.paf"}"
        paf_trunc = "{tmp_path / "ERR3152366_10k_trunc.paf"}"
        batchsize = 600
        maxb = 2
        '''))
    res = _run_driver("sim", toml, tmp_path)
    kinds = [e[0] for e in res["log"]]
    assert res["batches"] == 2 and kinds.count("update") == 2 and kinds.count("ingest") == 2
    assert kinds[:2] == ["create", "prescore_begin"] and res["log"][0][1]["n"] == 9
    assert (tmp_path / "out_dropsim" / "masks" / "boss.npz").is_file()
    ingested = [e[1] for e in res["log"] if e[0] == "ingest"]
    starts = [e[1] for e in res["log"] if e[0] == "read_starts"]
    assert all(0 < s <= n <= 600 for s, n in zip(starts, ingested))          # read starts: accepted reads only (simulation.py:171)


@needs_reference
def test_dropin_with_model_engine_reproduces_upstream_simulation(tmp_path):
    """Numbers, not only call order: upstream's own `BossRunsSim` classes + the mixin + a NumPy model of the library (oracle
    scoring behind the engine interface) run BASELINE config 1 (whole `data/BOSS_test_data`, `batchsize = 4000, maxb = 1`)
    and must arrive where upstream ALONE arrived — decisions, read-length staircase, threshold, every mask bit of every
    contig, as recorded from upstream in tests/golden/c1_full.npz (run A)."""
    import numpy as np
    gold = np.load(REPO / "tests" / "golden" / "c1_full.npz")
    data = REFERENCE / "data" / "BOSS_test_data"
    for f in ("zymo.fa", "ERR3152366_10k.fq", "ERR3152366_10k.paf", "ERR3152366_10k_trunc.paf"):
        if not (data / f).exists():
            pytest.skip(f"{f} not in the reference's test data")
        os.symlink(data / f, tmp_path / f)
    (tmp_path / "zymo.mmi").touch()
    toml = tmp_path / "c1.toml"
    toml.write_text(textwrap.dedent(f'''
        [general]
        name = "c1model"
        ref = "{tmp_path / "zymo.fa"}"
        mmi = "{tmp_path / "zymo.mmi"}"
        [optional]
        ploidy = 1
        bucket_threshold = 0
        [simulation]
        fq = "{tmp_path / "ERR3152366_10k.fq"}"
        paf_full = "{tmp_path / "ERR3152366_10k.paf"}"
        paf_trunc = "{tmp_path / "ERR3152366_10k_trunc.paf"}"
        batchsize = 4000
        maxb = 1
        '''))
    res = _run_driver("simmodel", toml, tmp_path)
    assert res["batches"] == 1 and res["shares"]
    assert res["counts"] == [int(x) for x in gold["A0_counts"]]
    assert res["approx_ccl"] == [int(x) for x in gold["A0_approx_ccl"]]
    want_thr = float(gold["A0_threshold"])
    assert abs(float.fromhex(res["threshold"]) - want_thr) <= 1e-12 * want_thr
    names = [str(n) for n in gold["A_contigs"]]
    assert sorted(res["strat"]) == sorted(names)
    differing = total = 0
    for n in names:
        got = np.unpackbits(np.frombuffer(bytes.fromhex(res["strat"][n]), dtype=np.uint8))
        want = np.unpackbits(gold[f"A0_{n}_strat"])
        assert got.shape == want.shape, n
        differing += int((got != want).sum())
        total += want.size
    # the model sums every window directly (like the CUDA kernels), upstream carries a running accumulator: entries whose
    # benefit sits within rounding of the threshold may fall either side (tests/tolerances.py MASK_REL)
    print('mask entries differing from upstream:', differing, 'of', total)
    assert differing <= max(2, total // 100_000), (differing, total)


# ------------------------------------------------------------------------------------------------------
# GPU: the mixin with the real engine on a stand-in for upstream's surface
# ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["hap_nb3", "dip_nb1"])
def test_mixin_matches_oracle_on_gpu(case, lib, tmp_path):
    import helpers as H
    import upstream_standin as up
    from boss_runs_b200 import dropin
    from golden_io import load_case
    g = load_case(case)
    cls, _ = dropin.make_classes(up.BossRuns, None, dropin.GpuOptions(strategy_format="both"))
    cls.gpu_debug = True
    args = up.make_args(barcodes=g.barcodes, reject_refs=",".join(g.reject_refs) if g.reject_refs else None, ploidy=g.ploidy,
                        bucket_threshold=g.bucket_threshold)
    exp = cls(args, g.records, tmp_path / "out")
    exp.init()
    assert "coverage" not in next(iter(exp.contigs_filt.values())).__dict__
    orc = H.oracle_run(g.records, g.ploidy, g.reject_refs, g.barcodes, g.bucket_threshold)
    for bi, (paf, seqs, bcs) in enumerate(g.batches):
        pd = H.parse_batch(paf, bcs, g.barcodes is not None)
        upd_o = H.oracle_step(orc, pd, seqs)
        exp.rl_dist.update({rid: recs[0].qlen for rid, recs in pd.items()})
        exp.mapper.map_sequences = lambda sequences, pd=pd: pd
        exp.process_batch_runs(seqs, None)
        assert bool(exp.last.switched_on) == upd_o
        H.compare_state(exp, orc, upd_o, f"dropin/{case}/b{bi}")
        npz = np.load(tmp_path / "out" / "masks" / "boss.npz")
        for name, c in exp.contigs.items():
            assert np.array_equal(npz[name], c.strat)
    from boss_runs_b200 import stratfile
    sb = stratfile.StrategyBits(tmp_path / "out" / "masks", barcodes=g.barcodes)
    sb.reload()
    for name, arr in sb.as_dict().items():
        assert np.array_equal(arr, exp.contigs[name].strat), name


@pytest.mark.gpu
def test_mixin_missing_time_cost(lib, tmp_path):
    """Q14 through the drop-in: a bucket is on, no read length seen yet -> AttributeError; switches updated, strategies not."""
    import upstream_standin as up
    from boss_runs_b200 import dropin, synth
    contigs = synth.random_contigs({"a": 120_000}, seed=2)
    cls, _ = dropin.make_classes(up.BossRuns)
    exp = cls(up.make_args(bucket_threshold=0), list(contigs.items()), tmp_path / "out")
    exp.init()
    exp.mapper.map_sequences = lambda sequences: {}
    with pytest.raises(AttributeError):
        exp.process_batch_runs({}, {})
    c = exp.contigs["a"]
    assert c.bucket_switches.all() and c.switched_on.all() and c.strat.all()
