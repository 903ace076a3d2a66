#!/usr/bin/env python
"""bench.py — strategy-update latency and Gsites/s (BASELINE.json metric) on 1..N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2|...]

One "step" = one strategy update over one batch of synthetic mappings: coverage scatter -> score/bin pass
-> bucket switches -> smoothing -> exponent histogram -> threshold -> bucket-gated masks.

* `value` / `ms_per_step`: inputs (tokenised batch, F-hat) already resident in HBM, timed with CUDA events
  on the launching stream, max over ranks.
* `e2e`: the same update through the reference-facing API (`BossRuns.process_batch_runs`) with HOST
  buffers: record marshalling, C++ CIGAR tokeniser, H2D, kernels, D2H of every contig's mask.
* `roofline`: the score/bin pass (dominant kernel) against the measured HBM peak of MEASURED_PEAKS.json.
* `cpu_baseline` / `--impl reference`: the oracle port (NumPy restatement of the reference, oracle/) timed on
  this box's host cores on a bounded sample of the same workload.

Workloads (BASELINE.json configs): c3 = 3.1 Gb diploid, 25 contigs (the metric's configuration; default),
c2 = 4.6 Mb haploid, c4 = 24 barcodes x 5 Mb, c5 = many-contig metagenome. `--scale` shrinks a workload
for development.
"""
from __future__ import annotations

import argparse
import io
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

METRIC = "strategy_update_throughput"
UNIT = "Gsites/s"


# ------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------
def workload_spec(name: str, scale: float):
    from boss_runs_b200 import synth
    if name == "c3":
        lens = synth.grch38_like_lengths(int(3_100_000_000 * scale), 25)
        return dict(name="c3: synthetic 3.1 Gb diploid, 25 contigs, state Poisson(8) + 2% dropout + 1% deep regions, 4000 x 10 kb read batches",
                    lengths=lens, ploidy=2, nb=1, reads=4000, mean_len=10_000.0, depth=8.0)
    if name == "c2":
        return dict(name="c2: synthetic 4.6 Mb haploid, 4000 x 10 kb read batches", lengths=[int(4_600_000 * scale)],
                    ploidy=1, nb=1, reads=4000, mean_len=10_000.0, depth=8.0)
    if name == "c4":
        return dict(name="c4: 24 barcodes x 5 Mb", lengths=[int(5_000_000 * scale)], ploidy=1, nb=24, reads=4000,
                    mean_len=10_000.0, depth=4.0)
    if name == "c5":
        rng = np.random.default_rng(13)
        lens = np.exp(rng.uniform(np.log(100_000), np.log(5_000_000), size=int(150 * scale) or 1)).astype(np.int64)
        return dict(name="c5: metagenome-style, 150 contigs 100 kb-5 Mb", lengths=[int(x) for x in lens], ploidy=1, nb=1,
                    reads=4000, mean_len=10_000.0, depth=8.0)
    raise SystemExit(f"unknown workload {name}")


def random_codes(lengths, seed=7):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 4, size=int(n), dtype=np.uint8) for n in lengths]


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 100 ms while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample
# ------------------------------------------------------------------------------------------------------
def cpu_sample(spec, steps: int, warmup: int, sample_sites: int = 20_000_000, sample_reads: int = 26):
    """Oracle (NumPy restatement of the reference's CPU path) on ONE contig of `sample_sites` with the same
    ploidy / read model / pre-loaded depth as the workload and a proportionally scaled batch.
    Returns (sites per second, description, seconds per update)."""
    sys.path.insert(0, str(REPO / "tests"))
    from oracle import boss_oracle as bo
    from boss_runs_b200 import synth
    from boss_runs_b200.hostmodel import parse_PAF

    L = int(min(sample_sites, sum(spec["lengths"])))
    L = max(L, 120_000)
    total = sum(spec["lengths"]) * spec["nb"]
    n_reads = max(4, int(round(spec["reads"] * L * spec["nb"] / total))) if total > L else spec["reads"]
    n_reads = min(n_reads, spec["reads"], sample_reads if total > L else spec["reads"])
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 4, size=L, dtype=np.uint8)
    barcodes = [f"barcode{i + 1:02d}" for i in range(spec["nb"])] if spec["nb"] > 1 else None

    class _Seq(str):
        pass
    run = bo.OracleRun.__new__(bo.OracleRun)
    run.barcodes, run.nb = barcodes, spec["nb"]
    hap = bo.ScoreModel(1)
    c = bo.ContigState("s1", codes, nb=spec["nb"], score0=hap.score0, ent0=hap.ent0)
    run.contigs = {"s1": c}
    run.contigs_filt = run.contigs
    run.n_sites = L
    run.model = bo.ScoreModel(spec["ploidy"])
    run.model.build_table()
    run.read_starts = bo.ReadStarts(run.contigs_filt)
    run.rl = bo.ReadLengths()
    run.bucket_threshold = 5
    # pre-loaded sequencing state, as if earlier batches had been ingested
    depth = rng.poisson(spec["depth"], size=(L, spec["nb"]))
    ref_cnt = rng.binomial(depth, 0.9)
    rest = depth - ref_cnt
    c.coverage[np.arange(L), codes, :] = ref_cnt.astype(np.uint16)
    c.coverage[:, 4, :] += (rest // 2).astype(np.uint16)
    c.coverage[np.arange(L), (codes + 1) & 3, :] += (rest - rest // 2).astype(np.uint16)
    c.change_mask[:] = True
    contigs = {"s1": codes}
    times = []
    for it in range(warmup + steps):
        rb = synth.read_batch(contigs, n_reads=n_reads, seed=900 + it, mean_len=spec["mean_len"],
                              n_barcodes=spec["nb"] if spec["nb"] > 1 else 0)
        pd = parse_PAF(io.StringIO(rb.paf_text))
        for rid, recs in pd.items():
            for r in recs:
                r.barcode = rb.barcodes.get(rid) if barcodes else None
        run.rl.update({rid: recs[0].qlen for rid, recs in pd.items()})
        t0 = time.perf_counter()
        if it == 0:
            cm = c.change_mask.copy()
        run.ingest(pd, rb.seqs)
        if it == 0:
            c.change_mask |= cm           # first update scores the whole pre-loaded state
        run.read_starts.count(pd)
        run.update()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.median(times))
    desc = (f"oracle port (NumPy restatement of the reference CPU path): 1 contig of {L} sites x {spec['nb']} barcode(s), "
            f"ploidy {spec['ploidy']}, pre-loaded depth ~{spec['depth']}, {n_reads} reads of ~{int(spec['mean_len'])} bp per update; "
            f"median of {steps} updates after {warmup} warm-up")
    return L * spec["nb"] / sec, desc, sec


def run_reference_arm(args, spec):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 5)
    warm = min(args.warmup, 1)
    v, desc, sec = cpu_sample(spec, steps=steps, warmup=max(warm, 1))
    line = {"metric": METRIC, "value": v / 1e9, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": steps,
            "warmup": max(warm, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": spec["name"]},
            "cpu_baseline": {"value": v / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc},
            "e2e": {"value": v / 1e9, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args, spec):
    import torch
    from boss_runs_b200 import build, synth
    from boss_runs_b200.hostmodel import parse_PAF
    from boss_runs_b200.runs import BossRuns
    
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: boss_runs_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if rank == 0:
        build.build()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        from boss_runs_b200.sharding import ShardedRun
    lengths = spec["lengths"]
    names = [f"ctg{i + 1}" for i in range(len(lengths))]
    codes = random_codes(lengths)
    total_sites = int(sum(lengths)) * spec["nb"]
    barcodes = [f"barcode{i + 1:02d}" for i in range(spec["nb"])] if spec["nb"] > 1 else None

    t_setup = time.time()
    if world > 1:
        run = ShardedRun(contigs=dict(zip(names, codes)), ploidy=spec["ploidy"], barcodes=barcodes, bucket_threshold=5,
                         device=local, strict_upstream_asserts=False, exchange=args.exchange, fabric_timeout_s=30.0)
    else:
        run = BossRuns(contigs=dict(zip(names, codes)), ploidy=spec["ploidy"], barcodes=barcodes, bucket_threshold=5,
                       device=local, strict_upstream_asserts=False)
    eng = run.engine
    synth_kw = dict(seed=11, mean_depth=spec["depth"], p_ref=0.90, p_del=0.04, frac_dropout=0.02, frac_deep=0.01)
    (run if world > 1 else eng).synth_coverage(**synth_kw)

    # read batches: text form for the end-to-end leg, packed + device-resident for the kernel leg
    n_batches = 3
    contig_arrays = dict(zip(names, codes))
    batches = []
    for b in range(n_batches):
        rb = synth.read_batch(contig_arrays, n_reads=spec["reads"], seed=1000 + b, mean_len=spec["mean_len"],
                              n_barcodes=spec["nb"] if spec["nb"] > 1 else 0)
        pd = parse_PAF(io.StringIO(rb.paf_text))
        for rid, recs in pd.items():
            for r in recs:
                r.barcode = rb.barcodes.get(rid) if barcodes else None
        batches.append((pd, rb.seqs, rb))
    run.rl_dist.update({rid: recs[0].qlen for pd, _, _ in batches for rid, recs in pd.items()})
    for pd, _, _ in batches:
        run.count_read_starts(pd)
    setup_s = time.time() - t_setup

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: end to end through the public API, host buffers ------------------------------------------
    e2e_times, e2e_parts = [], []
    h2d = d2h = 0
    for it in range(args.warmup + args.steps):
        pd, seqs, rb = batches[it % n_batches]
        barrier()
        t0 = time.perf_counter()
        # the body of BossRuns.process_batch_runs, statement by statement, so that its parts can be timed
        run._prescore_begin()               # split score/bin pass: the GPU scores every tile while the host prepares the batch
        inc = run.cc.convert_records(paf_dict=pd, seqs=seqs)
        t1 = time.perf_counter()
        run._prescore(inc)                  # ... and will re-score only the tiles this batch writes to
        t1b = time.perf_counter()
        run._effect_increments(inc)
        t2 = time.perf_counter()
        run.update_wrapper()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            e2e_times.append(dt)
            e2e_parts.append((t1 - t0, t1b - t1, t2 - t1b, time.perf_counter() - t2))
            # what the library staged and copied: per-read scalars + CIGAR op slots (4 B) + read bases packed 2 bits each
            h2d = sum(e.ingest_bytes() for e in getattr(run, "engines", [eng]))
            h2d += 20 * len(inc) * len(getattr(run, "engines", [eng]))     # the announced intervals (contig i32, tstart/tend i64)
            # masks reach the host as the 4 KB chunks that changed (written by the distribution kernel into the
            # pinned mirror Contig.strat views) + bucket switches + the result record
            d2h = int(run.last.mirror_bytes) + int(sum(c.bucket_switches.size for c in run.contigs_filt.values())) + 256
    e2e_t = torch.tensor([float(np.mean(e2e_times))], device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t.item())

    # ---- leg 1b: the same, starting one step earlier — from the mapper's raw PAF text ----------
    # upstream: Paf.parse_PAF(StringIO(paf_raw)) builds {read: [PafLine]} in Python before convert_records
    # (mapper.py:63-65); process_batch_text tokenises the text in C and never builds those objects
    txt_times = []
    for it in range(args.warmup + args.steps):
        pd, seqs, rb = batches[it % n_batches]
        barrier()
        t0 = time.perf_counter()
        run.process_batch_text(rb.paf_text, seqs)
        torch.cuda.synchronize()
        if it >= args.warmup:
            txt_times.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    parse_PAF(io.StringIO(batches[0][2].paf_text), min_len=200)
    py_parse_ms = (time.perf_counter() - t0) * 1e3
    txt_t = torch.tensor([float(np.mean(txt_times))], device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(txt_t, op=dist.ReduceOp.MAX)
    txt_s = float(txt_t.item())

    # ---- leg 2: inputs resident in HBM -------------------------------------------------------------------
    dev_batches = []
    for pd, seqs, rb in batches:
        inc = run.cc.convert_records(paf_dict=pd, seqs=seqs)
        packed = run.pack_for_device(inc)

        def to_dev(d):
            return {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in d.items()}
        dev_batches.append([to_dev(d) for d in packed] if isinstance(packed, list) else to_dev(packed))
    fhat_w = run.read_starts.update_f_pointmass()
    upd_kwargs = dict(approx_ccl=run.rl_dist.approx_ccl, time_cost=run.rl_dist.time_cost, bucket_threshold=run.bucket_threshold)
    run.device_update(fhat_windows=fhat_w, **upd_kwargs)        # uploads F-hat once; later calls reuse it

    def step(i):
        d = dev_batches[i % n_batches]
        run.ingest_device(d)
        return run.device_update(fhat_windows=None, **upd_kwargs)

    sampler = ClockSampler(local)
    sampler.start()                     # nvidia-smi needs ~100 ms to deliver its first line: start it ahead of the warm-up
    for i in range(args.warmup):
        step(i)
    l0 = eng.launch_count()
    score_ms, all_ms = [], []
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    mirror_steps = []
    for i in range(args.steps):
        out = step(args.warmup + i)
        t = eng.timing()
        score_ms.append(t["score_bin"])
        all_ms.append(t)
        mirror_steps.append(int(out.mirror_bytes))
    ev1.record()
    barrier()
    launches = eng.launch_count() - l0
    ms = ev0.elapsed_time(ev1) / args.steps
    ms_t = torch.tensor([ms], device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = float(ms_t.item())
    # A timed region of a few milliseconds is shorter than nvidia-smi's 100 ms sampling period: keep the same workload
    # running (untimed; the same number of extra updates on every rank, derived from the agreed ms) until the sampler has
    # seen ~0.5 s of it, so that the clocks line always describes the GPU under this load.
    n_extra = 0
    if ms * (args.steps + args.warmup) < 500.0:
        n_extra = min(int(500.0 / max(ms, 0.05)) + 1, 5000)
        for i in range(n_extra):
            step(args.warmup + args.steps + i)
        barrier()
    clocks = sampler.stop()
    clocks["sampled_over"] = f"warm-up + timed updates + {n_extra} untimed updates of the same workload"

    if rank != 0:
        return
    peaks = {}
    pk = REPO / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic bytes of the score/bin pass on THIS rank: 10 B counters + 1 B reference base per
    # site*barcode (the reference base is re-read per barcode) + 8 B per 100-site bin written
    my_sites = run.local_sites() * spec["nb"]
    alg_bytes = my_sites * 11 + (my_sites // 100) * 8
    if spec["nb"] > 1:
        alg_bytes += run.local_sites() * (10 * spec["nb"] + 4 + 4)     # row-summary pre-pass: counters again + 4 B flag write + read
    k_ms = float(np.mean(score_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    prof = REPO / "profiles" / "score_bin_traffic.json"
    if prof.exists():
        try:
            pj = json.loads(prof.read_text())
            if pj.get("workload") == args.workload and pj.get("sites") == my_sites:
                traffic = pj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    cpu = None
    if not args.no_cpu:
        v, desc, sec = cpu_sample(spec, steps=3, warmup=1)
        cpu = {"value": v / 1e9, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc, "s_per_update_on_sample": sec}
    mean_t = {k: float(np.mean([t[k] for t in all_ms])) for k in all_ms[0]}
    line = {
        "metric": METRIC, "value": total_sites / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["name"], "sites": int(sum(lengths)), "barcodes": spec["nb"], "ploidy": spec["ploidy"],
                   "reads_per_batch": spec["reads"], "l2": "inputs (counters >= 46 MB ... 31 GB) exceed L2; 3 batches cycled",
                   "sharding": f"genome axis split over {world} GPU(s)",
                   "exchange": getattr(run, "exchange_mode", "none")},
        "e2e": {"value": total_sites / e2e_s / 1e9, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h),
                "host_ms": dict(zip(("convert_records", "announce", "ingest", "update_wrapper"),
                                    (float(x) * 1e3 for x in np.mean(np.array(e2e_parts), axis=0)))),
                "update_wrapper_ms": getattr(run, "last_host_ms", None)},
        "e2e_from_paf_text": {"value": total_sites / txt_s / 1e9, "unit": UNIT, "ms_per_step": txt_s * 1e3,
                              "what": "BossRuns.process_batch_text: raw PAF text + read strings -> masks on the host (C tokeniser, no PafLine objects)",
                              "python_parse_PAF_ms_avoided": py_parse_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_score_bin", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)"
                     if pk.exists() else "fallback 6650 GB/s (of fallback)", "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": int(alg_bytes)},
        "kernel_ms": mean_t,
        "kernel_ms_steps": {k: [round(t[k], 3) for t in all_ms] for k in ("score_bin", "smooth", "hist", "distribute", "scatter", "update")},
        "mirror_bytes_steps": mirror_steps,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "setup_s": setup_s,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", default="auto", choices=["auto", "fabric", "phases"],
                    help="N>1: peer-memory fabric inside the update's kernels, or NCCL collectives between phases")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    spec = workload_spec(args.workload, args.scale)
    if args.impl == "reference":
        run_reference_arm(args, spec)
    else:
        run_b200(args, spec)


if __name__ == "__main__":
    main()
