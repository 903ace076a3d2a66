"""boss_runs_b200 — B200 (sm_100a) implementation of BOSS-RUNS' periodic strategy update.

Layout:
  csrc/            CUDA kernels + the C ABI of libbossgpu.so (declared in include/bossgpu.h)
  build.py         nvcc recipe (in-tree build)
  _lib.py          ctypes binding, error mapping
  engine.py        `Engine`: one GPU shard behind the C ABI
  priors.py        host constants of the scoring model (mirror of upstream `Priors`)
  hostmodel.py     host mirrors: PafLine/parse_PAF, ReadlengthDist, ReadStartDist
  runs.py          reference-facing API: Contig, Reference, CoverageConverter, BossRuns
  dropin.py        upstream's own `BossRuns` / `BossRunsSim` with the array half replaced (mixin; `[gpu]` TOML table; main())
  simulation.py    the simulator's decision step on the GPU-backed update
  sharding.py      multi-GPU: genome-axis partition, peer-memory fabric / NCCL exchange steps, routed batches
  aeons.py         BOSS-AEONS' Benefit / pool threshold / masks (stateless C-ABI call)
  stratfile.py     optional packed strategy file boss.bits + consumer lookup
  synth.py         synthetic references / read batches of BASELINE.json's shapes (tests, bench)

Importing the package never touches CUDA; the first `Engine` loads libbossgpu.so and fails loudly if it
is missing or finds no device. There is no CPU fallback.
"""
__version__ = "0.1.0"
