"""The C-ABI boundary without a GPU: the library builds, loads, exports every symbol include/bossgpu.h declares,
its struct layouts match the ctypes mirrors, its pure-host entry points work, and the product path fails loudly
(instead of falling back to anything) when there is no CUDA device."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from boss_runs_b200 import _lib

REPO = Path(__file__).resolve().parent.parent
HEADER = (REPO / "include" / "bossgpu.h").read_text()


def declared_symbols() -> list[str]:
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(bossgpu_[a-z0-9_]+)\s*\(", body)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/bossgpu.h but not exported by libbossgpu.so"
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding and header disagree"
    assert lib.bossgpu_abi_version() == _lib.ABI_VERSION == int(re.search(r"BOSSGPU_ABI_VERSION\s+(\d+)", HEADER).group(1))


def test_header_constants_match_binding():
    def const(name):
        return int(re.search(rf"#define\s+{name}\s+(-?\d+)", HEADER).group(1))
    assert (const("BOSSGPU_BIN"), const("BOSSGPU_BUCKET"), const("BOSSGPU_RSD_WINDOW"), const("BOSSGPU_FREEZE")) == \
        (_lib.BIN, _lib.BUCKET, _lib.RSD_WINDOW, _lib.FREEZE)
    assert (const("BOSSGPU_N_PATTERNS"), const("BOSSGPU_N_STEPS"), const("BOSSGPU_HIST_BINS"), const("BOSSGPU_N_TIMERS")) == \
        (_lib.N_PATTERNS, _lib.N_STEPS, _lib.HIST_BINS, _lib.N_TIMERS)
    for name, val in (("OK", 0), ("EINVAL", -1), ("ECUDA", -2), ("ENOMEM", -3), ("EBASE", -4), ("ESHAPE", -5), ("ESTATE", -6),
                      ("EEMPTY", -7), ("EPEER", -8), ("ENOTC", -9)):
        assert const(f"BOSSGPU_{name}") == getattr(_lib, name) == val


def test_struct_layouts(tmp_path):
    """sizeof/offsetof of every struct crossing the boundary, as gcc sees the header vs the ctypes mirrors."""
    import subprocess
    src = tmp_path / "layout.c"
    fields = {
        "bossgpu_segment": ["contig", "contig_len", "start", "len"],
        "bossgpu_config": ["abi_version", "device", "stream", "n_segments", "n_barcodes", "segments", "ref_codes",
                           "n_contigs_total", "halo_bins", "contig_len_all", "n_sites_total", "n_windows_total", "len_g", "phi",
                           "priors", "phi_pow", "score0_contig", "entropy0_contig"],
        "bossgpu_update_params": ["w", "mult", "tc", "bucket_threshold", "fhat_windows", "write_debug", "fhat_from_counts",
                                  "rs_alpha", "rs_denom", "rs_zero_value"],
        "bossgpu_update_result": ["switched_on", "strat_size", "threshold", "normaliser", "ubar0", "fhat_sum", "n_nonzero",
                                  "n_dropout", "n_accept", "mirror_bytes"],
        "bossgpu_aeons_params": ["mu_ds", "ccl_ds", "perc", "tc", "tbar0", "want_strategy"],
        "bossgpu_aeons_result": ["threshold", "normaliser", "ubar0", "n_nonzero", "strat_size"],
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{REPO}/include/bossgpu.h"', "int main(void){"]
    for st, fs in fields.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fs:
            lines.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines.append("return 0;}")
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)     # also: the header is plain C
    got = dict(l.rsplit(" ", 1) for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirror = {"bossgpu_segment": _lib.Segment, "bossgpu_config": _lib.Config, "bossgpu_update_params": _lib.UpdateParams,
              "bossgpu_update_result": _lib.UpdateResult, "bossgpu_aeons_params": _lib.AeonsParams,
              "bossgpu_aeons_result": _lib.AeonsResult}
    for st, fs in fields.items():
        assert int(got[st]) == C.sizeof(mirror[st]), st
        for f in fs:
            assert int(got[f"{st}.{f}"]) == getattr(mirror[st], f).offset, f"{st}.{f}"


def test_tokenizer_follows_upstream_regex(lib):
    """`(\\d+)([MIDNSHP=XB])` with findall semantics (sequences.py:672,768): unmatched text is skipped."""
    def tok(text: bytes, cap=64):
        out = np.zeros(cap, dtype=np.uint32)
        r, q = C.c_int64(), C.c_int64()
        k = lib.bossgpu_tokenize_cigar(text, len(text), out.ctypes.data, cap, C.byref(r), C.byref(q))
        return k, [(int(x >> 4), int(x & 15)) for x in out[:max(k, 0)]], r.value, q.value

    assert tok(b"10M2I3D5M") == (4, [(10, 0), (2, 1), (3, 2), (5, 0)], 18, 17)
    assert tok(b"5S10=2X1N") == (4, [(5, 0), (10, 0), (2, 0), (1, 0)], 18, 18)          # S,=,X,N are kept columns upstream
    assert tok(b"") == (0, [], 0, 0)
    assert tok(b"M5M") == (1, [(5, 0)], 5, 5)                                            # a letter without digits matches nothing
    assert tok(b"12Q7M") == (1, [(7, 0)], 7, 7)                                          # unknown letter: run is dropped
    k, *_ = tok(b"1M" * 10, cap=4)
    assert k < 0 and b"capacity" in lib.bossgpu_last_error()
    import re as _re
    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(1, 30))
        text = "".join(f"{int(rng.integers(1, 3000))}{'MIDNSHP=XB'[int(rng.integers(0, 10))]}" for _ in range(n))
        k, ops, r, q = tok(text.encode(), cap=64)
        want = [(int(a), {"I": 1, "D": 2}.get(b, 0)) for a, b in _re.findall(r"(\d+)([MIDNSHP=XB])", text)]
        assert ops == want and k == len(want)
        assert r == sum(a for a, c in want if c != 1) and q == sum(a for a, c in want if c != 2)


def test_pattern_rank_matches_oracle(lib):
    from oracle import boss_oracle as bo
    rng = np.random.default_rng(1)
    pats = bo.all_patterns()[rng.integers(0, bo.N_PATTERNS, size=2000)].astype(np.uint16)
    want = bo.pattern_rank(pats)
    for p, w in zip(pats, want):
        c = np.ascontiguousarray(p)
        assert lib.bossgpu_pattern_rank(c.ctypes.data) == w
    frozen = np.array([30, 0, 0, 0, 0], dtype=np.uint16)
    assert lib.bossgpu_pattern_rank(frozen.ctypes.data) == -1
    assert lib.bossgpu_pattern_rank(None) == -1


def test_error_mapping():
    """BOSSGPU_E* -> the exception the reference raises at the same place."""
    _lib.load()
    for code, exc in ((_lib.EBASE, IndexError), (_lib.ESHAPE, AssertionError), (_lib.EEMPTY, ValueError), (_lib.EINVAL, ValueError),
                      (_lib.ENOMEM, MemoryError), (_lib.ECUDA, _lib.BossGpuError), (_lib.ESTATE, _lib.BossGpuError)):
        with pytest.raises(exc):
            _lib.check(code)
    _lib.check(_lib.OK)


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run; nothing under boss_runs_b200 imports the oracle."""
    import torch
    for py in (REPO / "boss_runs_b200").glob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{py.name} imports the oracle"
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from boss_runs_b200.runs import BossRuns
    from boss_runs_b200 import synth
    with pytest.raises(_lib.BossGpuError, match="no usable CUDA device|CUDA"):
        BossRuns(contigs=synth.random_contigs({"c1": 100_000}, seed=1), ploidy=1)


def test_missing_library_is_loud(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libbossgpu.so")
    with pytest.raises(_lib.BossGpuError, match="no CPU fallback"):
        _lib.load()


def test_worker_pool_runs_every_task_and_survives_reuse(tmp_path):
    """csrc/workerpool.h (the ingest path's persistent host threads) on its own: every worker index runs exactly once per
    start(), pools of different sizes follow each other, a forked child gets fresh workers."""
    src = tmp_path / "pool.cpp"
    src.write_text(r'''
#include "workerpool.h"
#include <atomic>
#include <cstdio>
#include <sys/wait.h>
int main() {
    boss::WorkerPool& p = boss::WorkerPool::instance();
    const int sizes[] = {3, 15, 1, 8, 15, 2};
    for (int rep = 0; rep < 200; ++rep) {
        const int n = sizes[rep % 6];
        std::atomic<int> hits[16];
        for (auto& h : hits) h = 0;
        const std::function<void(int)> job = [&](int t) { hits[t].fetch_add(1); };
        std::unique_lock<std::mutex> own(p.owner, std::try_to_lock);
        if (!own.owns_lock()) return 2;
        p.start(n, job);
        p.wait();
        for (int t = 0; t < 16; ++t) if (hits[t] != (t < n ? 1 : 0)) { std::printf("rep %d worker %d ran %d times\n", rep, t, (int)hits[t]); return 1; }
    }
    const pid_t c = fork();
    if (c == 0) {
        std::atomic<int> sum{0};
        const std::function<void(int)> job = [&](int t) { sum.fetch_add(t + 1); };
        p.start(4, job);
        p.wait();
        _exit(sum == 10 ? 0 : 3);
    }
    int st = 0;
    waitpid(c, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { std::printf("child status %d\n", st); return 4; }
    std::puts("ok");
    return 0;
}
''')
    exe = tmp_path / "pool"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", str(REPO / "boss_runs_b200" / "csrc"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)
