"""Build recipe for libbossgpu.so (in-tree, sm_100a only).

`python -m boss_runs_b200.build` or `__graft_entry__.build()` compiles `csrc/bossgpu.cu` (which pulls in
every kernel header) with nvcc into `boss_runs_b200/libbossgpu.so`. nvcc cross-compiles without a GPU.
The built library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libbossgpu.so"
STAMP = PKG / ".libbossgpu.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-pthread",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libbossgpu.so cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                  + [PKG.parent / "include" / "bossgpu.h"])


def _digest() -> str:
    h = hashlib.sha256()
    for p in sources():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile if sources changed since the last build; returns the path of the shared library."""
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / "bossgpu.cu"), "-o", str(LIB)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libbossgpu.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    STAMP.write_text(digest)
    return LIB


def fastconv_path() -> Path:
    import sysconfig
    return PKG / ("_fastconv" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_fastconv(force: bool = False) -> Path:
    """Compile csrc/fastconv.c (CPython helper walking the batch's Python objects; plain gcc, no CUDA)."""
    import sysconfig
    src, out = CSRC / "fastconv.c", fastconv_path()
    stamp = PKG / ".fastconv.stamp"
    digest = hashlib.sha256(src.read_bytes()).hexdigest()
    if not force and out.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return out
    cc = os.environ.get("CC") or shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler for _fastconv")
    cmd = [cc, "-O2", "-fPIC", "-shared", "-Wall", f"-I{sysconfig.get_paths()['include']}", str(src), "-o", str(out)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("building _fastconv failed")
    stamp.write_text(digest)
    return out


if __name__ == "__main__":
    build_fastconv(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
