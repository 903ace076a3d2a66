"""ctypes binding of libbossgpu.so (the C ABI declared in include/bossgpu.h).

There is deliberately no fallback: if the library is missing, cannot be loaded, or finds no CUDA
device, the error propagates.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libbossgpu.so"

ABI_VERSION = 1
BIN, BUCKET, RSD_WINDOW, FREEZE = 100, 20_000, 2000, 30
N_PATTERNS, N_STEPS, HIST_BINS, N_TIMERS = 278_256, 10, 1088, 8

OK, EINVAL, ECUDA, ENOMEM, EBASE, ESHAPE, ESTATE, EEMPTY, EPEER, ENOTC = 0, -1, -2, -3, -4, -5, -6, -7, -8, -9
BUF_SWITCH, BUF_NORM, BUF_HIST, BUF_MASK, BUF_HALO_SEND, BUF_HALO_RECV, BUF_STRAT, BUF_COV_TOTAL = range(8)


class BossGpuError(RuntimeError):
    """CUDA / resource failure inside libbossgpu."""


class PeerTimeout(BossGpuError):
    """A peer shard did not reach an exchange step of the sharded update (BOSSGPU_EPEER)."""


class Segment(C.Structure):
    _fields_ = [("contig", C.c_int32), ("reserved", C.c_int32), ("contig_len", C.c_int64),
                ("start", C.c_int64), ("len", C.c_int64)]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("stream", C.c_void_p),
        ("n_segments", C.c_int32), ("n_barcodes", C.c_int32),
        ("segments", C.POINTER(Segment)), ("ref_codes", C.c_void_p),
        ("n_contigs_total", C.c_int32), ("halo_bins", C.c_int32),
        ("contig_len_all", C.c_void_p), ("n_sites_total", C.c_int64), ("n_windows_total", C.c_int64),
        ("len_g", C.c_int32), ("reserved2", C.c_int32),
        ("phi", C.c_void_p), ("priors", C.c_void_p), ("phi_pow", C.c_void_p),
        ("score0_contig", C.c_double), ("entropy0_contig", C.c_double),
    ]


class UpdateParams(C.Structure):
    _fields_ = [("w", C.c_int32 * N_STEPS), ("mult", C.c_double * N_STEPS), ("tc", C.c_double),
                ("bucket_threshold", C.c_double), ("fhat_windows", C.c_void_p),
                ("write_debug", C.c_int32), ("fhat_from_counts", C.c_int32),
                ("rs_alpha", C.c_double), ("rs_denom", C.c_double), ("rs_zero_value", C.c_double)]


class UpdateResult(C.Structure):
    _fields_ = [("switched_on", C.c_int32), ("strat_size", C.c_int32), ("threshold", C.c_double),
                ("normaliser", C.c_double), ("ubar0", C.c_double), ("fhat_sum", C.c_double),
                ("n_nonzero", C.c_int64), ("n_dropout", C.c_int64), ("n_accept", C.c_int64 * 2), ("mirror_bytes", C.c_int64)]


class AeonsParams(C.Structure):
    _fields_ = [("mu_ds", C.c_int32), ("ccl_ds", C.c_int32 * 10), ("perc", C.c_double * 10), ("tc", C.c_double),
                ("tbar0", C.c_double), ("want_strategy", C.c_int32), ("reserved", C.c_int32)]


class AeonsResult(C.Structure):
    _fields_ = [("threshold", C.c_double), ("normaliser", C.c_double), ("ubar0", C.c_double), ("n_nonzero", C.c_int64),
                ("strat_size", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/bossgpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "bossgpu_abi_version": (C.c_int, []),
    "bossgpu_last_error": (C.c_char_p, []),
    "bossgpu_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "bossgpu_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "bossgpu_destroy": (C.c_int, [_P]),
    "bossgpu_synchronize": (C.c_int, [_P]),
    "bossgpu_ingest_packed": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "bossgpu_ingest_records": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int]),
    "bossgpu_prescore_begin": (C.c_int, [_P]),
    "bossgpu_prescore": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "bossgpu_ingest_records_ptr": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int]),
    "bossgpu_ingest_records_routed": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int]),
    "bossgpu_strat_host": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "bossgpu_buckets_host": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "bossgpu_set_strat_mirror": (C.c_int, [_P, _P, C.c_int64, C.c_int]),
    "bossgpu_host_register": (C.c_int, [_P, C.c_int64]),
    "bossgpu_host_unregister": (C.c_int, [_P]),
    "bossgpu_get_seg_accept": (C.c_int, [_P, _P, C.c_int64]),
    "bossgpu_read_starts_add": (C.c_int, [_P, C.c_int64, _P, _P]),
    "bossgpu_get_read_starts": (C.c_int, [_P, _P, C.c_int64]),
    "bossgpu_tokenize_cigar": (C.c_int64, [C.c_char_p, C.c_int64, _P, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bossgpu_update": (C.c_int, [_P, C.POINTER(UpdateParams), C.POINTER(UpdateResult)]),
    "bossgpu_update_phase": (C.c_int, [_P, C.c_int, C.POINTER(UpdateParams), C.POINTER(UpdateResult)]),
    "bossgpu_set_shards": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "bossgpu_halo_pack": (C.c_int, [_P]),
    "bossgpu_halo_unpack": (C.c_int, [_P]),
    "bossgpu_exchange_buffer": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "bossgpu_fabric_info": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t), _P]),
    "bossgpu_ipc_open": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "bossgpu_ipc_close": (C.c_int, [C.c_int, _P]),
    "bossgpu_fabric_attach": (C.c_int, [_P, C.c_int32, _P, C.c_double]),
    "bossgpu_update_fused_begin": (C.c_int, [_P, C.POINTER(UpdateParams)]),
    "bossgpu_update_fused_end": (C.c_int, [_P, C.POINTER(UpdateResult)]),
    "bossgpu_get_strat": (C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    "bossgpu_get_strat_all": (C.c_int, [_P, _P, C.c_int64]),
    "bossgpu_get_strat_packed": (C.c_int, [_P, _P, C.c_int64]),
    "bossgpu_strat_rows": (C.c_int64, [_P, C.c_int32]),
    "bossgpu_get_coverage": (C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    "bossgpu_set_coverage": (C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    "bossgpu_get_scores": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int64]),
    "bossgpu_get_scores_ds": (C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    "bossgpu_get_benefit": (C.c_int, [_P, C.c_int32, _P, _P, _P, C.c_int64]),
    "bossgpu_get_buckets": (C.c_int, [_P, C.c_int32, _P, C.c_int64, _P]),
    "bossgpu_set_buckets": (C.c_int, [_P, C.c_int32, _P, C.c_int64]),
    "bossgpu_get_hist": (C.c_int, [_P, _P, _P]),
    "bossgpu_get_fhat": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "bossgpu_get_score_table": (C.c_int, [_P, _P, _P]),
    "bossgpu_pattern_rank": (C.c_int64, [_P]),
    "bossgpu_timing": (C.c_int, [_P, _P]),
    "bossgpu_launch_count": (C.c_int64, [_P]),
    "bossgpu_ingest_bytes": (C.c_int64, [_P]),
    "bossgpu_aeons_update": (C.c_int, [C.c_int, C.c_int64, _P, _P, _P, _P, C.POINTER(AeonsParams), _P, _P, _P, _P, C.POINTER(AeonsResult)]),
    "bossgpu_synth_coverage": (C.c_int, [_P, C.c_uint64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
}

_lib = None


def load() -> C.CDLL:
    """Load libbossgpu.so (once). Raises if it has not been built — there is no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise BossGpuError(
            f"{LIB_PATH} is missing. Build it with `python -m boss_runs_b200.build` "
            "(needs nvcc). boss_runs_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.bossgpu_abi_version() != ABI_VERSION:
        raise BossGpuError("libbossgpu.so ABI version mismatch; rebuild with `python -m boss_runs_b200.build --force`")
    _lib = lib
    return lib


def last_error() -> str:
    return load().bossgpu_last_error().decode(errors="replace")


def check(rc: int) -> None:
    """Map a BOSSGPU_E* code onto the exception type the reference raises at the same place."""
    if rc == OK:
        return
    msg = last_error()
    if rc == EBASE:
        raise IndexError(msg)            # np.add.at index out of bounds: reference.py:138-140
    if rc == ESHAPE:
        raise AssertionError(msg)        # sequences.py:732-733 (or NumPy's shape error at :785)
    if rc == EEMPTY:
        raise ValueError(msg)            # np.max of an empty array: sequences.py:588
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == ENOTC:
        raise AttributeError("'ReadlengthDist' object has no attribute 'time_cost'")     # readlengthdist.py:68, core.py:192 (Q14)
    if rc == EPEER:
        raise PeerTimeout(msg)
    raise BossGpuError(f"libbossgpu error {rc}: {msg}")


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


# pointer to the UTF-8 bytes of a str object (for ASCII strings: its own buffer, no copy); valid while the object lives
_as_utf8 = C.pythonapi.PyUnicode_AsUTF8AndSize
_as_utf8.restype = C.c_void_p
_as_utf8.argtypes = [C.py_object, C.c_void_p]


def str_ptr(s: str) -> int:
    return _as_utf8(s, None)
