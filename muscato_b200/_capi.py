"""ctypes binding of include/muscato_b200.h (the C ABI a cgo shim would bind too).

There is no fallback: if the shared library is missing or cannot be loaded, importing
the engine raises.  The library is NOT built implicitly at import time on a GPU box;
run `python -c "import __graft_entry__ as g; g.build()"` (or muscato_b200.build.build()).
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

MSC_MAX_WINDOWS = 32
MSC_MAX_WINDOW_WIDTH = 50
MSC_MAX_READ_LENGTH = 1024

MSC_OK, MSC_ERR_CONFIG, MSC_ERR_INPUT, MSC_ERR_CUDA, MSC_ERR_STATE, MSC_ERR_NOMEM, MSC_ERR_IO, MSC_ERR_AGAIN = range(8)
MSC_MATCH_FIRST, MSC_MATCH_BEST = 0, 1
MSC_NO_MATCH = 0x7F7F7F7F
MSC_STAGE_SCREEN, MSC_STAGE_CONFIRM, MSC_STAGE_COMBINE, MSC_STAGE_DEFER = 1, 2, 4, 8


class msc_config(C.Structure):
    _fields_ = [
        ("n_windows", C.c_int32),
        ("windows", C.c_int32 * MSC_MAX_WINDOWS),
        ("window_width", C.c_int32),
        ("max_read_length", C.c_int32),
        ("pmatch", C.c_double),
        ("min_dinuc", C.c_int32),
        ("mmtol", C.c_int32),
        ("max_matches", C.c_int64),
        ("match_mode", C.c_int32),
        ("device", C.c_int32),
        ("bloom_bits_per_key", C.c_int32),
        ("keep_ascii", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


class msc_match(C.Structure):
    _fields_ = [("read_id", C.c_uint32), ("gene_id", C.c_uint32), ("pos", C.c_uint32), ("nx", C.c_uint32)]


class msc_key_rec(C.Structure):
    _fields_ = [("window", C.c_uint32), ("read_id", C.c_uint32)]


class msc_cand_rec(C.Structure):
    _fields_ = [("gene_id", C.c_uint32), ("p", C.c_uint32), ("read_id", C.c_uint32), ("window", C.c_uint32)]


class msc_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_reads", "n_keys", "n_key_groups", "table_slots", "bloom_bytes",
        "n_targets", "target_bases", "positions_probed",
        "n_candidates", "n_pairs", "n_pass", "n_matches_pre", "n_matches",
        "n_overflow_groups", "h2d_bytes", "d2h_bytes", "kernel_launches")] + [(n, C.c_float) for n in (
        "ms_pack_reads", "ms_build", "ms_pack_targets", "ms_scan", "ms_expand", "ms_confirm", "ms_combine",
        "ms_scan_kernel")] + [("reserved_f", C.c_float * 8)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved_f"}
        d["bloom_pass"] = float(self.reserved_f[0])
        d["ms_prep"] = float(self.reserved_f[1])  # device time of msc_prep_reads (sort + collapse), H2D excluded
        d["front_passes"] = int(self.reserved_f[2])  # exact front: scan launches per step; 0 = Bloom front
        return d


# Every symbol include/muscato_b200.h declares: (name, restype, argtypes)
_SIGS = [
    ("msc_version", C.c_char_p, []),
    ("msc_struct_size", C.c_uint64, [C.c_int]),
    ("msc_create", C.c_void_p, [C.POINTER(msc_config), C.c_char_p, C.c_uint64]),
    ("msc_destroy", None, [C.c_void_p]),
    ("msc_last_error", C.c_char_p, [C.c_void_p]),
    ("msc_set_reads", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    ("msc_set_reads_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    ("msc_prep_reads", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("msc_fetch_read_groups", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("msc_unique_reads_bytes", C.c_uint64, [C.c_void_p]),
    ("msc_fetch_unique_reads", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("msc_set_targets", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    ("msc_set_targets_packed", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    ("msc_set_targets_from", C.c_int, [C.c_void_p, C.c_void_p]),
    ("msc_packed_target_words", C.c_uint64, [C.c_void_p]),
    ("msc_fetch_packed_targets", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    ("msc_rebuild", C.c_int, [C.c_void_p, C.c_int]),
    ("msc_screen", C.c_int, [C.c_void_p]),
    ("msc_confirm", C.c_int, [C.c_void_p]),
    ("msc_best_device", C.c_void_p, [C.c_void_p]),
    ("msc_stream", C.c_void_p, [C.c_void_p]),
    ("msc_combine", C.c_int, [C.c_void_p]),
    ("msc_matches_device", C.c_void_p, [C.c_void_p, C.POINTER(C.c_uint64)]),
    ("msc_fetch_matches", C.c_int, [C.c_void_p, C.POINTER(C.POINTER(msc_match)), C.POINTER(C.c_uint64)]),
    ("msc_fetch_matches_into", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    ("msc_fetch_nonmatch", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    ("msc_run", C.c_int, [C.c_void_p]),
    ("msc_rebuild_and_run", C.c_int, [C.c_void_p, C.c_int]),
    ("msc_run_stages", C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    ("msc_set_shards", C.c_int, [C.c_void_p, C.c_int32]),
    ("msc_shard_overflow", C.c_int, [C.c_void_p]),
    ("msc_overflow_keys", C.c_int, [C.c_void_p, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]),
    ("msc_diverted_record_bytes", C.c_uint32, [C.c_void_p]),
    ("msc_divert_groups", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                    C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_uint64)]),
    ("msc_replay_diverted", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.POINTER(msc_match)),
                                      C.POINTER(C.c_uint64)]),
    ("msc_set_stage_timing", C.c_int, [C.c_void_p, C.c_int]),
    ("msc_get_stats", C.c_int, [C.c_void_p, C.POINTER(msc_stats)]),
    ("msc_reset_stats", None, [C.c_void_p]),
    ("msc_free", None, [C.c_void_p]),
    ("msc_dump_keys", C.c_int, [C.c_void_p, C.POINTER(C.POINTER(msc_key_rec)), C.POINTER(C.c_uint64)]),
    ("msc_dump_candidates", C.c_int, [C.c_void_p, C.POINTER(C.POINTER(msc_cand_rec)), C.POINTER(C.c_uint64)]),
]

EXPORTED_SYMBOLS = [s[0] for s in _SIGS]

_lib = None


def load() -> C.CDLL:
    """Load libmuscato_b200.so (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
            "(nvcc, sm_100a).  muscato_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in _SIGS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
