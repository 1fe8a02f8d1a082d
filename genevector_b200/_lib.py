"""ctypes binding of libgvb200.so (C ABI declared in include/gvb200.h).

This is the only place the shared library is loaded.  There is no CPU fallback: if the
library is missing the import fails with a build hint, and every compute entry point of the
library itself refuses to run on a device that is not sm_100 (GV_ERR_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GVB200_LIB selects another build of the same ABI (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("GVB200_LIB") or os.path.join(_HERE, "libgvb200.so")

GV_OK = 0
GV_F32, GV_F64, GV_COUNTS_U8_COL16 = 0, 1, 2
GV_MOM_SIGN, GV_MOM_PEARSON, GV_MOM_COSINE, GV_MOM_JACCARD = 0, 1, 2, 3
GV_PATH_AUTO, GV_PATH_COUNTS, GV_PATH_GENERAL = 0, 1, 2
GV_VAL_COUNTS, GV_VAL_DETECTED, GV_VAL_FIXED_POINT, GV_VAL_RANKS = 0, 1, 2, 3
GV_MAX_BINS = 31
GV_MARG_STRIDE = 32
GV_EDGE_STRIDE = 32
GV_EDGE_SCALE_SLOT = 31
GV_EDGE_OFFSET_SLOT = 30
GV_QTABLE_DIM = 32
GV_ABI_VERSION = 6
GV_OPERAND_INT8, GV_OPERAND_FP4 = 0, 1
GV_ERR_TIMEOUT = -5


class GvError(RuntimeError):
    """A libgvb200 call returned a negative status."""

    def __init__(self, fn: str, status: int, message: str):
        super().__init__(f"{fn} failed with status {status}: {message}")
        self.status = status


_vp, _i32, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

# name -> (restype, argtypes); mirrors include/gvb200.h one to one
SIGNATURES = {
    "gv_abi_version": (_i32, []),
    "gv_last_error": (C.c_char_p, []),
    "gv_device_check": (_i32, []),
    "gv_discretize_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "gv_discretize_csr": (_i32, [_vp, _i32, _vp, _vp, _i64, _i32, _i64, _i32, _vp, _vp, _i64,
                                 _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp,
                                 _vp, _sz, _i32, _vp, _vp]),
    "gv_csr_check": (_i32, [_vp, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "gv_shard_transpose": (_i32, [_vp, _i32, _vp, _vp, _i64, _i32, _i64, _i64, _vp, _i32, _i32, _i64, _vp, _vp, _vp,
                                  _i32, _vp]),
    "gv_shard_bins": (_i32, [_i32, _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp,
                             _vp, _vp, _i32, _i64, _i32, _vp]),
    "gv_gene_entropy": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _i32, _vp, _vp, _sz, _vp]),
    "gv_rows_per_gene": (_i32, [_i32]),
    "gv_onehot_rows": (_i64, [_i32, _i32]),
    "gv_value_rows": (_i64, [_i32, _i32]),
    "gv_expand_onehot": (_i32, [_vp, _i64, _i32, _i32, _vp, _i64, _i32, _vp]),
    "gv_tile_m": (_i32, [_i32]),
    "gv_tile_schedule": (_i64, [_i64, _i32, _i32, _vp, _i64]),
    "gv_mi_tiles": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _i64, _i32, _vp, _i64, _vp, _i64,
                           _i32, _i32, _vp, _vp]),
    "gv_gram_dump": (_i32, [_vp, _i64, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp]),
    "gv_tile_schedule_rect": (_i64, [_i64, _i64, _i32, _vp, _i64]),
    "gv_graph_aggregate": (_i32, [_vp, _i64, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _i64, _i32,
                                  _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "gv_xcorr_tiles": (_i32, [_vp, _i64, _i32, _vp, _i64, _i32, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp,
                              _i64, _i32, _vp, _i64, _vp, _i64, _i32, _vp, _vp]),
    "gv_symmetrize": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp]),
    "gv_mma_peak": (_i32, [_i32, _i32, _i32, _i32, _vp, _vp]),
    "gv_moment_tiles": (_i32, [_i32, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp,
                               _i64, _vp, _i64, _vp, _i64, _i32, _vp, _i32, _vp]),
    "gv_compute_mi_pairs": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "gv_build_training_tensors": (_i32, [_vp, _i64, _i32, C.c_float, _i32, _i32, _vp, _vp, _vp, _vp]),
    "gv_upload_pageable": (_i32, [_vp, _vp, _sz, _i32, _vp]),
    "gv_upload_packed_counts": (_i32, [_vp, _vp, _vp, _vp, _sz, _i32, _vp, _vp]),
    "gv_pack_counts_host": (_i32, [_vp, _vp, _sz, _vp, _vp, _vp]),
    "gv_upload_registered": (_i32, [_vp, _vp, _sz, _vp, C.POINTER(_vp)]),
    "gv_upload_release": (_i32, [_vp]),
    "gv_runtime_shutdown": (_i32, []),
    "gv_shm_open": (_i32, [C.c_char_p, _sz, _i32, _i32, C.POINTER(_vp)]),
    "gv_shm_close": (_i32, [_vp, _sz, _i32, C.c_char_p]),
    "gv_flag_store": (_i32, [_vp, C.c_uint64]),
    "gv_flag_load": (C.c_uint64, [_vp]),
    "gv_flag_wait_ge": (_i32, [_vp, C.c_uint64, _i32]),
    "gv_copy_rect_d2h": (_i32, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "gv_ipc_alloc": (_i32, [_sz, C.POINTER(_vp), _vp]),
    "gv_ipc_open": (_i32, [_vp, C.POINTER(_vp)]),
    "gv_ipc_close": (_i32, [_vp]),
    "gv_ipc_free": (_i32, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load libgvb200.so once; raise ImportError with the build command if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no fallback. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C genevector_b200/csrc`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.gv_abi_version() != GV_ABI_VERSION:
        raise ImportError("libgvb200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(fn_name: str, status: int) -> int:
    if status < 0:
        raise GvError(fn_name, int(status), load().gv_last_error().decode("utf-8", "replace"))
    return status


def call(fn_name: str, *args) -> int:
    """Call an int-returning entry point and raise GvError on a negative status."""
    return check(fn_name, getattr(load(), fn_name)(*args))
