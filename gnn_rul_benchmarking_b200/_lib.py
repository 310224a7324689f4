"""ctypes binding of libstgconv_b200.so (the C ABI in include/stgconv_b200.h).

There is no CPU fallback: importing this module builds the library when it is missing or
stale (nvcc is in the image) and raises if it cannot be loaded.  Calls that need a GPU raise
RuntimeError from the status code when none is present.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

c_float_p = C.POINTER(C.c_float)
c_double_p = C.POINTER(C.c_double)


class StgBlockDesc(C.Structure):
    """struct stg_block_desc (include/stgconv_b200.h)."""
    _fields_ = [
        ("H", C.c_int32), ("w", C.c_int32), ("stride", C.c_int32), ("decay", C.c_float),
        ("Wm", C.c_void_p), ("bm", C.c_void_p), ("bn0_w", C.c_void_p), ("bn0_b", C.c_void_p),
        ("bn0_rm", C.c_void_p), ("bn0_rv", C.c_void_p), ("Wt", C.c_void_p), ("bt", C.c_void_p),
        ("bn1_w", C.c_void_p), ("bn1_b", C.c_void_p), ("bn1_rm", C.c_void_p), ("bn1_rv", C.c_void_p),
        ("out", C.c_void_p), ("out_bstride", C.c_int64), ("yp", C.c_void_p), ("stats", C.c_void_p),
    ]


class StgBlockGrads(C.Structure):
    """struct stg_block_grads."""
    _fields_ = [
        ("dout", C.c_void_p), ("dout_bstride", C.c_int64),
        ("dWm", C.c_void_p), ("dbm", C.c_void_p), ("dbn0_w", C.c_void_p), ("dbn0_b", C.c_void_p),
        ("dWt", C.c_void_p), ("dbt", C.c_void_p), ("dbn1_w", C.c_void_p), ("dbn1_b", C.c_void_p),
        ("dxp", C.c_void_p),
    ]


MAX_BLOCKS = 2      # STG_MAX_BLOCKS


class StgModelDims(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("T", C.c_int32), ("P", C.c_int32), ("K", C.c_int32),
                ("EH", C.c_int32), ("E", C.c_int32), ("H", C.c_int32),
                ("w", C.c_int32 * MAX_BLOCKS), ("stride", C.c_int32 * MAX_BLOCKS),
                ("decay", C.c_float), ("pe_dropout", C.c_float), ("bn_momentum", C.c_float), ("bn_eps", C.c_float)]


class StgBN(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p)]


class StgModelBlock(C.Structure):
    _fields_ = [("Wm", C.c_void_p), ("bm", C.c_void_p), ("bn0", StgBN), ("Wt", C.c_void_p), ("bt", C.c_void_p),
                ("bn1", StgBN)]


class StgModelParams(C.Structure):
    _fields_ = [("conv1_w", C.c_void_p), ("bn1", StgBN), ("conv2_w", C.c_void_p), ("bn2", StgBN),
                ("lin_w", C.c_void_p), ("lin_b", C.c_void_p), ("bn3", StgBN), ("pe", C.c_void_p),
                ("blk", StgModelBlock * MAX_BLOCKS), ("fc_w", C.c_void_p * 4), ("fc_b", C.c_void_p * 4)]


class StgTcnParams(C.Structure):
    _fields_ = [("conv1_w", C.c_void_p), ("bn1", StgBN), ("conv2_w", C.c_void_p), ("bn2", StgBN)]


class StgDropout(C.Structure):
    _fields_ = [("keep", C.c_void_p), ("seed", C.c_uint64), ("step_dev", C.c_void_p)]


MODEL_SIGNATURES = {
    "stg_model_workspace_bytes": (C.c_size_t, [C.POINTER(StgModelDims)]),
    "stg_model_forward": (C.c_int, [C.POINTER(StgModelDims), C.POINTER(StgModelParams), C.c_void_p, C.c_void_p,
                                    C.c_size_t, C.c_int, C.POINTER(StgDropout), C.c_void_p, C.c_void_p]),
    "stg_model_backward": (C.c_int, [C.POINTER(StgModelDims), C.POINTER(StgModelParams), C.POINTER(StgModelParams),
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(StgDropout), C.c_void_p,
                                     C.c_void_p]),
    "stg_model_loss_backward": (C.c_int, [C.POINTER(StgModelDims), C.POINTER(StgModelParams),
                                          C.POINTER(StgModelParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.POINTER(StgDropout), C.c_void_p, C.c_void_p, C.c_void_p]),
    "stg_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float,
                                C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
}

# name -> (restype, argtypes); every symbol include/stgconv_b200.h declares
SIGNATURES = {
    "stg_last_error": (C.c_char_p, []),
    "stg_version": (C.c_char_p, []),
    "stg_block_xmoments": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "stg_block_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(StgBlockDesc), C.c_int,
                                    C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]),
    "stg_block_backward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(StgBlockDesc),
                                     C.POINTER(StgBlockGrads), C.c_int, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "stg_profile_enable": (C.c_int, [C.c_int]),
    "stg_profile_reset": (C.c_int, []),
    "stg_profile_slots": (C.c_int, []),
    "stg_profile_name": (C.c_char_p, [C.c_int]),
    "stg_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}
MODEL_SIGNATURES["stg_allreduce_adam"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                    C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                                    C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p])
MODEL_SIGNATURES["stg_metrics"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_adj_forward"] = (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_adj_backward"] = (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                                  C.c_int, C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_agg_forward"] = (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_void_p])
MODEL_SIGNATURES["stg_agg_backward"] = (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_tcn_forward"] = (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(StgTcnParams),
                                                 C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p])
MODEL_SIGNATURES["stg_tcn_backward"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.POINTER(StgTcnParams), C.POINTER(StgTcnParams), C.c_float,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_patch_stats"] = (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_patch_stats11"] = (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_patch_stats12"] = (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_gat_forward"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                 C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                                 C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_gat_backward"] = (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                  C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_void_p])
MODEL_SIGNATURES["stg_rnn_batch_tile"] = (C.c_int, [C.c_int])
MODEL_SIGNATURES["stg_rnn_saved_floats"] = (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int])
MODEL_SIGNATURES["stg_rnn_forward"] = (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                                 C.c_void_p, C.c_void_p])
MODEL_SIGNATURES["stg_rnn_backward"] = (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                                  C.c_void_p, C.c_void_p])
SIGNATURES.update(MODEL_SIGNATURES)

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Returns the loaded CDLL; builds it first if the sources changed.  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("STG_NO_BUILD") != "1":
        _build.build()
    if not os.path.exists(_build.LIB):
        raise RuntimeError(f"{_build.LIB} is missing and could not be built: the CUDA extension is required "
                           "(there is no CPU fallback)")
    lib = C.CDLL(_build.LIB)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class StgError(RuntimeError):
    pass


def check(status: int, what: str = "") -> None:
    if status == 0:
        return
    msg = load().stg_last_error().decode()
    if status in (-1, -2):
        raise ValueError(f"{what}: {msg} (status {status})")
    raise StgError(f"{what}: {msg} (status {status})")


class kernel_profile:
    """Context manager around stg_profile_*: per-kernel CUDA-event durations of the launches made
    inside the `with` block.  .result() -> {kernel_name: (total_ms, launches)} (synchronises)."""

    def __enter__(self):
        lib = load()
        lib.stg_profile_reset()
        lib.stg_profile_enable(1)
        return self

    def __exit__(self, *exc):
        load().stg_profile_enable(0)
        return False

    def result(self):
        lib = load()
        out = {}
        for slot in range(lib.stg_profile_slots()):
            ms, n = C.c_double(0.0), C.c_int64(0)
            check(lib.stg_profile_read(slot, C.byref(ms), C.byref(n)), "stg_profile_read")
            if n.value:
                out[lib.stg_profile_name(slot).decode()] = (ms.value, n.value)
        return out
