"""ctypes binding of libpeneo_b200.so (the C ABI declared in include/peneo_b200.h).

There is no CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpeneo_b200.so")

NUM_HEADS = 5
HEAD_CLASSES = (2, 3, 3, 3, 3)
PREC_FP32, PREC_BF16, PREC_TF32 = 0, 1, 2
DT_F32, DT_BF16, DT_F16, DT_I64 = 0, 1, 2, 3
E_OVERFLOW = -4

# every symbol include/peneo_b200.h declares (tests check that the library exports them all)
SYMBOLS = (
    "peneo_abi_version",
    "peneo_last_error",
    "peneo_device_supported",
    "peneo_pack_bytes",
    "peneo_pack_weights",
    "peneo_token_proj_workspace_bytes",
    "peneo_token_proj_fwd",
    "peneo_gather_tokens",
    "peneo_token_dropout_bwd",
    "peneo_pair_heads_fwd",
    "peneo_heads_bwd_workspace_bytes",
    "peneo_heads_bwd",
    "peneo_pair_loss_workspace_bytes",
    "peneo_pair_loss_fwd",
    "peneo_pair_loss_bwd",
    "peneo_fused_loss_supported",
    "peneo_pair_heads_loss_fwd",
    "peneo_heads_loss_bwd",
    "peneo_pair_loss_ohem_workspace_bytes",
    "peneo_pair_loss_ohem_fwd",
    "peneo_pair_loss_ohem_bwd",
    "peneo_scatter_tags",
    "peneo_decode_spots_workspace_bytes",
    "peneo_decode_spots",
    "peneo_pair_heads_spots_workspace_bytes",
    "peneo_pair_heads_spots_fwd",
    "peneo_decode_resolve_doc_ints",
    "peneo_decode_resolve_workspace_bytes",
    "peneo_decode_resolve",
    "peneo_selftest",
    "peneo_probe_rates",
)


class Dims(C.Structure):
    _fields_ = [("hin", C.c_int32), ("hid", C.c_int32), ("d", C.c_int32), ("shrink", C.c_int32),
                ("num_layers", C.c_int32)]


class Params(C.Structure):
    _fields_ = [
        ("shrink_w1", C.c_void_p), ("shrink_b1", C.c_void_p), ("shrink_w2", C.c_void_p), ("shrink_b2", C.c_void_p),
        ("combine_w", C.c_void_p), ("combine_b", C.c_void_p),
        ("mid_w", C.c_void_p * (NUM_HEADS * 8)), ("mid_b", C.c_void_p * (NUM_HEADS * 8)),
        ("out_w", C.c_void_p * NUM_HEADS), ("out_b", C.c_void_p * NUM_HEADS),
    ]


class Dropout(C.Structure):
    """peneo_dropout: drop probability + per-step seed (NULL pointer = eval mode)."""

    _fields_ = [("p", C.c_float), ("seed", C.c_uint64)]


class SavedAct(C.Structure):
    """peneo_saved_act: activations the fused forward stores for a backward pass without recompute (both or NULL)."""

    _fields_ = [("h", C.c_void_p), ("s", C.c_void_p)]


class Grads(C.Structure):
    _fields_ = Params._fields_


PtrArray5 = C.c_void_p * NUM_HEADS

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m peneo_b200.build` "
            "(peneo_b200 has no CPU or PyTorch fallback path)"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    lib.peneo_abi_version.restype = C.c_int
    lib.peneo_last_error.restype = C.c_char_p
    lib.peneo_device_supported.argtypes = [C.c_int]
    lib.peneo_pack_bytes.restype = sz
    lib.peneo_pack_bytes.argtypes = [C.POINTER(Dims), C.c_int]
    lib.peneo_pack_weights.argtypes = [C.POINTER(Dims), C.c_int, C.POINTER(Params), vp, vp]
    lib.peneo_token_proj_workspace_bytes.restype = sz
    lib.peneo_token_proj_workspace_bytes.argtypes = [C.POINTER(Dims), C.c_int, i64]
    lib.peneo_token_proj_fwd.argtypes = [C.POINTER(Dims), C.c_int, vp, vp, C.c_int, i64, i64, vp, vp, C.POINTER(Dropout), vp]
    lib.peneo_gather_tokens.argtypes = [vp, C.c_int, i32, i32, i32, i64, i64, vp, C.c_int, C.POINTER(Dropout), vp]
    lib.peneo_token_dropout_bwd.argtypes = [vp, i64, i32, C.POINTER(Dropout), vp]
    lib.peneo_pair_heads_fwd.argtypes = [C.POINTER(Dims), C.c_int, vp, vp, i32, i32, PtrArray5, C.POINTER(Dropout), vp]
    lib.peneo_heads_bwd_workspace_bytes.restype = sz
    lib.peneo_heads_bwd_workspace_bytes.argtypes = [C.POINTER(Dims), C.c_int, i32, i32]
    lib.peneo_heads_bwd.argtypes = [C.POINTER(Dims), C.c_int, vp, vp, C.c_int, i64, i32, i32, PtrArray5,
                                    C.POINTER(Grads), vp, vp, C.POINTER(Dropout), vp]
    lib.peneo_pair_loss_workspace_bytes.restype = sz
    lib.peneo_pair_loss_workspace_bytes.argtypes = [i32, i32]
    lib.peneo_pair_loss_fwd.argtypes = [i32, i32, PtrArray5, PtrArray5, C.POINTER(C.c_float), C.POINTER(C.c_float), vp,
                                        vp, vp]
    lib.peneo_pair_loss_bwd.argtypes = [i32, i32, PtrArray5, PtrArray5, C.POINTER(C.c_float), C.POINTER(C.c_float), vp,
                                        vp, PtrArray5, vp]
    lib.peneo_fused_loss_supported.argtypes = [C.POINTER(Dims), C.c_int]
    lib.peneo_pair_heads_loss_fwd.argtypes = [C.POINTER(Dims), C.c_int, vp, vp, i32, i32, PtrArray5, PtrArray5,
                                              C.POINTER(C.c_float), C.POINTER(C.c_float), vp, vp, C.POINTER(Dropout), vp,
                                              C.POINTER(SavedAct)]
    lib.peneo_heads_loss_bwd.argtypes = [C.POINTER(Dims), C.c_int, vp, vp, C.c_int, i64, i32, i32, PtrArray5, PtrArray5,
                                         C.POINTER(C.c_float), C.POINTER(C.c_float), vp, vp, C.POINTER(Grads), vp, vp,
                                         C.POINTER(Dropout), vp, C.POINTER(SavedAct)]
    lib.peneo_pair_loss_ohem_workspace_bytes.restype = sz
    lib.peneo_pair_loss_ohem_workspace_bytes.argtypes = [i32, i32]
    lib.peneo_pair_loss_ohem_fwd.argtypes = [i32, i32, PtrArray5, PtrArray5, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                             i32, i32, vp, vp, vp]
    lib.peneo_pair_loss_ohem_bwd.argtypes = [i32, i32, PtrArray5, PtrArray5, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                             vp, vp, PtrArray5, vp]
    lib.peneo_scatter_tags.argtypes = [vp, i64, i32, i32, vp, vp]
    lib.peneo_decode_spots_workspace_bytes.restype = sz
    lib.peneo_decode_spots_workspace_bytes.argtypes = [i32, i32]
    lib.peneo_decode_spots.argtypes = [i32, i32, PtrArray5, C.c_int, i32, vp, vp, vp, vp, vp, vp]
    lib.peneo_pair_heads_spots_workspace_bytes.restype = sz
    lib.peneo_pair_heads_spots_workspace_bytes.argtypes = [i32, i32]
    lib.peneo_pair_heads_spots_fwd.argtypes = [C.POINTER(Dims), C.c_int, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.peneo_decode_resolve_doc_ints.restype = sz
    lib.peneo_decode_resolve_doc_ints.argtypes = [i32, i32]
    lib.peneo_decode_resolve_workspace_bytes.restype = sz
    lib.peneo_decode_resolve_workspace_bytes.argtypes = [i32, i32]
    lib.peneo_decode_resolve.argtypes = [i32, i32, i32, vp, vp, vp, vp, C.c_int, C.c_float, vp, vp, vp]
    lib.peneo_selftest.argtypes = [C.POINTER(C.c_uint32), C.c_char_p, sz]
    lib.peneo_probe_rates.argtypes = [C.POINTER(C.c_double), C.c_int]
    if lib.peneo_abi_version() != 4:
        raise RuntimeError("libpeneo_b200.so ABI version mismatch; rebuild with `python -m peneo_b200.build --force`")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().peneo_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def dropout_arg(drop):
    """(p, seed) tuple or None -> ctypes argument for a `const peneo_dropout*` parameter."""
    if drop is None or drop[0] <= 0.0:
        return None
    return C.byref(Dropout(float(drop[0]), int(drop[1]) & 0xFFFFFFFFFFFFFFFF))


def ptrs5(tensors: Sequence) -> PtrArray5:
    return PtrArray5(*[t.data_ptr() for t in tensors])


def floats(vals: Sequence[float]):
    return (C.c_float * len(vals))(*[float(v) for v in vals])
