"""ctypes binding of libtextreid_b200.so (the C ABI declared in include/textreid_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRB_LIB") or os.path.join(_HERE, "libtextreid_b200.so")   # TRB_LIB: A/B tuning builds

TOPK_DEPTH = 10

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_int = C.c_int
_f = C.c_float


class MocoShape(C.Structure):
    _fields_ = [("N", _i32), ("D", _i32), ("K", _i32), ("C", _i32)]


class MocoHParams(C.Structure):
    _fields_ = [("T", _f), ("epsilon", _f), ("alpha", _f), ("beta", _f), ("scale_pos", _f), ("scale_neg", _f)]


class EmaChunk(C.Structure):
    _fields_ = [("k", _p), ("q", _p), ("n", _i64)]


# name -> (restype, argtypes); must list every symbol declared in include/textreid_b200.h
SIGNATURES = {
    "trb_version": (_int, []),
    "trb_last_error_string": (C.c_char_p, []),
    "trb_l2_normalize_rows_f32": (_int, [_p, _p, _p, _i64, _i64, _f, _p]),
    "trb_retrieval_thresholds_f32": (_int, [_p, _p, _p, _p, _p, _i64, _i64, _p]),
    "trb_retrieval_stream_f32": (_int, [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _int, _p, _p, _p, _p]),
    "trb_rank_similarity_f32": (_int, [_p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "trb_rank_rerank_f64": (_int, [_p, _i64, _i64, _i64, _i64, _p, _p, _int, C.c_double, _p, _p, _p, _p, _p, _p]),
    "trb_rank_scores_f64": (_int, [_p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "trb_jaccard_f64": (_int, [_p, _p, _int, C.c_double, _p, _i64, _i64, _p]),
    "trb_similarity_f32": (_int, [_p, _p, _p, _i64, _i64, _i64, _p]),
    "trb_retrieval_finish": (_int, [_p, _p, _int, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "trb_retrieval_metrics": (_int, [_p, _p, _i64, C.POINTER(_i32), _int, _p, _p, _p]),
    "trb_retrieval_tc_lists_per_split": (_int, []),
    "trb_packed_rows": (_i64, [_i64]),
    "trb_packed_bytes": (_i64, [_i64, _i64]),
    "trb_pack_rows_bf16": (_int, [_p, _int, _p, _int, _f, _p, _i64, _i64, _p]),
    "trb_retrieval_stream_tc": (_int, [_p, _p, _i64, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _int, _int, _int,
                                       _p, _p, _p, _p]),
    "trb_moco_loss_workspace_bytes": (_i64, [C.POINTER(MocoShape), _int]),
    "trb_moco_loss": (_int, [_p, _p, _p, _p, _p, _p, _int, _p, _p, _p, _p, _p, _p, _p, C.POINTER(MocoShape),
                             C.POINTER(MocoHParams), _int, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "trb_moco_loss_launches": (_int, [C.POINTER(MocoShape), _int]),
    "trb_moco_step": (_int, [_p, _p, _p, _p, _p, _p, _int, _p, _p, _p, _p, _p, _p, _p, _p, C.POINTER(MocoShape),
                             C.POINTER(MocoHParams), _int, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "trb_moco_step_launches": (_int, [C.POINTER(MocoShape), _int]),
    "trb_moco_loss_debug_stamps": (_int, [_p, C.POINTER(MocoShape), _p]),
    "trb_moco_loss_debug_logits": (_int, [_p, C.POINTER(MocoShape), _p]),
    "trb_moco_grad_combine": (_int, [_p, _p, _p, _p, _p, _p, _p, _int, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "trb_combine3_f32": (_int, [_p, _p, _p, _p, _p, _i64, _p]),
    "trb_scale_inplace_f32": (_int, [_p, _p, _i64, _p]),
    "trb_ema_update_f32": (_int, [_p, _p, _i64, _f, _f, _p]),
    "trb_ema_update_chunks_f32": (_int, [_p, _i64, _i64, _f, _f, _p]),
    "trb_enqueue": (_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libtextreid_b200.so is not built (%s). Run `python -m textreid_b200.build` "
            "(or __graft_entry__.build()). There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


# kernels launched per successful library call (bench.py reports the sum as "gpu_launches")
KERNELS_PER_CALL = {
    "trb_l2_normalize_rows_f32": 1, "trb_retrieval_thresholds_f32": 1, "trb_retrieval_stream_f32": 1,
    "trb_rank_similarity_f32": 1, "trb_rank_rerank_f64": 1, "trb_rank_scores_f64": 1, "trb_jaccard_f64": 1, "trb_similarity_f32": 1, "trb_retrieval_finish": 1, "trb_retrieval_metrics": 1,
    "trb_pack_rows_bf16": 1, "trb_retrieval_stream_tc": 1, "trb_moco_loss": 0, "trb_moco_step": 0, "trb_moco_grad_combine": 1, "trb_combine3_f32": 1,
    "trb_scale_inplace_f32": 1, "trb_ema_update_f32": 1, "trb_ema_update_chunks_f32": 1, "trb_enqueue": 2,
}
_launches = 0


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def launch_count() -> int:
    return _launches


def add_launches(n: int) -> None:
    global _launches
    _launches += int(n)


def check(rc: int, what: str) -> None:
    global _launches
    _launches += KERNELS_PER_CALL.get(what.split("(")[0], 0)
    if rc != 0:
        msg = load().trb_last_error_string()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("textreid_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path"
                               % t.device.type)
