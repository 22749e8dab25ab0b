"""ctypes binding of libvpr_b200.so (the C-ABI declared in include/vpr_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VPR_B200_LIB") or os.path.join(PKG_DIR, "libvpr_b200.so")    # override: an alternate build (A/B runs)

VPR_F32, VPR_BF16 = 0, 1
BWD_ALL, BWD_PARAMS, BWD_INPUTS = 0, 1, 2
OUT_F32, OUT_BF16, OUT_ATOMIC_F32 = 0, 1, 2
LATLON_MSE, LATLON_HUBER, LATLON_EXTERNAL = 0, 1, 2
ANGLE_NONE, ANGLE_ATAN2_MAAE, ANGLE_ACOS, ANGLE_MSE, ANGLE_SCALAR_MAAE = 0, 1, 2, 3, 4
HEADS_FORWARD_ONLY, HEADS_PREPACKED, HEADS_SKIP_WGRAD, HEADS_WGRAD_ONLY, HEADS_PREZEROED = 1, 2, 4, 8, 16
HEAD_DROP_SEED_OFFSET = 0x9E3779B97F4A7C15      # the angle head's input dropout draws from seed + this
ABI_VERSION = 5

_fp = C.c_void_p


class SaladShape(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("B", "n", "d", "hidden", "m", "l", "g", "iters")]


class SaladParams(C.Structure):
    _fields_ = [(k, _fp) for k in (
        "cluster_w0", "cluster_b0", "cluster_w1", "cluster_b1", "score_w0", "score_b0", "score_w1",
        "score_b1", "token_w0", "token_b0", "token_w1", "token_b1", "dust_bin")]


class SaladGrads(C.Structure):
    _fields_ = SaladParams._fields_


class HeadsShape(C.Structure):
    _fields_ = [("B", C.c_int), ("D", C.c_int), ("hidden", C.c_int), ("latlon_loss", C.c_int),
                ("angle_loss", C.c_int), ("huber_delta", C.c_float), ("loss_scale", C.c_float),
                ("global_B", C.c_int), ("reg_dropout_p", C.c_float), ("seed", C.c_uint64),
                ("head_dropout_p", C.c_float), ("flags", C.c_int), ("seed_offset_dev", C.c_void_p)]


COMM_MAX_WORLD, COMM_MAX_CTAS = 8, 160


class CommPeers(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("data", C.c_void_p * COMM_MAX_WORLD),
                ("staging", C.c_void_p * COMM_MAX_WORLD), ("flags", C.c_void_p * COMM_MAX_WORLD),
                ("mc_data", C.c_void_p), ("staging_elems", C.c_size_t)]


class HeadsParams(C.Structure):
    _fields_ = [(k, _fp) for k in ("reg_w0", "reg_b0", "reg_w1", "reg_b1", "ang_w", "ang_b")]


class HeadsGrads(C.Structure):
    _fields_ = HeadsParams._fields_


EXPORTS = {
    "vpr_last_error": (C.c_char_p, []),
    "vpr_abi_version": (C.c_int, []),
    "vpr_launch_count": (C.c_ulonglong, []),
    "vpr_salad_save_bytes": (C.c_size_t, [C.POINTER(SaladShape), C.c_int]),
    "vpr_salad_workspace_bytes": (C.c_size_t, [C.POINTER(SaladShape), C.c_int]),
    "vpr_salad_fwd": (C.c_int, [C.POINTER(SaladParams), _fp, _fp, _fp, _fp, _fp, C.c_size_t,
                                C.POINTER(SaladShape), C.c_int, C.c_float, C.c_uint64, _fp, _fp]),
    "vpr_counter_advance": (C.c_int, [_fp, _fp]),
    "vpr_salad_bwd": (C.c_int, [C.POINTER(SaladParams), _fp, _fp, _fp, _fp, _fp, C.c_size_t,
                                C.POINTER(SaladGrads), _fp, _fp, C.POINTER(SaladShape), C.c_int, C.c_float,
                                C.c_int, C.c_int, _fp]),
    "vpr_dropout_mask": (C.c_int, [C.c_uint64, C.c_float, C.c_int64, C.c_int64, _fp, _fp]),
    "vpr_sinkhorn_fwd": (C.c_int, [_fp, _fp, _fp, C.c_int, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "vpr_sinkhorn_bwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, _fp, C.c_int, C.c_int, C.c_int,
                                   C.c_int, _fp]),
    "vpr_heads_workspace_bytes": (C.c_size_t, [C.POINTER(HeadsShape), C.c_int]),
    "vpr_heads_pack": (C.c_int, [C.POINTER(HeadsParams), _fp, C.c_size_t, C.POINTER(HeadsShape), C.c_int, _fp]),
    "vpr_heads_zero": (C.c_int, [_fp, C.POINTER(HeadsGrads), _fp, C.c_size_t, C.POINTER(HeadsShape), C.c_int, _fp]),
    "vpr_heads_loss_fwd_bwd": (C.c_int, [C.POINTER(HeadsParams), _fp, _fp, _fp, _fp, _fp, _fp,
                                         C.POINTER(HeadsGrads), _fp, _fp, C.c_size_t, C.POINTER(HeadsShape),
                                         C.c_int, _fp]),
    "vpr_comm_flag_bytes": (C.c_size_t, []),
    "vpr_comm_staging_elems": (C.c_size_t, [C.c_size_t, C.c_int]),
    "vpr_allreduce_sum_f32": (C.c_int, [C.POINTER(CommPeers), C.c_size_t, C.c_size_t, C.c_int, _fp]),
    "vpr_seam_ln_split_fwd": (C.c_int, [_fp, _fp, _fp, C.c_float, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, _fp]),
    "vpr_seam_ln_split_bwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, _fp]),
    "vpr_metrics_latlon": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int64, _fp, _fp]),
    "vpr_metrics_angle": (C.c_int, [_fp, _fp, C.c_int, C.c_int64, _fp, _fp, _fp]),
    "vpr_recall_at_k": (C.c_int, [_fp, _fp, _fp, C.c_int64, C.c_int, C.c_int64, _fp, C.c_int, _fp, _fp]),
    "vpr_preprocess_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "vpr_preprocess_batch_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vpr_preprocess_batch": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                       C.POINTER(C.c_float), _fp, C.c_int, _fp, _fp, C.c_size_t, _fp]),
    "vpr_preprocess_image": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                       _fp, C.c_int, _fp, _fp, C.c_size_t, _fp]),
    "vpr_knn_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "vpr_knn_topk": (C.c_int, [_fp, _fp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, _fp, _fp, _fp,
                               C.c_size_t, _fp]),
    "vpr_knn_merge": (C.c_int, [_fp, _fp, C.c_int, C.c_int64, C.c_int, _fp, _fp, _fp]),
    "vpr_gemm": (C.c_int, [C.c_int, _fp, C.c_int, C.c_int64, C.c_int64, _fp, C.c_int, C.c_int64, C.c_int64,
                           C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_int, C.c_int64, C.c_int64, _fp, C.c_int,
                           _fp, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_int, C.c_int, _fp, C.c_float,
                           C.c_uint64, _fp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library once.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU or PyTorch fallback for this path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.vpr_abi_version() != ABI_VERSION:
            raise RuntimeError("libvpr_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vpr_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device(t: torch.Tensor):
    """Context manager that makes ``t``'s device current for the duration of a library call: the kernels launch on
    the current device, and the library refuses to run on a device other than the one it first ran on."""
    return torch.cuda.device(t.device)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return VPR_BF16
    if dt == torch.float32:
        return VPR_F32
    raise RuntimeError(f"unsupported dtype {dt}: use torch.float32 or torch.bfloat16")
