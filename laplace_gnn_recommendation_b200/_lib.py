"""ctypes binding of liblaplace_b200.so (the C ABI declared in include/laplace_b200.h).

PyTorch is used only for device memory and streams: every call passes raw ``data_ptr()``s and the
current CUDA stream handle.  There is no CPU fallback: if the shared library is missing or a call
fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "liblaplace_b200.so")

c_i32, c_i64, c_f32, c_vp, c_sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class LgbCsr(C.Structure):
    """struct lgb_csr (include/laplace_b200.h)."""
    _fields_ = [
        ("n_rows", c_i64), ("n_cols", c_i64), ("nnz", c_i64),
        ("rowptr", c_vp), ("colidx", c_vp), ("val", c_vp), ("row_order", c_vp),
        ("chunk", c_i32), ("_pad", c_i32),
        ("n_long", c_i64), ("n_tasks", c_i64),
        ("long_rows", c_vp), ("long_ptr", c_vp), ("task_row", c_vp), ("task_start", c_vp), ("task_end", c_vp),
        ("colidx_hot", c_vp), ("hot_cols", c_vp), ("n_hot", c_i32), ("_pad2", c_i32),
        ("seg_row", c_vp), ("seg_t0", c_vp), ("seg_t1", c_vp), ("row_seg0", c_vp), ("n_seg", c_i64),
        ("task_exec", c_vp), ("task_seg", c_vp),
    ]


class LgbBprArgs(C.Structure):
    """struct lgb_bpr_args (include/laplace_b200.h)."""
    _fields_ = [
        ("uf", c_vp), ("u0", c_vp), ("pf", c_vp), ("p0", c_vp), ("nf", c_vp), ("n0", c_vp),
        ("iu", c_vp), ("ip", c_vp), ("in_", c_vp),
        ("B", c_i64), ("B_norm", c_i64), ("d", c_i32), ("lambda_", c_f32), ("gscale", c_f32), ("flags", c_i32),
        ("user_lo", c_i64), ("user_hi", c_i64),
        ("gout", c_vp),
        ("duf", c_vp), ("du0", c_vp), ("dpf", c_vp), ("dp0", c_vp), ("dnf", c_vp), ("dn0", c_vp),
        ("loss", c_vp), ("ws", c_vp),
    ]


class LgbExchange(C.Structure):
    """struct lgb_exchange (include/laplace_b200.h)."""
    _fields_ = [("multicast_base", c_vp), ("peer_base", c_vp * 16), ("pad_base", c_vp * 16),
                ("rank", c_i32), ("world", c_i32), ("n_channels", c_i32), ("_pad", c_i32)]


# name -> (restype, argtypes); must list EVERY function include/laplace_b200.h declares
# (tests/test_abi.py parses the header and checks this table and the .so against it).
PROTOTYPES = {
    "lgb_abi_version": (C.c_int, []),
    "lgb_last_error": (C.c_char_p, []),
    "lgb_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "lgb_csr_build_ws_bytes": (C.c_int, [c_i64, c_i64, C.POINTER(c_sz)]),
    "lgb_csr_build": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lgb_csr_build_check": (C.c_int, [c_vp, c_vp]),
    "lgb_csr_transpose_ws_bytes": (C.c_int, [c_i64, c_i64, C.POINTER(c_sz)]),
    "lgb_csr_transpose": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lgb_gather_f32": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "lgb_gcn_norm": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "lgb_spmm_plan_count": (C.c_int, [c_vp, c_i64, c_i32, C.POINTER(c_i64), c_vp, c_sz, c_vp]),
    "lgb_spmm_plan_ws_bytes": (C.c_int, [c_i64, C.POINTER(c_sz)]),
    "lgb_spmm_plan_fill": (C.c_int, [c_vp, c_i64, c_i32, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lgb_degree_order_ws_bytes": (C.c_int, [c_i64, C.POINTER(c_sz)]),
    "lgb_degree_order": (C.c_int, [c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "lgb_spmm": (C.c_int, [C.POINTER(LgbCsr), c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_f32, c_i32, c_vp, c_vp]),
    "lgb_spmm_split": (C.c_int, [C.POINTER(LgbCsr), c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_f32, c_i32, c_vp, c_i64, c_vp, c_vp]),
    "lgb_spmm_rowsparse": (C.c_int, [C.POINTER(LgbCsr), c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_i32, c_vp, c_vp]),
    "lgb_rows_bitmap": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "lgb_scale_rows_nonzero": (C.c_int, [c_vp, c_i64, c_i32, c_f32, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "lgb_segment_max": (C.c_int, [C.POINTER(LgbCsr), c_vp, c_i32, c_vp, c_vp, c_vp]),
    "lgb_segment_max_bwd": (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "lgb_row_div_by_degree": (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "lgb_gcn_values": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "lgb_accumulate": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_f32, c_vp, c_vp]),
    "lgb_mean_rows": (C.c_int, [C.POINTER(c_vp), c_i32, c_i64, c_f32, c_vp, c_vp]),
    "lgb_zero": (C.c_int, [c_vp, c_sz, c_vp]),
    "lgb_scale_concat": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_i32, c_f32, c_vp, c_vp]),
    "lgb_bpr_blocks": (c_i64, [c_i64]),
    "lgb_bpr": (C.c_int, [C.POINTER(LgbBprArgs), c_vp]),
    "lgb_gather_rows_owned": (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_i64, c_i64, c_vp, c_vp]),
    "lgb_scatter_add_rows_owned": (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_i64, c_i64, c_vp, c_vp]),
    "lgb_edge_concat_fwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp]),
    "lgb_edge_concat_bwd": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "lgb_edge_dot_fwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "lgb_edge_dot_bwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "lgb_edge_mlp2_fwd": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "lgb_linear_wgrad_ws_bytes": (C.c_int, [c_i64, c_i32, c_i32, C.POINTER(c_sz)]),
    "lgb_linear_wgrad": (C.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "lgb_topk_exclude": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "lgb_topk_exclude_tiled": (C.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "lgb_neg_reject_mask": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "lgb_sort_keys_ws_bytes": (C.c_int, [c_i64, C.POINTER(c_sz)]),
    "lgb_edge_keys_sorted": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "lgb_segment_expand": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "lgb_bucketize_segmented": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "lgb_adam_step": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_i32, c_vp]),
    "lgb_multimem_allreduce_f32": (C.c_int, [c_vp, c_i64, c_i32, c_i32, c_vp]),
    "lgb_peer_allreduce_f32": (C.c_int, [C.POINTER(C.c_uint64), c_i64, c_i32, c_i32, c_vp]),
    "lgb_exchange_pad_words": (C.c_int, [c_i32]),
    "lgb_exchange_allreduce_f32": (C.c_int, [C.POINTER(LgbExchange), c_i64, c_i64, c_i32, c_i32, c_vp]),
}

_lib: Optional[C.CDLL] = None
LAUNCHES = 0  # number of library compute calls issued by this process (bench.py reports kernels launched)


def load() -> C.CDLL:
    """Load the in-tree shared library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m laplace_gnn_recommendation_b200.build` "
            "(there is no CPU or PyTorch fallback for the propagation kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.lgb_abi_version() != 4:
        raise RuntimeError(f"liblaplace_b200.so ABI {lib.lgb_abi_version()} != 4")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().lgb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"liblaplace_b200 {what} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Raw device pointer (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr() if t.numel() > 0 else None


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def on_device(t: torch.Tensor) -> bool:
    """True when ``t`` lives where the kernels run (a gate for OPTIONAL fused paths; mandatory paths use require_cuda)."""
    return bool(t.is_cuda)


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "laplace_gnn_recommendation_b200 runs on CUDA tensors only (sm_100a kernels, no CPU fallback); "
                f"got a tensor on {t.device}")


def f32c(t: torch.Tensor) -> torch.Tensor:
    """Contiguous fp32 view/copy of a tensor (the ABI wants row-major fp32 with stride d)."""
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def i64c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t if t.is_contiguous() else t.contiguous()


def count_launch(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n
