"""ctypes loader for librmpnn_b200.so (the C ABI in include/rmpnn_b200.h)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_STAGES = 6
MAX_SEGS = 3
EPI_STORE, EPI_NORMALIZE, EPI_RESID_LN, EPI_SEGSUM, EPI_RESID_ADD = 0, 1, 2, 3, 4
PREC_FP32, PREC_BF16, PREC_FP32X3 = 0, 1, 2
OPERAND_BF16, OPERAND_FP16 = 0, 1


class LibraryMissing(RuntimeError):
    pass


class MlpDesc(C.Structure):
    """Mirror of ``struct rmp_mlp_desc``."""
    _fields_ = [
        ("seg_src", C.c_void_p * MAX_SEGS),
        ("seg_idx", C.c_void_p * MAX_SEGS),
        ("seg_width", C.c_int32 * MAX_SEGS),
        ("seg_bf16", C.c_int32 * MAX_SEGS),
        ("num_segs", C.c_int32),
        ("Wt", C.c_void_p * MAX_STAGES),
        ("bias", C.c_void_p * MAX_STAGES),
        ("k_in", C.c_int32 * MAX_STAGES),
        ("n_out", C.c_int32 * MAX_STAGES),
        ("gelu", C.c_int32 * MAX_STAGES),
        ("num_stages", C.c_int32),
        ("epilogue", C.c_int32),
        ("gain", C.c_void_p),
        ("beta", C.c_void_p),
        ("resid", C.c_void_p),
        ("seg_id", C.c_void_p),
        ("alpha", C.c_float),
        ("row_scale", C.c_void_p),
        ("out", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("rows_dev", C.c_void_p),
        ("rows_slot", C.c_int32),
        ("rows_cap", C.c_int32),
    ]


P, I, F, U64, I64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_int64

# name -> argtypes; every entry point returns int except the string getters.
SIGNATURES = {
    "rmp_device_arch": [],
    "rmp_abi_struct_size": [],
    "rmp_graph_index": [P, I, I, I, P, P, P, P, P, P],
    "rmp_knn": [P, P, I, I, I, I, I, P, P, P],
    "rmp_edge_list": [P, P, P, P, P, P, I, I, I, I, P, P, P, P],
    "rmp_rdesign_raw_features": [P, I, I, P, P, P, P, I, I, P, P, P, P, P],
    "rmp_rdesign_raw_features_k128": [P, I, I, P, P, P, P, I, I, P, P, P, P, P],
    "rmp_row_mlp": [C.POINTER(MlpDesc), P],
    "rmp_row_mlp_bf16": [C.POINTER(MlpDesc), P, P],
    "rmp_row_mlp_tc": [C.POINTER(MlpDesc), P, I, P],
    "rmp_row_mlp_x3": [C.POINTER(MlpDesc), P, P, P, P, P, I, P],
    "rmp_add_layernorm": [P, P, P, P, F, P, I, I, I, I, P, P],
    "rmp_readout_head": [P, I, P, P, F, U64, P, I, I, P, P, P, P],
    "rmp_cast_bf16": [P, P, I64, P],
    "rmp_collate": [P, I, P, P, I, I, I, I, P, P, P],
    "rmp_forest_predict": [P, I, I, I, P, I, P, I, P, P, I, F, P, P, P],
    "rmp_res_graph": [P, P, I, I, I, I, P, P, P],
    "rmp_res_edge_rows": [P, P, I, I, I, P, P, P, P, P],
    "rmp_res_raw_features": [P, P, P, I, I, I, I, I, I, I, I, I, I, P, P, P],
    "rmp_graph_norm": [P, P, P, I, I, I, P, P, I, P, P, P, P],
    "rmp_res_attention": [P, P, I, I, I, I, P, P],
    "rmp_res_attention_tc": [P, P, I, I, I, I, P, P],
    "rmp_scale_rows": [P, P, I64, I, P],
    "rmp_readout_head_masked": [P, I, P, P, P, F, U64, I, P, P, P, P],
    "rmp_edge_embed_bf16": [P, P, P, P, P, I, P, P, P, P, P, P, P],
    "rmp_node_mlp_bf16": [P, P, I, P, P, I, P, P, P, P, P, P, P, P, P, P, P, I, P],
    "rmp_edge_mlp_bf16": [P, P, I, P, P, P, P, P, P, I, F, P, P, P],
    "rmp_edge_mlp_bf16_traced": [P, P, I, P, P, P, P, P, P, I, F, P, P, P, P],
    "rmp_edge_mlp_train_fwd_bf16": [P, P, I, P, P, P, P, P, P, I, F, P, P, P, F, U64, P, C.c_uint32, P],
    "rmp_act_fwd_bf16": [P, P, I64, I, F, U64, P, C.c_uint32, P],
    "rmp_act_bwd_bf16": [P, P, P, P, F, P, P, I, F, U64, P, C.c_uint32, P],
    "rmp_scatter_add_rows": [P, P, P, I, P],
    "rmp_segment_sum_rows": [P, P, P, I, P],
    "rmp_segment_sum_rows_perm": [P, P, P, P, I, P],
    "rmp_gemm_dw_bf16": [P, P, P, I, P],
    "rmp_adam_step": [P, P, P, P, I64, F, F, F, F, F, I, P, P],
    "rmp_normalize_fwd_bf16": [P, P, P, P, I, F, P],
    "rmp_normalize_bwd_bf16": [P, P, P, P, P, P, P, I, F, P],
    "rmp_gemm_rows_bf16": [P, P, P, P, P, P, P, P, P, P, I, I, I, F, U64, P, C.c_uint32, P],
    "rmp_gemm_rows_kcat_bf16": [P, P, I, P, I, I, P],
    "rmp_colsum": [P, I, P, P, P, I, I, P, P, P],
    "rmp_add_layernorm_fwd": [P, P, I, P, P, F, I, P, P, P, P, P],
    "rmp_layernorm_bwd": [P, P, P, P, P, I, P, P, P, P, P],
}
STRING_GETTERS = ("rmp_version", "rmp_last_error_string")


def lib_path() -> str:
    return os.environ.get("RMPNN_B200_LIB", os.path.join(_HERE, "librmpnn_b200.so"))


def lib():
    """The loaded library.  Raises LibraryMissing (never falls back) if it is not built."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise LibraryMissing(
                f"{path} not found: build it with `python __graft_entry__.py build` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no fallback path")
        L = C.CDLL(path)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.rmp_launch_count.restype = C.c_uint64
        L.rmp_launch_count.argtypes = []
        for name in STRING_GETTERS:
            getattr(L, name).restype = C.c_char_p
            getattr(L, name).argtypes = []
        _LIB = L
    return _LIB


def check(rc: int, what: str = ""):
    if rc == 0:
        return
    msg = lib().rmp_last_error_string().decode()
    if rc == 2:                                   # cudaErrorMemoryAllocation
        import torch
        raise torch.OutOfMemoryError(f"{what}: {msg}")
    raise RuntimeError(f"{what or 'rmpnn_b200'} failed (code {rc}): {msg}")
