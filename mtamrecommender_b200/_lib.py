"""ctypes binding of libmtam_b200.so (the C-ABI declared in include/mtam.h).

There is no fallback: if the shared library is missing or fails to load, importing a product
module that needs it raises.  Build with `python -m mtamrecommender_b200.build`.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmtam_b200.so")

ABI_VERSION = 2
KINDS = {"MTAM": 0, "PISTREC": 1, "SASREC": 2, "TA_SASREC": 3, "TISASREC": 4, "BPRMF": 5, "MTAM_VIA_T_GRU": 6,
         "MTAM_NO_TIME_AWARE_RNN": 7, "MTAM_VIA_RNN": 8}
GEMM_FP32, GEMM_TF32X3, GEMM_TF32 = 0, 1, 2
OPTIMIZERS = {"adam": 0, "sgd": 1}
S_LOSS, S_LOSS_ORIGIN, S_L2_NORM, S_GLOBAL_NORM, S_CLIP_SCALE, S_COUNT = 0, 1, 2, 3, 4, 8
PARAM_DEAD, PARAM_TABLE = 1, 2
NAME_MAX = 160


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("kind", C.c_int32), ("max_batch", C.c_int32), ("L", C.c_int32),
                ("D", C.c_int32), ("H", C.c_int32), ("N", C.c_int32), ("user_rows", C.c_int32),
                ("item_rows", C.c_int32), ("category_rows", C.c_int32), ("position_rows", C.c_int32),
                ("reg", C.c_float), ("clip", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("gemm_mode", C.c_int32), ("optimizer", C.c_int32), ("dropout", C.c_float),
                ("dropout_seed", C.c_uint32), ("reserved", C.c_int32 * 4)]


class Sizes(C.Structure):
    _fields_ = [("param_floats", C.c_uint64), ("workspace_bytes", C.c_uint64)]


class Batch(C.Structure):
    _fields_ = [("B", C.c_int32), ("user_id", C.c_void_p), ("item_list", C.c_void_p),
                ("category_list", C.c_void_p), ("position_list", C.c_void_p), ("time_list", C.c_void_p),
                ("timelast_list", C.c_void_p), ("timenow_list", C.c_void_p), ("target_item_id", C.c_void_p),
                ("target_item_category", C.c_void_p), ("target_item_time", C.c_void_p),
                ("seq_length", C.c_void_p)]


class RecordStore(C.Structure):
    _fields_ = [("n_records", C.c_int64), ("offsets", C.c_void_p), ("user_id", C.c_void_p),
                ("target_item_id", C.c_void_p), ("target_item_category", C.c_void_p), ("target_item_time", C.c_void_p),
                ("seq_length", C.c_void_p), ("item", C.c_void_p), ("category", C.c_void_p), ("position", C.c_void_p),
                ("time", C.c_void_p), ("timelast", C.c_void_p), ("timenow", C.c_void_p)]


class ParamInfo(C.Structure):
    _fields_ = [("name", C.c_char * NAME_MAX), ("rows", C.c_int32), ("cols", C.c_int32), ("ndim", C.c_int32),
                ("ld", C.c_int32), ("offset", C.c_uint64), ("flags", C.c_int32)]


class SparseView(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("D", C.c_int32), ("item_cat_rows", C.c_void_p),
                ("position_rows", C.c_void_p), ("user_rows", C.c_void_p), ("has_user", C.c_int32),
                ("dense_begin", C.c_uint64), ("user_offset", C.c_uint64), ("item_offset", C.c_uint64),
                ("category_offset", C.c_uint64), ("position_offset", C.c_uint64)]


PHASES = ("embed_fwd", "gru_x_gemm", "gru_fwd", "kv_gemm", "hop_fwd", "ce_fwd", "ce_bwd", "hop_bwd",
          "hop_param_grads", "gru_bwd", "gru_param_grads", "embed_bwd", "dense_norm", "scatter", "adam")

# every symbol include/mtam.h declares: (restype, argtypes)
_VP, _I32, _I64, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
SIGNATURES = {
    "mtam_pack_records": (C.c_int, [_VP, _VP, _I64, _I32, _I32, _VP]),
    "mtam_gather": (C.c_int, [_VP, _I32, _I32, _VP, _I64, _VP, _VP]),
    "mtam_scatter_add_workspace": (_SZ, [_I64, _I32, _I32]),
    "mtam_scatter_add": (C.c_int, [_VP, _I32, _I32, _VP, _VP, _I32, _I64, _I32, _VP, _SZ, _VP, _VP, _VP]),
    "mtam_sort_workspace": (_SZ, [_I64, _I32]),
    "mtam_sort_indices": (C.c_int, [_VP, _I64, _I32, _VP, _SZ, C.POINTER(_VP), C.POINTER(_VP), _VP]),
    "mtam_scatter_add_sorted_workspace": (_SZ, [_I64, _I32]),
    "mtam_scatter_add_sorted": (C.c_int, [_VP, _I32, _VP, _VP, _VP, _I32, _I64, _I32, _VP, _SZ, _VP]),
    "mtam_plan": (C.c_int, [C.POINTER(Config), C.POINTER(Sizes)]),
    "mtam_create": (C.c_int, [C.POINTER(Config), _VP, _VP, _VP, _VP, _VP, _SZ, C.POINTER(_VP)]),
    "mtam_destroy": (C.c_int, [_VP]),
    "mtam_last_error": (C.c_char_p, [_VP]),
    "mtam_param_count": (C.c_int, [_VP]),
    "mtam_param_info_get": (C.c_int, [_VP, _I32, C.POINTER(ParamInfo)]),
    "mtam_get_adam_step": (C.c_int, [_VP, C.POINTER(_I64)]),
    "mtam_set_adam_step": (C.c_int, [_VP, _I64]),
    "mtam_forward": (C.c_int, [_VP, C.POINTER(Batch), _VP, _VP, _VP, _VP]),
    "mtam_train_step": (C.c_int, [_VP, C.POINTER(Batch), C.c_double, _VP, _VP]),
    "mtam_forward_backward": (C.c_int, [_VP, C.POINTER(Batch), _I32, _VP, _VP, _VP]),
    "mtam_finish_grads": (C.c_int, [_VP, _VP, _I32, _VP]),
    "mtam_apply": (C.c_int, [_VP, C.c_double, _VP, _VP, _VP]),
    "mtam_apply_begin": (C.c_int, [_VP, C.c_double, _VP, _VP, _VP]),
    "mtam_apply_range": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VP]),
    "mtam_apply_end": (C.c_int, [_VP, _VP]),
    "mtam_set_bpr_negative": (C.c_int, [_VP, _I32]),
    "mtam_forward_rows": (C.c_int, [_VP, C.POINTER(Batch), _VP, _VP, _VP, _VP]),
    "mtam_backward_rows": (C.c_int, [_VP, C.POINTER(Batch), _VP, _VP, _I32, _VP, _VP]),
    "mtam_sumsq_workspace": (_SZ, [_I64]),
    "mtam_sumsq": (C.c_int, [_VP, _I64, _VP, _VP, _SZ, _VP]),
    "mtam_set_dropout_state": (C.c_int, [_VP, C.c_uint32, C.c_uint32]),
    "mtam_set_item_grad_event": (C.c_int, [_VP, _VP]),
    "mtam_scatter_sparse_into": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "mtam_prepare_step": (C.c_int, [_VP, C.c_double, _VP]),
    "mtam_sparse_pieces": (C.c_int, [_VP, C.POINTER(SparseView)]),
    "mtam_profile_enable": (C.c_int, [_VP, _I32]),
    "mtam_profile_read": (C.c_int, [_VP, _VP, _I32]),
    "mtam_gemm": (C.c_int, [_I32] * 6 + [_VP, _I32, _VP, _I32, _VP, _I32, _VP, _I32, _I32, _VP, _SZ, _VP]),
    "mtam_gemm_workspace": (_SZ, [_I32, _I32, _I32]),
    "mtam_launch_count": (C.c_longlong, []),
    "mtam_eval_topk": (C.c_int, [_VP, C.POINTER(Batch), _I32, _VP, _VP, _VP]),
    "mtam_score_topk": (C.c_int, [_I32, _VP, _I32, _I32, _VP, _I32, _I32, _I32, _VP, _VP, _VP, _SZ, _VP]),
    "mtam_score_topk_workspace": (_SZ, [_I32, _I32, _I32]),
    "mtam_score_bucket_max": (C.c_int, [_VP, _I32, _I32, _VP, _I32, _I32, _VP, _I32, _VP]),
    "mtam_set_topk_bucket_crossover": (C.c_int, [_I32]),
    "mtam_merge_topk": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _VP, _VP, _VP]),
    "mtam_softmax_ce_workspace": (_SZ, [_I32, _I32, _I32]),
    "mtam_softmax_ce_forward": (C.c_int, [_I32, _VP, _I32, _I32, _VP, _I32, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mtam_softmax_ce_backward": (C.c_int, [_I32, _VP, _I32, _I32, _VP, _I32, _VP, _VP, C.c_float, _VP, _VP, _VP, _SZ, _VP]),
    "mtam_hr_ndcg": (C.c_int, [_VP, _I32, _I32, _VP, _VP, _VP]),
}

_lib = None


class MtamError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads the library once.  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MtamError(f"{LIB_PATH} not found: build it with `python -m mtamrecommender_b200.build` "
                        "(there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().mtam_last_error(None)
        raise MtamError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")
