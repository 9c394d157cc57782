"""ctypes binding of librecsys_b200.so (the C ABI declared in include/recsys_b200.h).

There is no fallback: if the shared library is missing or the device is not a B200 the import of
the engine fails loudly (RuntimeError), it never routes through PyTorch eager or the CPU.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librecsys_b200.so")

REC_MAX_HEADS = 5
REC_MAX_NETS = 2
REC_MAX_TOPK = 32
REC_MAX_KLIST = 8
REC_ABI_VERSION = 1

_fp = C.POINTER(C.c_float)


class RecConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "item_num", "action_dim", "embedding_dim", "hidden_dim", "state_size", "bidirectional",
        "n_heads", "n_nets", "use_packed_seq", "frozen_pad_row", "max_batch", "vocab_lo", "vocab_hi",
        "max_topk")]


class RecNetParams(C.Structure):
    _fields_ = (
        [("emb", C.c_void_p), ("emb_m", C.c_void_p), ("emb_v", C.c_void_p)]
        + [(f"{n}{s}", C.c_void_p * 2) for n in ("w_ih", "w_hh", "b_ih", "b_hh") for s in ("", "_m", "_v")]
        + [(f"{n}{s}", C.c_void_p * REC_MAX_HEADS) for n in ("head_w", "head_b") for s in ("", "_m", "_v")]
    )


class RecBatch(C.Structure):
    _fields_ = [("B", C.c_int32), ("s", C.c_void_p), ("a", C.c_void_p), ("r", C.c_void_p),
                ("s_next", C.c_void_p), ("true_len", C.c_void_p), ("true_next_len", C.c_void_p),
                ("is_end", C.c_void_p)]


class RecTrainHparams(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("gamma", C.c_float), ("alpha", C.c_float), ("q_weights", C.c_float * 3),
                ("div_emb", C.c_void_p), ("div_dim", C.c_int32), ("topk_div", C.c_int32),
                ("topk_nov", C.c_int32), ("nov_reward", C.c_float), ("unpopular", C.c_void_p),
                ("out_to_in", C.c_void_p), ("pad_pos_end", C.c_int32), ("dropout_p", C.c_float),
                ("dropout_seed", C.c_uint64), ("dropout_mask", C.c_void_p)]


class RecEvalOpts(C.Structure):
    _fields_ = [("head_idx", C.c_int32), ("n_k", C.c_int32), ("ks", C.c_int32 * REC_MAX_KLIST),
                ("n_cov", C.c_int32), ("cov_ks", C.c_int32 * REC_MAX_KLIST), ("topk_div", C.c_int32),
                ("topk_nov", C.c_int32), ("nov_reward", C.c_float), ("div_emb", C.c_void_p),
                ("div_dim", C.c_int32), ("unpopular", C.c_void_p), ("out_to_in", C.c_void_p),
                ("pad_pos_end", C.c_int32)]


class RecEvalAccum(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("hits", "ndcg", "reps", "div_sum", "nov_sum", "loss_sum", "cov_bits")]


# every symbol include/recsys_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "rec_abi_version": (C.c_int, []),
    "rec_create": (C.c_int, [C.POINTER(RecConfig), _P, C.POINTER(_P)]),
    "rec_destroy": (None, [_P]),
    "rec_last_error": (C.c_char_p, [_P]),
    "rec_bind_params": (C.c_int, [_P, C.c_int, C.POINTER(RecNetParams)]),
    "rec_set_adam_step": (C.c_int, [_P, C.c_int, C.c_int64]),
    "rec_get_adam_step": (C.c_int64, [_P, C.c_int]),
    "rec_forward_state": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P]),
    "rec_head_logits": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int64]),
    "rec_train_step_supervised": (C.c_int, [_P, C.POINTER(RecBatch), C.POINTER(RecTrainHparams), _P]),
    "rec_train_step_q": (C.c_int, [_P, C.POINTER(RecBatch), C.POINTER(RecTrainHparams), C.c_int, _P]),
    "rec_train_step_sarm": (C.c_int, [_P, C.POINTER(RecBatch), C.POINTER(RecTrainHparams), _P]),
    "rec_train_step_supervised_host": (C.c_int, [_P, C.POINTER(RecBatch), C.POINTER(RecTrainHparams), _P]),
    "rec_train_step_q_host": (C.c_int, [_P, C.POINTER(RecBatch), C.POINTER(RecTrainHparams), C.c_int, _P]),
    "rec_record_floats": (C.c_int, [_P]),
    "rec_train_phase_a": (C.c_int, [_P, C.POINTER(RecBatch), C.POINTER(RecTrainHparams), C.c_int, _P]),
    "rec_train_phase_b": (C.c_int, [_P, _P, C.c_int, _P]),
    "rec_train_phase_c": (C.c_int, [_P, _P, _P, _P]),
    "rec_train_phase_d": (C.c_int, [_P, _P]),
    "rec_dp_packed_bytes": (C.c_int64, [_P, C.c_int]),
    "rec_dp_grad_floats": (C.c_int64, [_P]),
    "rec_dp_forward": (C.c_int, [_P, C.POINTER(RecBatch), C.c_int, _P]),
    "rec_dp_unpack": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(RecBatch)]),
    "rec_train_phase_a_heads": (C.c_int, [_P, C.POINTER(RecBatch), C.POINTER(RecTrainHparams), C.c_int, _P]),
    "rec_dp_backward": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "rec_dp_apply": (C.c_int, [_P, _P, _P]),
    "rec_set_embedding_shard": (C.c_int, [_P, C.c_int64, C.c_int64]),
    "rec_emb_rows_gather": (C.c_int, [_P, C.c_int, _P, C.c_int64, _P]),
    "rec_emb_rows_scatter": (C.c_int, [_P, C.c_int, _P, C.c_int64, _P]),
    "rec_packed_batch_bytes": (C.c_int64, [_P, C.c_int]),
    "rec_gather_batch": (C.c_int, [_P, C.POINTER(RecBatch), C.c_int64, _P, C.c_int, C.POINTER(RecBatch)]),
    "rec_build_replay_rows": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int64, C.c_int64, C.c_int, C.POINTER(RecBatch)]),
    "rec_pack_batch": (C.c_int, [_P, C.POINTER(RecBatch), _P]),
    "rec_unpack_batch": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(RecBatch)]),
    "rec_eval_batch": (C.c_int, [_P, C.c_int, C.POINTER(RecBatch), C.POINTER(RecEvalOpts),
                                 C.POINTER(RecEvalAccum), _P, _P]),
    "rec_eval_hold_params": (C.c_int, [_P, C.c_int]),
    "rec_eval_shard_candidates": (C.c_int, [_P, C.c_int, C.POINTER(RecBatch), C.c_int, C.c_int, _P]),
    "rec_eval_merge": (C.c_int, [_P, C.POINTER(RecBatch), C.POINTER(RecEvalOpts), _P, C.c_int,
                                 C.POINTER(RecEvalAccum), _P, _P]),
    "rec_set_cuda_graphs": (C.c_int, [_P, C.c_int]),
    "rec_set_stream": (C.c_int, [_P, _P]),
    "rec_set_tensor_cores": (C.c_int, [_P, C.c_int]),
    "rec_launch_count": (C.c_int64, [_P]),
    "rec_enable_kernel_timing": (C.c_int, [_P, C.c_int]),
    "rec_last_kernel_ms": (C.c_float, [_P, C.c_int]),
    "rec_debug_copy_astar": (C.c_int, [_P, _P, C.c_int]),
    "rec_debug_set_trace": (C.c_int, [_P, _P]),
    "rec_debug_tc_gemm": (C.c_int, [C.c_int, _P, _P, _P, _P]),
}

_lib = None


def load_library():
    """dlopen the in-tree shared object and type every entry point. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ikea-recommender-system_b200/csrc`). There is no CPU/eager fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = ABI drift, fail loudly
        fn.restype = res
        fn.argtypes = args
    v = lib.rec_abi_version()
    if v != REC_ABI_VERSION:
        raise RuntimeError(f"librecsys_b200.so ABI version {v} != binding version {REC_ABI_VERSION}")
    _lib = lib
    return lib


def last_error(lib, handle) -> str:
    msg = lib.rec_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(lib, handle, rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error(lib, handle)}")
