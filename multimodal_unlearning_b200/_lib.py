"""ctypes binding of include/clear_b200.h — the only place in the package that touches the C ABI.

There is deliberately no fallback: if libclear_b200.so is missing or a call fails, a RuntimeError is
raised.  (The CPU restatement under oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libclear_b200.so"

# enum ClearLossKind
KIND_CE, KIND_KL_FWD, KIND_KL_REV, KIND_NPO, KIND_ENTROPY = range(5)

EXPORTED_SYMBOLS = (
    "clear_last_error", "clear_abi_version", "clear_device_info", "clear_gemm_bf16", "clear_head_vtiles",
    "clear_head_project", "clear_head_lse_partials", "clear_head_rowloss", "clear_head_count_valid", "clear_head_reduce_rows",
    "clear_head_reduce_sequences", "clear_head_backward", "clear_saliency_accumulate",
    "clear_saliency_histogram", "clear_saliency_select", "clear_saliency_chunk_elems", "clear_saliency_mask",
    "clear_masked_adamw", "clear_reduce_partials_elems", "clear_grad_sqnorm", "clear_clip_coef", "clear_param_l1l0",
    "clear_param_l1_grad", "clear_grad_scale", "clear_saliency_mask_bits", "clear_mask_pack", "clear_mask_unpack",
    "clear_head_dweight_slab", "clear_dweight_allgather", "clear_head_compact_rows", "clear_rows_gather", "clear_rows_scatter",
    "clear_head_loss_workspace_bytes", "clear_head_loss_fwd_bwd",
    # ABI 3
    "clear_head_lse_partials_f32", "clear_rows_gather_dev", "clear_rows_scatter_dev", "clear_index_compose",
    "clear_saliency_hist_total", "clear_saliency_select_global_tau", "clear_saliency_select_global_fill",
    "clear_adamw8_block_elems", "clear_masked_adamw_8bit", "clear_head_project_pair",
    "clear_event_create", "clear_event_elapsed_ms", "clear_event_destroy",
)
ABI_VERSION = 3


class GemmArgs(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("d", C.c_void_p), ("lse_part", C.c_void_p),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
                ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
                ("a_major", C.c_int32), ("b_major", C.c_int32),
                ("d_f32", C.c_int32), ("accumulate", C.c_int32),
                ("cta_group", C.c_int32), ("raster_m_fast", C.c_int32), ("k_splits", C.c_int32)]


class ProjectArgs(C.Structure):
    _fields_ = [("hidden", C.c_void_p), ("weight", C.c_void_p), ("logits", C.c_void_p), ("lse_part", C.c_void_p),
                ("hidden_ld", C.c_int64), ("weight_ld", C.c_int64), ("logits_ld", C.c_int64),
                ("n_rows", C.c_int64), ("hidden_dim", C.c_int32), ("vocab", C.c_int32), ("cta_group", C.c_int32),
                ("n_rows_dev", C.c_void_p), ("n_rows_dev_offset", C.c_int32)]


class RowLossArgs(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("lse_part", C.c_void_p), ("ref_logits", C.c_void_p),
                ("ref_lse_part", C.c_void_p), ("labels", C.c_void_p), ("row_weight", C.c_void_p),
                ("scale_dev", C.c_void_p), ("loss_rows", C.c_void_p),
                ("logits_ld", C.c_int64), ("ref_ld", C.c_int64), ("n_rows", C.c_int64),
                ("vocab", C.c_int32), ("kind", C.c_int32), ("grad_scale", C.c_float), ("beta", C.c_float),
                ("write_grad", C.c_int32), ("ref_logits_f32", C.c_int32), ("n_rows_dev", C.c_void_p),
                ("n_rows_dev_offset", C.c_int32), ("status_flags", C.c_void_p)]


class BackwardArgs(C.Structure):
    _fields_ = [("dlogits", C.c_void_p), ("hidden", C.c_void_p), ("weight", C.c_void_p),
                ("dhidden", C.c_void_p), ("dweight", C.c_void_p),
                ("dlogits_ld", C.c_int64), ("hidden_ld", C.c_int64), ("weight_ld", C.c_int64),
                ("dhidden_ld", C.c_int64), ("dweight_ld", C.c_int64), ("n_rows", C.c_int64),
                ("hidden_dim", C.c_int32), ("vocab", C.c_int32), ("dweight_f32", C.c_int32),
                ("accumulate", C.c_int32), ("cta_group", C.c_int32), ("dhidden_ws", C.c_void_p),
                ("dweight_peers", C.c_void_p * 8), ("n_peers", C.c_int32),
                ("n_rows_dev", C.c_void_p), ("n_rows_dev_offset", C.c_int32)]


REDUCE_REFERENCE, REDUCE_SUM, REDUCE_SEQ_MEAN = 0, 1, 2


class HeadArgs(C.Structure):
    _fields_ = [("hidden", C.c_void_p), ("weight", C.c_void_p), ("labels", C.c_void_p), ("ref_hidden", C.c_void_p),
                ("ref_weight", C.c_void_p), ("ref_logits", C.c_void_p), ("row_weight", C.c_void_p), ("loss", C.c_void_p),
                ("loss_rows", C.c_void_p), ("seq_loss", C.c_void_p), ("dhidden", C.c_void_p), ("dweight", C.c_void_p),
                ("dweight_peers", C.c_void_p * 8), ("n_peers", C.c_int32),
                ("hidden_ld", C.c_int64), ("weight_ld", C.c_int64), ("ref_hidden_ld", C.c_int64), ("ref_weight_ld", C.c_int64),
                ("ref_logits_ld", C.c_int64), ("dhidden_ld", C.c_int64), ("dweight_ld", C.c_int64), ("n_rows", C.c_int64),
                ("hidden_dim", C.c_int32), ("vocab", C.c_int32), ("kind", C.c_int32), ("beta", C.c_double), ("coef", C.c_double),
                ("reduction", C.c_int32), ("n_seqs", C.c_int32), ("block_rows", C.c_int32), ("dweight_f32", C.c_int32),
                ("dweight_accumulate", C.c_int32), ("cta_group", C.c_int32),
                ("row_index", C.c_void_p), ("n_src_rows", C.c_int64), ("compact_ignored", C.c_int32),
                ("ref_logits_f32", C.c_int32), ("status_flags", C.c_void_p), ("dweight_ready_event", C.c_void_p),
                ("stage_events", C.c_void_p), ("n_stage_events", C.c_int32)]


class AdamWArgs(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("mask", C.c_void_p), ("n", C.c_int64),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("step", C.c_int32), ("grad_scale_dev", C.c_void_p),
                ("mask_bits", C.c_void_p)]


class AdamW8Args(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("state1", C.c_void_p), ("state2", C.c_void_p),
                ("absmax1", C.c_void_p), ("absmax2", C.c_void_p), ("code1", C.c_void_p), ("code2", C.c_void_p),
                ("qlut1", C.c_void_p), ("qlut2", C.c_void_p), ("mask", C.c_void_p), ("mask_bits", C.c_void_p), ("n", C.c_int64),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("step", C.c_int32), ("grad_scale_dev", C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    """Load libclear_b200.so (built by `python -m multimodal_unlearning_b200.build`).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m multimodal_unlearning_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the CLEAR loss head.")
    lib = C.CDLL(str(LIB_PATH))
    lib.clear_last_error.restype = C.c_char_p
    lib.clear_abi_version.restype = C.c_int
    lib.clear_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    lib.clear_gemm_bf16.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
    lib.clear_head_vtiles.argtypes = [C.c_int]
    lib.clear_head_project.argtypes = [C.POINTER(ProjectArgs), C.c_void_p]
    lib.clear_head_project_pair.argtypes = [C.POINTER(ProjectArgs), C.POINTER(ProjectArgs), C.c_void_p]
    lib.clear_head_lse_partials.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.clear_head_rowloss.argtypes = [C.POINTER(RowLossArgs), C.c_void_p]
    lib.clear_head_count_valid.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.clear_head_reduce_rows.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.clear_head_reduce_sequences.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.clear_head_backward.argtypes = [C.POINTER(BackwardArgs), C.c_void_p]
    lib.clear_saliency_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.clear_saliency_histogram.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
    lib.clear_saliency_select.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.c_int32,
                                          C.c_void_p, C.c_void_p]
    lib.clear_saliency_chunk_elems.argtypes = []
    lib.clear_saliency_chunk_elems.restype = C.c_int64
    lib.clear_saliency_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                                        C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.clear_masked_adamw.argtypes = [C.POINTER(AdamWArgs), C.c_void_p]
    lib.clear_reduce_partials_elems.argtypes = []
    lib.clear_reduce_partials_elems.restype = C.c_int64
    lib.clear_grad_sqnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.clear_saliency_mask_bits.argtypes = lib.clear_saliency_mask.argtypes
    lib.clear_mask_pack.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.clear_mask_unpack.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.clear_clip_coef.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.clear_param_l1l0.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.clear_param_l1_grad.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.clear_grad_scale.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.clear_head_dweight_slab.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.clear_dweight_allgather.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int32,
                                            C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    lib.clear_head_compact_rows.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.clear_rows_gather.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
    lib.clear_rows_scatter.argtypes = lib.clear_rows_gather.argtypes
    lib.clear_head_lse_partials_f32.argtypes = lib.clear_head_lse_partials.argtypes
    lib.clear_rows_gather_dev.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_int64, C.c_void_p]
    lib.clear_rows_scatter_dev.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                           C.c_void_p]
    lib.clear_index_compose.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.clear_saliency_hist_total.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.clear_saliency_select_global_tau.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_int32, C.c_void_p, C.c_void_p]
    lib.clear_saliency_select_global_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                      C.c_void_p, C.c_void_p]
    lib.clear_event_create.argtypes = [C.POINTER(C.c_void_p)]
    lib.clear_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
    lib.clear_event_destroy.argtypes = [C.c_void_p]
    lib.clear_adamw8_block_elems.restype = C.c_int32
    lib.clear_masked_adamw_8bit.argtypes = [C.POINTER(AdamW8Args), C.c_void_p]
    lib.clear_head_loss_workspace_bytes.argtypes = [C.POINTER(HeadArgs), C.POINTER(C.c_size_t)]
    lib.clear_head_loss_fwd_bwd.argtypes = [C.POINTER(HeadArgs), C.c_void_p, C.c_size_t, C.c_void_p]
    for name in EXPORTED_SYMBOLS:
        getattr(lib, name)  # AttributeError here = the .so is stale / does not match include/clear_b200.h
    if lib.clear_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libclear_b200.so ABI version {lib.clear_abi_version()} != {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().clear_last_error()
        raise RuntimeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")
