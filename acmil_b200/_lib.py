"""ctypes binding of libacmil_b200.so (include/acmil_b200.h).

The library is built in-tree by ``acmil_b200.build`` (nvcc, sm_100a).  There is no
Python / CPU fallback: if the library cannot be loaded, or no CUDA device is present when a
compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.environ.get("ACMIL_B200_LIB_DIR") or os.path.join(_HERE, "lib"), "libacmil_b200.so")

MAX_BRANCH, MAX_MASKED, MAX_SLIDES, MAX_CLASS = 8, 32, 128, 16
ACT_TANH, ACT_RELU, ACT_GELU = 0, 1, 2
IMPL_AUTO, IMPL_FFMA, IMPL_UMMA = 0, 1, 2
ACT_IDS = {"tanh": ACT_TANH, "relu": ACT_RELU, "gelu": ACT_GELU}


def record_floats(n_branch: int, d_inner: int, n_masked: int) -> int:
    """4-byte units of one bag's partial record (GpRecord in csrc/gp_common.cuh: m, l, acc, cnt, score, idx, h; every
    section padded to a multiple of 4).  acmil_gp_sizes reports the same number times n_slides times 4 as partial_bytes."""
    pad4 = lambda n: (n + 3) & ~3  # noqa: E731
    k, l, nmc = n_branch, d_inner, max(n_masked, 0)
    return pad4(k) * 3 + pad4(k * l) + 2 * pad4(k * nmc) + pad4(k * nmc * l)


class AcmilError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"acmil_b200 error {code}: {msg}")
        self.code = code


class GpShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "d_in", "d_inner", "d_attn", "n_branch", "front", "front_bias", "front_act", "act_a",
        "gated", "gate_bias", "score_bias", "reserved")]


class GpWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_w1", "d_b1", "d_wv", "d_bv", "d_wu", "d_bu", "d_ww", "d_bw")]


class GpConsts(C.Structure):
    _fields_ = [("b1", C.c_float * 128), ("bv", C.c_float * 128), ("bu", C.c_float * 128),
                ("ww", (C.c_float * 128) * 8), ("bw", C.c_float * 8), ("inv_scale", C.c_float * 4),
                ("valid", C.c_int32), ("reserved", C.c_int32 * 3)]


class GpBatch(C.Structure):
    _fields_ = [("d_x", C.c_void_p), ("row_offsets", C.POINTER(C.c_int64)), ("n_slides", C.c_int32),
                ("n_masked", C.c_int32), ("shard_row_begin", C.POINTER(C.c_int64)), ("d_a_out", C.c_void_p),
                ("a_ld", C.c_int64), ("x_f16", C.c_int32), ("reserved", C.c_int32), ("d_z", C.c_void_p)]


class GpHeads(C.Structure):
    _fields_ = [("n_class", C.c_int32), ("n_branch_heads", C.c_int32), ("d_wc", C.c_void_p), ("d_bc", C.c_void_p),
                ("slide_head", C.c_int32), ("shared_head", C.c_int32), ("d_ws", C.c_void_p), ("d_bs", C.c_void_p)]


class GpOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_sub", "d_slide", "d_afeat", "d_bag_feat", "d_lse_m", "d_lse_l", "d_topk_idx", "d_masked_idx")]


MAX_PEERS = 16


class GpExchange(C.Structure):
    """acmil_gp_exchange (include/acmil_b200.h): peer-mapped gather buffers / flag words of every rank."""
    _fields_ = [("n_ranks", C.c_int32), ("rank", C.c_int32), ("d_gather", C.c_void_p * MAX_PEERS),
                ("d_flags", C.c_void_p * MAX_PEERS), ("d_epoch", C.c_void_p), ("d_ticket", C.c_void_p),
                ("gather_bytes", C.c_size_t)]


class GpBwdGateArgs(C.Structure):
    """acmil_gp_bwd_gate_args (include/acmil_b200.h)."""
    _fields_ = ([(n, C.c_void_p) for n in ("d_h", "d_z", "d_scores", "d_lse_m", "d_lse_l", "d_afeat", "d_g_afeat", "d_g_bag",
                                           "d_g_scores", "d_ww")]
                + [(n, C.c_int64) for n in ("n", "a_ld", "gs_ld", "ldt")]
                + [(n, C.c_int32) for n in ("d_inner", "d_attn", "n_branch", "act_a", "gated", "reserved")]
                + [(n, C.c_void_p) for n in ("d_dz", "d_dzt", "d_dhp", "d_partials", "d_small")])


class GemmDesc(C.Structure):
    """acmil_gemm_desc (include/acmil_transmil.h)."""
    _fields_ = ([(n, C.c_void_p) for n in ("a", "b", "c", "ct", "bias", "addend", "split_ws")]
                + [(n, C.c_int32) for n in ("m", "n", "k", "batch")]
                + [(n, C.c_int64) for n in ("lda", "ldb", "ldc", "ldct", "ld_addend", "a_batch_stride", "b_batch_stride",
                                            "c_batch_stride", "ct_batch_stride", "addend_batch_stride")]
                + [("col_block_width", C.c_int32), ("k_split", C.c_int32), ("col_block_stride", C.c_int64),
                   ("alpha", C.c_float), ("beta", C.c_float), ("diag", C.c_float), ("act", C.c_int32),
                   ("precise", C.c_int32), ("batch_inner", C.c_int32), ("bias_per_row", C.c_int32), ("reserved", C.c_int32)]
                + [(n, C.c_int64) for n in ("a_batch_stride2", "b_batch_stride2", "c_batch_stride2", "ct_batch_stride2",
                                            "addend_batch_stride2")]
                + [("b_split", C.c_void_p), ("b_split_rows", C.c_int32), ("b_split_row0", C.c_int32),
                   ("softmax_stats_out", C.c_void_p), ("softmax_stats_in", C.c_void_p)])


class NystromShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "n", "dim", "heads", "dim_head", "num_landmarks", "pinv_iterations", "residual", "conv_kernel",
        "n_out", "padded_out", "precise")] + [("reserved", C.c_int32 * 4)]


class NystromWeights(C.Structure):
    _fields_ = [("d_ln_w", C.c_void_p), ("d_ln_b", C.c_void_p), ("ln_eps", C.c_float), ("reserved", C.c_int32),
                ("d_wqkv", C.c_void_p), ("d_wout", C.c_void_p), ("d_bout", C.c_void_p), ("d_wconv", C.c_void_p),
                ("d_split_qkv", C.c_void_p), ("d_split_out", C.c_void_p)]


class NystromShard(C.Structure):
    """acmil_nystrom_shard (include/acmil_transmil.h)."""
    _fields_ = [(n, C.c_int32) for n in (
        "n_loc", "lead_zero", "dim", "heads", "dim_head", "num_landmarks", "m_loc", "group_len", "pinv_iterations", "residual",
        "conv_kernel", "precise", "n_out", "head_first", "head_count", "halo")]


class NystromShardBufs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_x", "d_residual", "d_out", "d_ql_loc", "d_kl_loc", "d_ql", "d_kl", "d_z", "d_kv_part", "d_st_m", "d_st_l", "d_kv",
        "d_vt_ext", "d_workspace")] + [("workspace_bytes", C.c_size_t)]


class VitShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("batch", "img", "patch", "in_ch", "dim", "depth", "heads", "mlp_dim", "n_class",
                                         "precise")] + [("ln_eps", C.c_float), ("reserved", C.c_int32 * 5)]


class VitBlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_ln1_w", "d_ln1_b", "d_qkv_w", "d_qkv_b", "d_proj_w", "d_proj_b", "d_ln2_w",
                                          "d_ln2_b", "d_fc1_w", "d_fc1_b", "d_fc2_w", "d_fc2_b", "d_split_qkv",
                                          "d_split_proj", "d_split_fc1", "d_split_fc2")]


class VitWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("d_cls_token", "d_pos_embed", "d_patch_w", "d_patch_b", "d_norm_w", "d_norm_b",
                                          "d_head_w", "d_head_b")] + [("blocks", C.POINTER(VitBlockWeights)),
                                                                     ("d_split_patch", C.c_void_p)]


# every symbol include/*.h declares: (restype, argtypes)
_SIZE_P = C.POINTER(C.c_size_t)
SYMBOLS = {
    "acmil_last_error": (C.c_char_p, []),
    "acmil_abi_version": (C.c_int, []),
    "acmil_device_count": (C.c_int, []),
    "acmil_launch_count": (C.c_int64, []),
    "acmil_prof_enable": (C.c_int, [C.c_int]),
    "acmil_prof_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "acmil_gp_packed_bytes": (C.c_int, [C.POINTER(GpShape), _SIZE_P]),
    "acmil_gp_pack": (C.c_int, [C.POINTER(GpShape), C.POINTER(GpWeights), C.c_void_p, C.c_size_t,
                                C.POINTER(GpConsts), C.c_void_p]),
    "acmil_gp_umma_supported": (C.c_int, [C.POINTER(GpShape)]),
    "acmil_gp_partial_x": (C.c_int, [C.POINTER(GpShape), C.c_void_p, C.POINTER(GpConsts), C.POINTER(GpBatch), C.c_int,
                                     C.c_void_p, C.c_size_t, C.POINTER(GpExchange), C.c_void_p]),
    "acmil_gp_finish_x": (C.c_int, [C.POINTER(GpShape), C.POINTER(GpBatch), C.POINTER(GpExchange), C.POINTER(C.c_int32),
                                    C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(GpHeads),
                                    C.POINTER(GpOutputs), C.c_void_p]),
    "acmil_gp_sizes": (C.c_int, [C.POINTER(GpShape), C.POINTER(GpBatch), C.c_int, _SIZE_P, _SIZE_P]),
    "acmil_gp_partial": (C.c_int, [C.POINTER(GpShape), C.c_void_p, C.POINTER(GpConsts), C.POINTER(GpBatch), C.c_int,
                                   C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "acmil_gp_overflow_flags": (C.c_int, [C.POINTER(GpShape), C.POINTER(GpBatch), C.c_int, C.c_void_p,
                                          C.POINTER(C.c_int32), C.c_void_p]),
    "acmil_gp_finish": (C.c_int, [C.POINTER(GpShape), C.POINTER(GpBatch), C.c_void_p, C.c_size_t, C.c_int,
                                  C.POINTER(C.c_int32), C.c_void_p, C.c_int32, C.POINTER(GpHeads),
                                  C.POINTER(GpOutputs), C.c_void_p]),
    "acmil_gp_finish_rand": (C.c_int, [C.POINTER(GpShape), C.POINTER(GpBatch), C.c_void_p, C.c_size_t, C.c_int,
                                       C.POINTER(C.c_int32), C.c_void_p, C.c_int32, C.c_int32, C.POINTER(GpHeads),
                                       C.POINTER(GpOutputs), C.c_void_p]),
    "acmil_gp_attn_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "acmil_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "acmil_gp_bwd_workspace_floats": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "acmil_gp_bwd_gate": (C.c_int, [C.POINTER(GpBwdGateArgs), C.c_void_p]),
    "acmil_gp_bwd_relu_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "acmil_transpose_f32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "acmil_div_loss_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "acmil_div_loss_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_void_p]),
    # include/acmil_transmil.h
    "acmil_gemm_nt": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "acmil_gemm_split_bytes": (C.c_int, [C.c_int32, C.c_int32, _SIZE_P]),
    "acmil_gemm_split_b": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "acmil_layernorm_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_float,
                                       C.c_void_p, C.c_int64, C.c_void_p]),
    "acmil_nystrom_workspace_bytes": (C.c_int, [C.POINTER(NystromShape), _SIZE_P]),
    "acmil_nystrom_attn_fwd": (C.c_int, [C.POINTER(NystromShape), C.POINTER(NystromWeights), C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "acmil_softmax_rows_inplace": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "acmil_nystrom_shard_workspace_bytes": (C.c_int, [C.POINTER(NystromShard), _SIZE_P]),
    "acmil_nystrom_shard_phase": (C.c_int, [C.POINTER(NystromShard), C.POINTER(NystromWeights), C.POINTER(NystromShardBufs),
                                            C.c_int32, C.c_void_p]),
    "acmil_lse_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p]),
    "acmil_ppeg_fwd_rows": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "acmil_vit_workspace_bytes": (C.c_int, [C.POINTER(VitShape), _SIZE_P]),
    "acmil_vit_fwd": (C.c_int, [C.POINTER(VitShape), C.POINTER(VitWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
    "acmil_preprocess_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _SIZE_P]),
    "acmil_preprocess_u8": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "acmil_f32_to_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    # include/acmil_resnet.h
    "acmil_im2col": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 10 + [C.c_void_p]),
    "acmil_maxpool_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p]),
    "acmil_avgpool_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "acmil_ppeg_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


def load(build_if_missing: bool = True):
    """Loads (building first if the .so is absent and nvcc is available) and returns the CDLL."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        # (ACMIL_B200_NO_REBUILD=1: load what is there -- variant builds made with other ACMIL_NVCC_EXTRA flags)
        stale = (os.path.exists(LIB_PATH) and os.environ.get("ACMIL_B200_NO_REBUILD") != "1"
                 and not _build.up_to_date())
        if not os.path.exists(LIB_PATH) or stale:
            # missing, or built from other sources / flags than the ones in the tree (the stamp holds their digest)
            if not build_if_missing:
                raise ImportError(f"{LIB_PATH} is {'stale' if stale else 'missing'}; run `python -m acmil_b200.build`")
            try:
                _build.build()
            except RuntimeError:
                if not stale:
                    raise
                import warnings
                warnings.warn(f"{LIB_PATH} does not match the sources and could not be rebuilt; loading it as it is")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.acmil_abi_version() != 1:
            raise ImportError("libacmil_b200.so ABI version mismatch")
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        raise AcmilError(rc, load().acmil_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(load().acmil_launch_count())
