"""ctypes binding of ``libmtn_b200.so`` (C ABI declared in ``include/mtn_b200.h``).

The product path has no fallback: if the shared library is missing or a call fails
the caller gets an exception carrying ``mtn_last_error()``.  PyTorch is used only
as the owner of device memory and streams -- tensors are passed as raw pointers.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmtn_b200.so")
ABI_VERSION = 7

ACT_NONE, ACT_RELU = 0, 1


class MtnError(RuntimeError):
    pass


class LinearArgs(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int), ("W", C.c_void_p), ("ldw", C.c_int),
                ("bias", C.c_void_p), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("act", C.c_int), ("addend", C.c_void_p), ("ld_add", C.c_int),
                ("add_period", C.c_int), ("out_f32", C.c_void_p), ("ld32", C.c_int),
                ("out_f16", C.c_void_p), ("ld16", C.c_int),
                ("batch", C.c_int), ("stride_A", C.c_longlong), ("stride_W", C.c_longlong),
                ("stride_bias", C.c_longlong), ("stride_add", C.c_longlong), ("stride_out_f32", C.c_longlong),
                ("stride_out_f16", C.c_longlong), ("out16_pre_add", C.c_int),
                ("drop_seed", C.c_void_p), ("drop_site", C.c_uint32), ("drop_thresh", C.c_uint32),
                ("drop_after_add", C.c_int)]


class GemmArgs(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_int), ("a_mn", C.c_int),
                ("B", C.c_void_p), ("ldb", C.c_int), ("b_mn", C.c_int),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
                ("alpha", C.c_void_p), ("bias", C.c_void_p), ("act", C.c_int),
                ("relu_mask", C.c_void_p), ("ld_mask", C.c_int),
                ("addend", C.c_void_p), ("ld_add", C.c_int), ("add_period", C.c_int),
                ("accumulate", C.c_int), ("out_f32", C.c_void_p), ("ld32", C.c_int),
                ("out_f16", C.c_void_p), ("ld16", C.c_int), ("out16_pre_add", C.c_int),
                ("batch", C.c_int), ("stride_A", C.c_longlong), ("stride_B", C.c_longlong),
                ("stride_bias", C.c_longlong), ("stride_add", C.c_longlong), ("stride_out_f32", C.c_longlong),
                ("stride_out_f16", C.c_longlong), ("mask_scale", C.c_float),
                ("drop_seed", C.c_void_p), ("drop_site", C.c_uint32), ("drop_thresh", C.c_uint32),
                ("drop_after_add", C.c_int), ("colsum_a", C.c_void_p), ("multimem", C.c_int)]


class LinearDgradArgs(C.Structure):
    _fields_ = [("dY", C.c_void_p), ("lddy", C.c_int), ("W", C.c_void_p), ("ldw", C.c_int),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("alpha", C.c_void_p),
                ("relu_mask", C.c_void_p), ("ld_mask", C.c_int), ("mask_scale", C.c_float),
                ("addend", C.c_void_p), ("ld_add", C.c_int),
                ("dX_f32", C.c_void_p), ("ld32", C.c_int), ("dX_f16", C.c_void_p), ("ld16", C.c_int),
                ("batch", C.c_int), ("stride_dY", C.c_longlong), ("stride_W", C.c_longlong),
                ("stride_add", C.c_longlong), ("stride_dX_f32", C.c_longlong), ("stride_dX_f16", C.c_longlong)]


class LinearWgradArgs(C.Structure):
    _fields_ = [("dY", C.c_void_p), ("lddy", C.c_int), ("X", C.c_void_p), ("ldx", C.c_int),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("alpha", C.c_void_p),
                ("dW", C.c_void_p), ("lddw", C.c_int),
                ("batch", C.c_int), ("stride_dY", C.c_longlong), ("stride_X", C.c_longlong),
                ("stride_dW", C.c_longlong), ("dbias", C.c_void_p), ("multimem", C.c_int)]


class LayerNormBwdArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("a_2", C.c_void_p), ("eps", C.c_float), ("rows", C.c_int), ("d", C.c_int),
                ("dy", C.c_void_p), ("dy_scale", C.c_void_p), ("dres", C.c_void_p), ("dx", C.c_void_p),
                ("da_2", C.c_void_p), ("db_2", C.c_void_p), ("param_alpha", C.c_void_p),
                ("dx_f16", C.c_void_p), ("dx_colsum", C.c_void_p),
                ("drop_seed", C.c_void_p), ("drop_site", C.c_uint32), ("drop_thresh", C.c_uint32), ("multimem", C.c_int)]


class EmbedBwdArgs(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("lut", C.c_void_p), ("pe", C.c_void_p),
                ("rows", C.c_int), ("L", C.c_int), ("d", C.c_int), ("vocab", C.c_int), ("scale", C.c_float),
                ("a_2", C.c_void_p), ("eps", C.c_float), ("dy", C.c_void_p),
                ("dlut", C.c_void_p), ("da_2", C.c_void_p), ("db_2", C.c_void_p), ("param_alpha", C.c_void_p),
                ("drop_seed", C.c_void_p), ("drop_site", C.c_uint32), ("drop_thresh", C.c_uint32)]


class AttnCoreBwdArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("ldq", C.c_int), ("k", C.c_void_p), ("ldk", C.c_int),
                ("v", C.c_void_p), ("ldv", C.c_int), ("dO", C.c_void_p), ("lddo", C.c_int),
                ("stats", C.c_void_p), ("delta", C.c_void_p),
                ("mask_bits", C.c_void_p), ("mask_rows_q", C.c_int),
                ("B", C.c_int), ("h", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int), ("d_k", C.c_int),
                ("dq", C.c_void_p), ("lddq", C.c_int), ("dk", C.c_void_p), ("lddk", C.c_int),
                ("dv", C.c_void_p), ("lddv", C.c_int),
                ("drop_seed", C.c_void_p), ("drop_site", C.c_uint32), ("drop_thresh", C.c_uint32),
                ("dq_f16", C.c_void_p), ("lddq16", C.c_int)]


class AttnCoreArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("ldq", C.c_int), ("k", C.c_void_p), ("ldk", C.c_int),
                ("v", C.c_void_p), ("ldv", C.c_int), ("mask_bits", C.c_void_p),
                ("mask_rows_q", C.c_int), ("B", C.c_int), ("h", C.c_int), ("Lq", C.c_int),
                ("Lk", C.c_int), ("d_k", C.c_int), ("out", C.c_void_p), ("ldo", C.c_int), ("stats", C.c_void_p),
                ("drop_seed", C.c_void_p), ("drop_site", C.c_uint32), ("drop_thresh", C.c_uint32),
                ("q_batch_stride", C.c_longlong), ("k_batch_stride", C.c_longlong), ("v_batch_stride", C.c_longlong),
                ("o_batch_stride", C.c_longlong)]


class AttnSiteArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int), ("d", C.c_int), ("h", C.c_int),
                ("x", C.c_void_p), ("x_out", C.c_void_p),
                ("ln_a", C.c_void_p), ("ln_b", C.c_void_p), ("ln_eps", C.c_float),
                ("w_q", C.c_void_p), ("b_q", C.c_void_p),
                ("w_kv", C.c_void_p), ("b_kv", C.c_void_p),
                ("w_o", C.c_void_p), ("b_o", C.c_void_p),
                ("mem_f16", C.c_void_p),
                ("kv", C.c_void_p), ("ld_kv", C.c_int), ("kv_k_col", C.c_int), ("kv_v_col", C.c_int),
                ("mask_bits", C.c_void_p), ("mask_rows_q", C.c_int),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class AttnSiteFusedArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int), ("d", C.c_int), ("h", C.c_int),
                ("xn_f16", C.c_void_p), ("ld_xn", C.c_int), ("x", C.c_void_p), ("ld_x", C.c_int),
                ("w_q", C.c_void_p), ("ld_wq", C.c_int), ("b_q", C.c_void_p),
                ("w_o", C.c_void_p), ("ld_wo", C.c_int), ("b_o", C.c_void_p),
                ("kv", C.c_void_p), ("ld_kv", C.c_int), ("kv_k_col", C.c_int), ("kv_v_col", C.c_int),
                ("mask_bits", C.c_void_p), ("mask_rows_q", C.c_int)]


class FfnArgs(C.Structure):
    _fields_ = [("rows", C.c_int), ("d", C.c_int), ("d_ff", C.c_int),
                ("x", C.c_void_p), ("x_out", C.c_void_p),
                ("ln_a", C.c_void_p), ("ln_b", C.c_void_p), ("ln_eps", C.c_float),
                ("w_1", C.c_void_p), ("b_1", C.c_void_p), ("w_2", C.c_void_p), ("b_2", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class DecodeSite(C.Structure):
    _fields_ = [("kind", C.c_int), ("ln_eps", C.c_float), ("ln_a", C.c_void_p), ("ln_b", C.c_void_p),
                ("w_in", C.c_void_p), ("b_in", C.c_void_p), ("w_out", C.c_void_p), ("b_out", C.c_void_p),
                ("k", C.c_void_p), ("v", C.c_void_p), ("q_cache", C.c_void_p),
                ("ld_kv", C.c_int), ("Lk", C.c_int), ("kv_batch_stride", C.c_longlong),
                ("mask_bits", C.c_void_p), ("mask_words", C.c_int)]


class DecodeClusterArgs(C.Structure):
    _fields_ = [("sites", C.POINTER(DecodeSite)), ("n_sites", C.c_int),
                ("B", C.c_int), ("d", C.c_int), ("h", C.c_int), ("d_ff", C.c_int), ("t", C.c_int),
                ("x_in", C.c_void_p), ("out", C.c_void_p),
                ("norm_a", C.c_void_p), ("norm_b", C.c_void_p), ("norm_eps", C.c_float),
                ("taps", C.c_void_p), ("stamps", C.c_void_p),
                ("gen_w", C.c_void_p), ("gen_b", C.c_void_p), ("gen_V", C.c_int), ("gen_V8", C.c_int),
                ("tokens", C.c_void_p), ("tokens_stride", C.c_longlong), ("rows_per_dialogue", C.c_int)]


# every symbol include/mtn_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mtn_abi_version": (C.c_int, []),
    "mtn_last_error": (C.c_char_p, []),
    "mtn_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "mtn_embed_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mtn_feature_prep_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mtn_feature_prep_f16_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mtn_log_softmax_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "mtn_label_smoothing_workspace_bytes": (C.c_size_t, [C.c_int]),
    "mtn_label_smoothing_loss_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_float,
                                               C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mtn_layernorm_grouped_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p]),
    "mtn_cast_f32_to_f16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "mtn_mask_words": (C.c_int, [C.c_int]),
    "mtn_mask_pack": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mtn_linear_fwd": (C.c_int, [C.POINTER(LinearArgs), C.c_void_p]),
    "mtn_ln_linear_supported": (C.c_int, [C.c_int]),
    "mtn_ln_linear_debug_timestamps": (C.c_int, [C.c_void_p]),
    "mtn_ln_linear_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "mtn_attn_core_fwd": (C.c_int, [C.POINTER(AttnCoreArgs), C.c_void_p]),
    "mtn_attn_site_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "mtn_attn_site_fwd": (C.c_int, [C.POINTER(AttnSiteArgs), C.c_void_p]),
    "mtn_rows_linear_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "mtn_rows_linear_fwd": (C.c_int, [C.POINTER(LinearArgs), C.c_void_p]),
    "mtn_rows_ln_linear_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "mtn_rows_ln_linear_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                         C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "mtn_decode_attn_supported": (C.c_int, [C.c_int, C.c_int]),
    "mtn_decode_attn_fwd": (C.c_int, [C.POINTER(AttnCoreArgs), C.c_void_p]),
    "mtn_attn_site_fused_supported": (C.c_int, [C.c_int, C.c_int]),
    "mtn_attn_site_fused_fwd": (C.c_int, [C.POINTER(AttnSiteFusedArgs), C.c_void_p]),
    "mtn_prog_begin": (C.c_int, []),
    "mtn_prog_recording": (C.c_int, []),
    "mtn_prog_stage_bytes": (C.c_int, []),
    "mtn_prog_end": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "mtn_prog_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mtn_decode_cluster_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "mtn_decode_cluster_max_sites": (C.c_int, []),
    "mtn_decode_cluster_fwd": (C.c_int, [C.POINTER(DecodeClusterArgs), C.c_void_p]),
    "mtn_ffn_fused_supported": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "mtn_ffn_fused_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "mtn_ffn_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "mtn_ffn_fwd": (C.c_int, [C.POINTER(FfnArgs), C.c_void_p]),
    "mtn_check_linear_fwd": (C.c_int, [C.POINTER(LinearArgs), C.c_void_p]),
    "mtn_check_attn_core_fwd": (C.c_int, [C.POINTER(AttnCoreArgs), C.c_void_p]),
    # ---- backward (ABI v3)
    "mtn_gemm_f16": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "mtn_check_gemm_f16": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "mtn_linear_dgrad": (C.c_int, [C.POINTER(LinearDgradArgs), C.c_void_p]),
    "mtn_linear_wgrad": (C.c_int, [C.POINTER(LinearWgradArgs), C.c_void_p]),
    "mtn_cast_colsum": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                  C.c_int, C.c_void_p]),
    "mtn_embed_dropout_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                        C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                        C.c_uint32, C.c_void_p]),
    "mtn_seed_bump": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtn_adam_advance": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "mtn_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                C.c_void_p]),
    "mtn_grad_absmax": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "mtn_grad_scale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mtn_zero": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "mtn_scale_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "mtn_layernorm_bwd": (C.c_int, [C.POINTER(LayerNormBwdArgs), C.c_void_p]),
    "mtn_embed_bwd": (C.c_int, [C.POINTER(EmbedBwdArgs), C.c_void_p]),
    "mtn_attn_delta": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]),
    "mtn_attn_core_bwd": (C.c_int, [C.POINTER(AttnCoreBwdArgs), C.c_void_p]),
    "mtn_log_softmax_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                      C.c_void_p]),
    "mtn_label_smoothing_loss_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_float,
                                               C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t,
                                               C.c_void_p]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises MtnError when it has not been built --
    there is deliberately no Python/PyTorch fallback for the hot path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MtnError("%s not found: build it with `python -c 'import __graft_entry__ as g; "
                           "g.build()'` or `make -C mtn_b200/csrc`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.mtn_abi_version() != ABI_VERSION:
            raise MtnError("ABI mismatch: library %d, binding %d" % (L.mtn_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MtnError("mtn_b200 error %d: %s" % (rc, lib().mtn_last_error().decode()))


# ------------------------------------------------------------------------------
# optional per-launch trace (bench.py roofline leg): when TRACE is a list every kernel
# wrapper appends {name, flops, bytes, start, end} with CUDA events around the launch.
# ------------------------------------------------------------------------------
TRACE = None
# when RECORD is a list, every launch also appends (name, flops, bytes, relaunch) where relaunch()
# re-issues the identical kernel on the current stream -- bench.py replays the launches of one
# kernel type back to back inside a CUDA graph to time them without host gaps.
RECORD = None


def _launch(name, flops, nbytes, fn, keep=()):
    if RECORD is not None:      # `keep` pins the operand tensors for as long as the closure lives
        RECORD.append((name, flops, nbytes, lambda: check(fn()), keep))
    if TRACE is None:
        return check(fn())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = fn()
    e.record()
    TRACE.append({"name": name, "flops": flops, "bytes": nbytes, "start": s, "end": e})
    check(rc)


# Launch-stream override: kernels go to STREAM (a torch.cuda.Stream) while PyTorch's current stream -- and with it
# the caching allocator's pool -- stays the caller's.  Buffers are therefore always allocated stream-ordered on the
# caller's stream and side streams only launch kernels (the engines join every side stream before they return).
STREAM = None


class on_stream(object):
    def __init__(self, stream):
        self.stream = stream

    def __enter__(self):
        global STREAM
        self.prev, STREAM = STREAM, self.stream

    def __exit__(self, *exc):
        global STREAM
        STREAM = self.prev


def launch_stream():
    return STREAM if STREAM is not None else torch.cuda.current_stream()


def stream_ptr():
    return C.c_void_p(launch_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise MtnError("%s must be a CUDA tensor (the hot path has no CPU implementation)" % name)
    if t.dtype != dtype:
        raise MtnError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.stride(-1) != 1:
        raise MtnError("%s must have unit stride in its last dimension" % name)


# ------------------------------------------------------------------------------
# thin tensor-level wrappers (used by mtn.py, engine.py and the tests)
# ------------------------------------------------------------------------------
def layernorm(x, a_2, b_2, eps, out_f32=None, out_f16=None, rows_per_group=None):
    """x: [..., d] f32 contiguous.  Writes into out_f32 / out_f16 (same shape).
    rows_per_group: a_2 / b_2 are [groups, d] and row r uses parameter set r // rows_per_group."""
    _req(x, torch.float32, "x"); _req(a_2, torch.float32, "a_2"); _req(b_2, torch.float32, "b_2")
    _req(out_f32, torch.float32, "out_f32"); _req(out_f16, torch.float16, "out_f16")
    d = x.shape[-1]
    rows = x.numel() // d
    assert x.is_contiguous() and (out_f32 is None or out_f32.is_contiguous()) and \
        (out_f16 is None or out_f16.is_contiguous())
    nbytes = rows * d * (4 + (4 if out_f32 is not None else 0) + (2 if out_f16 is not None else 0))
    rpg = rows if rows_per_group is None else int(rows_per_group)
    assert a_2.is_contiguous() and b_2.is_contiguous() and a_2.numel() * rpg >= rows * d
    _launch("layernorm", 0, nbytes,
            lambda: lib().mtn_layernorm_grouped_fwd(ptr(x), ptr(a_2), ptr(b_2), float(eps), rows, d, rpg, ptr(out_f32),
                                                    ptr(out_f16), stream_ptr()),
            keep=(x, a_2, b_2, out_f32, out_f16))


def cast_f16(src, dst=None):
    """2-D (or flattened-to-2-D) f32 -> f16 copy; row strides are honoured."""
    _req(src, torch.float32, "src")
    s2 = src.reshape(-1, src.shape[-1]) if src.dim() != 2 else src
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.float16, device=src.device)
    _req(dst, torch.float16, "dst")
    d2 = dst.reshape(-1, dst.shape[-1]) if dst.dim() != 2 else dst
    assert d2.shape == s2.shape
    _launch("cast_f16", 0, s2.numel() * 6,
            lambda: lib().mtn_cast_f32_to_f16(ptr(s2), s2.stride(0), ptr(d2), d2.stride(0), s2.shape[0],
                                              s2.shape[1], stream_ptr()), keep=(s2, d2))
    return dst


def mask_words(Lk):
    return lib().mtn_mask_words(int(Lk))


def mask_pack(mask):
    """mask: bool/uint8 [B, rows_q, Lk] (rows_q == 1 for key-padding masks) -> int32 bit words
    [B, rows_q, words]."""
    assert mask.dim() == 3 and mask.is_cuda
    m8 = mask.to(torch.uint8).contiguous()
    B, R, Lk = m8.shape
    bits = torch.empty((B, R, mask_words(Lk)), dtype=torch.int32, device=mask.device)
    _launch("mask_pack", 0, m8.numel() + bits.numel() * 4,
            lambda: lib().mtn_mask_pack(ptr(m8), B, R, Lk, ptr(bits), stream_ptr()), keep=(m8, bits))
    return bits


# few-row dispatch (KV-cached decoding): linear() / attn_core() take the small-M kernels of csrc/decode_rows.cu when
# the shape qualifies and ROWS_KERNELS is on (engine.decode_step switches it on for its launches)
ROWS_KERNELS = False


def rows_linear_ok(M, N, K, addend, out_f32, add_period, drop, out16_pre_add):
    return (ROWS_KERNELS and bool(lib().mtn_rows_linear_supported(int(M), int(N), int(K))) and drop is None and not out16_pre_add and
            (addend is None or (out_f32 is not None and addend.data_ptr() == out_f32.data_ptr() and add_period == 0 and
                                addend.stride(0) == out_f32.stride(0))))


def linear(A, W, bias=None, act=ACT_NONE, addend=None, add_period=0, out_f32=None, out_f16=None,
           _check_kernel=False, out16_pre_add=False, drop=None, drop_after_add=False):
    """C = act(A W^T + bias) + addend.  A: [M, K] f16 (row stride allowed), W: [N, K] f16."""
    _req(A, torch.float16, "A"); _req(W, torch.float16, "W"); _req(bias, torch.float32, "bias")
    _req(addend, torch.float32, "addend"); _req(out_f32, torch.float32, "out_f32")
    _req(out_f16, torch.float16, "out_f16")
    assert A.dim() == 2 and W.dim() == 2 and A.shape[1] == W.shape[1]
    a = LinearArgs()
    a.A, a.lda, a.W, a.ldw = A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0)
    a.bias = bias.data_ptr() if bias is not None else None
    a.M, a.N, a.K, a.act = A.shape[0], W.shape[0], A.shape[1], act
    if addend is not None:
        assert addend.dim() == 2
        a.addend, a.ld_add, a.add_period = addend.data_ptr(), addend.stride(0), int(add_period)
    if out_f32 is not None:
        assert out_f32.dim() == 2 and tuple(out_f32.shape) == (a.M, a.N)
        a.out_f32, a.ld32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_f16 is not None:
        assert out_f16.dim() == 2 and tuple(out_f16.shape) == (a.M, a.N)
        a.out_f16, a.ld16 = out_f16.data_ptr(), out_f16.stride(0)
    a.out16_pre_add = 1 if out16_pre_add else 0
    _set_drop(a, drop)
    a.drop_after_add = 1 if drop_after_add else 0
    fn = lib().mtn_check_linear_fwd if _check_kernel else lib().mtn_linear_fwd
    if not _check_kernel and rows_linear_ok(a.M, a.N, a.K, addend, out_f32, add_period, drop, out16_pre_add):
        fn = lib().mtn_rows_linear_fwd
    nbytes = 2 * (a.M * a.K + a.N * a.K) + a.M * a.N * ((4 if out_f32 is not None else 0) +
                                                        (2 if out_f16 is not None else 0) +
                                                        (4 if addend is not None else 0))
    _launch("linear", 2 * a.M * a.N * a.K, nbytes, lambda: fn(C.byref(a), stream_ptr()),
            keep=(A, W, bias, addend, out_f32, out_f16, drop))


def ln_linear_supported(d):
    return bool(lib().mtn_ln_linear_supported(int(d)))


def ln_linear(x, a_2, b_2, eps, W, bias=None, act=ACT_NONE, out_f16=None):
    """out_f16 = act(LN(x) W^T + bias) in ONE launch (LayerNorm fused into the projection's operand staging).
    x: [M, d] f32 contiguous, W: [N, d] f16 (row stride allowed), out_f16: [M, N] f16 (row stride allowed).
    Same result as layernorm(out_f16=...) + linear(out_f16=...) (bit-identical at d = 256 / 512)."""
    _req(x, torch.float32, "x"); _req(a_2, torch.float32, "a_2"); _req(b_2, torch.float32, "b_2")
    _req(W, torch.float16, "W"); _req(bias, torch.float32, "bias"); _req(out_f16, torch.float16, "out_f16")
    assert x.dim() == 2 and x.is_contiguous() and W.dim() == 2 and W.shape[1] == x.shape[1]
    M, d = x.shape
    N = W.shape[0]
    assert out_f16.dim() == 2 and tuple(out_f16.shape) == (M, N)
    assert a_2.is_contiguous() and b_2.is_contiguous() and a_2.numel() >= d and b_2.numel() >= d
    _launch("ln_linear", 2 * M * N * d, M * d * 4 + N * d * 2 + M * N * 2,
            lambda: lib().mtn_ln_linear_fwd(ptr(x), ptr(a_2), ptr(b_2), float(eps), M, d, ptr(W), W.stride(0), ptr(bias), N,
                                            int(act), ptr(out_f16), out_f16.stride(0), stream_ptr()),
            keep=(x, a_2, b_2, W, bias, out_f16))


def rows_ln_linear_ok(M, N, d):
    # every CTA of a row block (N / 8 of them) recomputes the block's LayerNorm statistics: worth it for N <= 512
    # (3.9 us vs 1.8 + 3.3 us as two launches), not for wide outputs (N = 2048: 13 us vs 7.7; profiles/r02_decode_kernels.txt)
    return ROWS_KERNELS and N <= 512 and bool(lib().mtn_rows_ln_linear_supported(int(M), int(N), int(d)))


def rows_ln_linear(x, a_2, b_2, eps, W, bias=None, act=ACT_NONE, out_f16=None):
    """out_f16 [M, N] = act(LN(x) W^T + bias) for M <= 128 in ONE launch (csrc/decode_rows.cu); bit-identical to
    layernorm(out_f16=...) + linear(out_f16=...).  x: [M, d] f32 (row stride allowed), out_f16: row stride allowed."""
    _req(x, torch.float32, "x"); _req(a_2, torch.float32, "a_2"); _req(b_2, torch.float32, "b_2")
    _req(W, torch.float16, "W"); _req(bias, torch.float32, "bias"); _req(out_f16, torch.float16, "out_f16")
    assert x.dim() == 2 and W.dim() == 2 and W.shape[1] == x.shape[1] and out_f16.dim() == 2
    M, d = x.shape
    N = W.shape[0]
    assert tuple(out_f16.shape) == (M, N) and a_2.is_contiguous() and b_2.is_contiguous()
    _launch("rows_ln_linear", 2 * M * N * d, M * d * 4 + N * d * 2 + M * N * 2,
            lambda: lib().mtn_rows_ln_linear_fwd(ptr(x), x.stride(0), ptr(a_2), ptr(b_2), float(eps), M, d, ptr(W), W.stride(0),
                                                 ptr(bias), N, int(act), ptr(out_f16), out_f16.stride(0), stream_ptr()),
            keep=(x, a_2, b_2, W, bias, out_f16))


def attn_core(q, k, v, B, h, Lq, Lk, d_k, out, mask_bits=None, _check_kernel=False, stats=None, drop=None):
    """q: [B*Lq, >=h*d_k] f16 view (row stride = leading dimension), k/v: [B*Lk, ...];
    out: [B*Lq, >=h*d_k] f16.  mask_bits: output of mask_pack or None.
    Any of q / k / v / out may instead be 3-D [B, >=L, >=h*d_k] views (KV-cached decoding: strides (batch, row, 1);
    the first Lq / Lk rows of every batch element are used)."""
    a = AttnCoreArgs()
    strides = {}
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.float16, n)
        assert t.dim() in (2, 3)
        strides[n] = (t.stride(0), 0) if t.dim() == 2 else (t.stride(1), t.stride(0))
        if t.dim() == 3:
            assert t.shape[0] == B and t.shape[1] >= (Lq if n in ("q", "out") else Lk)
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = (q.data_ptr(), strides["q"][0], k.data_ptr(), strides["k"][0],
                                         v.data_ptr(), strides["v"][0])
    a.q_batch_stride, a.k_batch_stride, a.v_batch_stride, a.o_batch_stride = (strides["q"][1], strides["k"][1],
                                                                              strides["v"][1], strides["out"][1])
    if mask_bits is not None:
        assert mask_bits.dtype == torch.int32 and mask_bits.is_contiguous() and mask_bits.shape[0] == B
        assert mask_bits.shape[2] == mask_words(Lk)
        a.mask_bits, a.mask_rows_q = mask_bits.data_ptr(), mask_bits.shape[1]
    a.B, a.h, a.Lq, a.Lk, a.d_k = B, h, Lq, Lk, d_k
    a.out, a.ldo = out.data_ptr(), strides["out"][0]
    if stats is not None:      # [B, h, Lq, 2] f32 softmax statistics for the backward pass
        _req(stats, torch.float32, "stats")
        assert stats.is_contiguous() and stats.numel() == B * h * Lq * 2
        a.stats = stats.data_ptr()
    _set_drop(a, drop)
    fn = lib().mtn_check_attn_core_fwd if _check_kernel else lib().mtn_attn_core_fwd
    if not _check_kernel and ROWS_KERNELS and Lq <= 8 and d_k == 64 and stats is None and drop is None:
        fn = lib().mtn_decode_attn_fwd
    _launch("attn_core", 4 * B * h * Lq * Lk * d_k, 2 * h * d_k * B * (2 * Lq + 2 * Lk),
            lambda: fn(C.byref(a), stream_ptr()), keep=(q, k, v, out, mask_bits, stats, drop))


def attn_site_fused_supported(d, h):
    return bool(lib().mtn_attn_site_fused_supported(int(d), int(h)))


def attn_site_fused(xn16, x, w_q, b_q, w_o, b_o, kv, k_col, v_col, B, h, Lq, Lk, mask_bits=None):
    """x [B*Lq, d] f32 += Wo . attention(xn16 Wq^T + bq, K, V) + bo in ONE launch (csrc/site_fused.cu).
    xn16: [B*Lq, d] f16 = LayerNorm(x); kv: f16 [B*Lk, ld] with K at columns [k_col, k_col+d), V at [v_col, v_col+d);
    w_q / w_o: [d, d] f16 (row stride allowed); mask_bits: output of mask_pack or None."""
    for t, n in ((xn16, "xn16"), (w_q, "w_q"), (w_o, "w_o"), (kv, "kv")):
        _req(t, torch.float16, n)
        assert t.dim() == 2
    _req(x, torch.float32, "x"); _req(b_q, torch.float32, "b_q"); _req(b_o, torch.float32, "b_o")
    d = x.shape[1]
    assert x.dim() == 2 and x.shape[0] == B * Lq and tuple(xn16.shape) == (B * Lq, d) and kv.shape[0] == B * Lk
    assert tuple(w_q.shape) == (d, d) and tuple(w_o.shape) == (d, d) and b_q.numel() == d and b_o.numel() == d
    a = AttnSiteFusedArgs()
    a.B, a.Lq, a.Lk, a.d, a.h = B, Lq, Lk, d, h
    a.xn_f16, a.ld_xn, a.x, a.ld_x = xn16.data_ptr(), xn16.stride(0), x.data_ptr(), x.stride(0)
    a.w_q, a.ld_wq, a.b_q = w_q.data_ptr(), w_q.stride(0), b_q.data_ptr()
    a.w_o, a.ld_wo, a.b_o = w_o.data_ptr(), w_o.stride(0), b_o.data_ptr()
    a.kv, a.ld_kv, a.kv_k_col, a.kv_v_col = kv.data_ptr(), kv.stride(0), int(k_col), int(v_col)
    if mask_bits is not None:
        assert mask_bits.dtype == torch.int32 and mask_bits.is_contiguous() and mask_bits.shape[0] == B
        assert mask_bits.shape[2] == mask_words(Lk)
        a.mask_bits, a.mask_rows_q = mask_bits.data_ptr(), mask_bits.shape[1]
    rows = B * Lq
    _launch("attn_site_fused", 4 * rows * d * d + 4 * B * Lq * Lk * d, rows * d * (2 + 8) + 2 * B * Lk * 2 * d + 4 * d * d,
            lambda: lib().mtn_attn_site_fused_fwd(C.byref(a), stream_ptr()),
            keep=(xn16, x, w_q, b_q, w_o, b_o, kv, mask_bits))


class StepProgram(object):
    """A decoding-step program (csrc/decode_rows.cu): the few-row launches of one KV-cached step as the stage list of ONE
    persistent cooperative kernel.  ``with prog.record():`` runs the step's Python code -- the few-row entry points and
    ``layernorm`` append stages instead of launching (host-side only: legal during CUDA-graph capture); ``prog.launch()``
    then runs them (graph-capturable: the stage list is uploaded from pinned memory with a stream-ordered copy).
    The constructor allocates (pinned host + device buffers): construct OUTSIDE graph capture."""

    MAX_STAGES = 256

    def __init__(self):
        self.n = 0
        nbytes = self.MAX_STAGES * int(lib().mtn_prog_stage_bytes())
        self.host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        self.counter = torch.zeros(4, dtype=torch.int32, device="cuda")
        self.keep = []
        self._uploaded = False

    class _Rec(object):
        def __init__(self, prog):
            self.prog = prog

        def __enter__(self):
            global RECORD
            check(lib().mtn_prog_begin())
            self.prev, RECORD = RECORD, []        # RECORD pins the operand tensors of every recorded call
            return self.prog

        def __exit__(self, et, ev, tb):
            global RECORD
            rec, RECORD = RECORD, self.prev
            n = C.c_int(0)
            if et is not None:
                lib().mtn_prog_end(None, 0, C.byref(n))
                return False
            p = self.prog
            check(lib().mtn_prog_end(p.host.data_ptr(), p.host.numel(), C.byref(n)))
            p.n, p._uploaded = int(n.value), False
            p.keep = [r[4] for r in rec]
            return False

    def record(self):
        return self._Rec(self)

    def launch(self):
        if not self._uploaded:
            nbytes = self.n * int(lib().mtn_prog_stage_bytes())
            self.dev[:nbytes].copy_(self.host[:nbytes], non_blocking=True)
            self._uploaded = True
        check(lib().mtn_prog_launch(self.dev.data_ptr(), self.n, self.counter.data_ptr(), stream_ptr()))


def decode_cluster_supported(B, d, h, d_ff, n_sites):
    L = lib()
    return bool(L.mtn_decode_cluster_supported(int(B), int(d), int(h), int(d_ff))) and n_sites <= L.mtn_decode_cluster_max_sites()


class DecodeClusterPlan(object):
    """The sublayer list of one decoding state for ``mtn_decode_cluster_fwd`` (csrc/decode_cluster.cu: one KV-cached step
    = ONE kernel, one thread-block cluster per dialogue group).  Built once per state by engine.DecoderEngine; holds the
    host-side ``MtnDecodeSite`` array (passed by value into the kernel's parameter space at every launch, so nothing is
    uploaded and the launch is graph-capturable) and keeps every referenced tensor alive."""

    def __init__(self, B, d, h, d_ff, rows_per_dialogue=1):
        """B target rows = D dialogues x rows_per_dialogue hypotheses (beam search; greedy decoding: 1)."""
        self.B, self.d, self.h, self.d_ff, self.R = int(B), int(d), int(h), int(d_ff), int(rows_per_dialogue)
        assert self.R >= 1 and self.B % self.R == 0
        self.D = self.B // self.R
        self.sites, self.keep = [], []

    def _site(self, kind, ln, w_in, b_in, w_out, b_out):
        for t, n in ((w_in, "w_in"), (w_out, "w_out")):
            _req(t, torch.float16, n)
            assert t.is_contiguous(), n
        for t, n in ((ln[0], "a_2"), (ln[1], "b_2"), (b_in, "b_in"), (b_out, "b_out")):
            _req(t, torch.float32, n)
            assert t.is_contiguous(), n
        s = DecodeSite()
        s.kind, s.ln_eps, s.ln_a, s.ln_b = kind, float(ln[2]), ln[0].data_ptr(), ln[1].data_ptr()
        s.w_in, s.b_in, s.w_out, s.b_out = w_in.data_ptr(), b_in.data_ptr(), w_out.data_ptr(), b_out.data_ptr()
        self.keep += [ln[0], ln[1], w_in, b_in, w_out, b_out]
        self.sites.append(s)
        return s

    def self_attention(self, ln, w_qkv, b_qkv, w_o, b_o, cache):
        """cache: f16 [B, T_max, 3d] holding [Q|K|V] of every position decoded so far."""
        d = self.d
        _req(cache, torch.float16, "cache")
        assert cache.dim() == 3 and cache.shape[0] == self.B and cache.shape[2] == 3 * d and cache.is_contiguous()
        assert tuple(w_qkv.shape) == (3 * d, d) and tuple(w_o.shape) == (d, d) and b_qkv.numel() == 3 * d and b_o.numel() == d
        s = self._site(0, ln, w_qkv, b_qkv, w_o, b_o)
        s.q_cache, s.k, s.v = cache.data_ptr(), cache.data_ptr() + 2 * d, cache.data_ptr() + 4 * d
        s.ld_kv, s.Lk, s.kv_batch_stride = 3 * d, 0, cache.stride(0)
        self.keep.append(cache)

    def cross_attention(self, ln, w_q, b_q, w_o, b_o, kv, k_col, v_col, Lk, mask_bits=None):
        """kv: f16 [D*Lk, ld] with K at columns [k_col, k_col + d), V at [v_col, v_col + d) (the memory stage's hoisted
        projections, one set per DIALOGUE); mask_bits: mask_pack output [D, 1, words] or None."""
        d = self.d
        _req(kv, torch.float16, "kv")
        assert kv.dim() == 2 and kv.shape[0] == self.D * Lk and kv.stride(1) == 1
        assert tuple(w_q.shape) == (d, d) and tuple(w_o.shape) == (d, d) and b_q.numel() == d and b_o.numel() == d
        s = self._site(1, ln, w_q, b_q, w_o, b_o)
        s.k, s.v = kv.data_ptr() + 2 * int(k_col), kv.data_ptr() + 2 * int(v_col)
        s.ld_kv, s.Lk, s.kv_batch_stride = kv.stride(0), int(Lk), int(Lk) * kv.stride(0)
        if mask_bits is not None:
            assert mask_bits.dtype == torch.int32 and mask_bits.is_contiguous() and mask_bits.shape[0] == self.D
            assert mask_bits.shape[1] == 1 and mask_bits.shape[2] == mask_words(Lk)
            s.mask_bits, s.mask_words = mask_bits.data_ptr(), mask_bits.shape[2]
            self.keep.append(mask_bits)
        self.keep.append(kv)

    def feed_forward(self, ln, w_1, b_1, w_2, b_2):
        d, dff = self.d, self.d_ff
        assert tuple(w_1.shape) == (dff, d) and tuple(w_2.shape) == (d, dff) and b_1.numel() == dff and b_2.numel() == d
        self._site(2, ln, w_1, b_1, w_2, b_2)

    def finish(self, norm):
        self.arr = (DecodeSite * len(self.sites))(*self.sites)
        self.norm = norm
        self.keep += [norm[0], norm[1]]
        return self

    def step(self, t, x_in, out, taps=None, stamps=None, gen=None, tokens=None):
        """One position: x_in [B, d] f32 -> out [B, d] f32 (Decoder.norm applied); cache rows t are written.
        gen = (w f16 [V8, d], b f32 [V8], V) + tokens (int64 [B], any stride): the arg-max of the generator's logits of the
        output rows is taken in the same kernel (greedy decoding)."""
        _req(x_in, torch.float32, "x_in"); _req(out, torch.float32, "out")
        assert tuple(x_in.shape) == (self.B, self.d) and tuple(out.shape) == (self.B, self.d)
        assert x_in.is_contiguous() and out.is_contiguous()
        a = DecodeClusterArgs()
        a.sites, a.n_sites = self.arr, len(self.sites)
        a.B, a.d, a.h, a.d_ff, a.t = self.B, self.d, self.h, self.d_ff, int(t)
        a.rows_per_dialogue = self.R
        a.x_in, a.out = x_in.data_ptr(), out.data_ptr()
        a.norm_a, a.norm_b, a.norm_eps = self.norm[0].data_ptr(), self.norm[1].data_ptr(), float(self.norm[2])
        if taps is not None:
            _req(taps, torch.float32, "taps")
            assert taps.is_contiguous() and taps.numel() == len(self.sites) * self.B * self.d
            a.taps = taps.data_ptr()
        if stamps is not None:
            assert stamps.dtype == torch.int64 and stamps.is_cuda and stamps.numel() >= 8 * len(self.sites)
            a.stamps = stamps.data_ptr()
        if gen is not None:
            gw, gb, V = gen
            _req(gw, torch.float16, "gen_w"); _req(gb, torch.float32, "gen_b")
            assert gw.is_contiguous() and gw.shape[1] == self.d and gw.shape[0] % 8 == 0 and gb.numel() == gw.shape[0] >= V
            assert tokens is not None and tokens.dtype == torch.int64 and tokens.is_cuda and tokens.dim() == 1 and tokens.numel() == self.B
            a.gen_w, a.gen_b, a.gen_V, a.gen_V8 = gw.data_ptr(), gb.data_ptr(), int(V), gw.shape[0]
            a.tokens, a.tokens_stride = tokens.data_ptr(), tokens.stride(0)
        _launch("decode_cluster", 0, 0, lambda: lib().mtn_decode_cluster_fwd(C.byref(a), stream_ptr()),
                keep=(self, x_in, out, taps, stamps, gen, tokens))


def ffn_fused_supported(rows, d, d_ff):
    return bool(lib().mtn_ffn_fused_supported(int(rows), int(d), int(d_ff)))


def ffn_fused(xn16, x, w_1, b_1, w_2, b_2):
    """x [rows, d] f32 += relu(xn16 W1^T + b1) W2^T + b2 in ONE launch, the hidden activation stays on chip
    (csrc/ffn_fused.cu).  xn16: [rows, d] f16 = LayerNorm(x); w_1: [d_ff, d] f16; w_2: [d, d_ff] f16 (contiguous)."""
    for t, n in ((xn16, "xn16"), (w_1, "w_1"), (w_2, "w_2")):
        _req(t, torch.float16, n)
        assert t.dim() == 2
    _req(x, torch.float32, "x"); _req(b_1, torch.float32, "b_1"); _req(b_2, torch.float32, "b_2")
    rows, d = x.shape
    d_ff = w_1.shape[0]
    assert tuple(xn16.shape) == (rows, d) and tuple(w_1.shape) == (d_ff, d) and tuple(w_2.shape) == (d, d_ff)
    assert w_1.is_contiguous() and w_2.is_contiguous() and b_1.numel() == d_ff and b_2.numel() == d
    _launch("ffn_fused", 4 * rows * d * d_ff, rows * d * (2 + 8) + 4 * d * d_ff,
            lambda: lib().mtn_ffn_fused_fwd(xn16.data_ptr(), xn16.stride(0), x.data_ptr(), x.stride(0), rows, d, d_ff,
                                            w_1.data_ptr(), b_1.data_ptr(), w_2.data_ptr(), b_2.data_ptr(), stream_ptr()),
            keep=(xn16, x, w_1, b_1, w_2, b_2))


def embed(ids, lut, pe, scale, ln=None, out_f32=None, out_f16=None, drop=None):
    """ids: [B, L] int64; lut: [vocab, d] f32; pe: [>=L, d] f32 (the table, 2-D).  Returns / fills
    [B, L, d] = LN?(lut[ids] * scale + pe[:L]);  ln = (a_2, b_2, eps) or None."""
    assert ids.dtype == torch.int64 and ids.dim() == 2 and ids.is_cuda
    _req(lut, torch.float32, "lut"); _req(pe, torch.float32, "pe")
    _req(out_f32, torch.float32, "out_f32"); _req(out_f16, torch.float16, "out_f16")
    B, L = ids.shape
    d = lut.shape[1]
    idc = ids.contiguous()
    assert pe.dim() == 2 and pe.shape[0] >= L and pe.is_contiguous() and lut.is_contiguous()
    a2, b2, eps = ln if ln is not None else (None, None, 0.0)
    _launch("embed", 0, B * L * d * (8 + (4 if out_f32 is not None else 0) + (2 if out_f16 is not None else 0)),
            lambda: lib().mtn_embed_dropout_fwd(ptr(idc), ptr(lut), ptr(pe), B * L, L, d, lut.shape[0], float(scale), ptr(a2),
                                                ptr(b2), float(eps), ptr(out_f32), ptr(out_f16),
                                                ptr(drop[0]) if drop else None, drop[1] if drop else 0,
                                                drop[2] if drop else 0, stream_ptr()),
            keep=(idc, lut, pe, a2, b2, out_f32, out_f16, drop))


def feature_prep(ft, out_f16=None, out_f32=None):
    """ft: [B, L, F] f32 raw features -> (mask bool [B, 1, L], out).  Frames that are all 1.0 are padding
    (data_utils.py:29) and are zeroed (data_utils.py:30)."""
    ftc = ft.contiguous()
    B, L, F = ftc.shape
    mask = torch.empty(B, 1, L, dtype=torch.bool, device=ft.device)
    if ft.dtype == torch.float16:          # features stored / uploaded as f16: same result, half the bytes
        _req(ft, torch.float16, "ft")
        assert out_f32 is None
        out_f16 = out_f16 if out_f16 is not None else torch.empty(B, L, F, dtype=torch.float16, device=ft.device)
        _launch("feature_prep", 0, B * L * F * 4,
                lambda: lib().mtn_feature_prep_f16_fwd(ptr(ftc), B * L, F, ptr(mask), ptr(out_f16), stream_ptr()),
                keep=(ftc, mask, out_f16))
        return mask, out_f16
    _req(ft, torch.float32, "ft")
    if out_f16 is None and out_f32 is None:
        out_f16 = torch.empty(B, L, F, dtype=torch.float16, device=ft.device)
    _launch("feature_prep", 0, B * L * F * (4 + (2 if out_f16 is not None else 0) + (4 if out_f32 is not None else 0)),
            lambda: lib().mtn_feature_prep_fwd(ptr(ftc), B * L, F, ptr(mask), ptr(out_f16), ptr(out_f32), stream_ptr()),
            keep=(ftc, mask, out_f16, out_f32))
    return mask, (out_f16 if out_f16 is not None else out_f32)


def log_softmax(x, V, out=None, argmax=None):
    """x: [rows, ld>=V] f32 logits.  out: [rows, V] f32 log-probs (or None); argmax: [rows] int64 (or None)."""
    _req(x, torch.float32, "x"); _req(out, torch.float32, "out")
    assert x.dim() == 2 and (out is None or (out.dim() == 2 and out.shape[0] == x.shape[0]))
    assert argmax is None or (argmax.dtype == torch.int64 and argmax.is_contiguous())
    _launch("log_softmax", 0, x.shape[0] * V * 8,
            lambda: lib().mtn_log_softmax_fwd(ptr(x), x.stride(0), x.shape[0], V, ptr(out),
                                              out.stride(0) if out is not None else 0, ptr(argmax), stream_ptr()),
            keep=(x, out, argmax))


def label_smoothing_loss(logits, V, target, padding_idx, smoothing, loss, scale=1.0, accumulate=False):
    """loss[0] (+)= scale * KL-sum of the label-smoothed target vs softmax(logits[:, :V]) (label_smoothing.py:20-32).
    logits: [rows, ld >= V] f32 (logits or log-probs); target: [rows] int64; loss: 1-element f32 tensor."""
    _req(logits, torch.float32, "logits"); _req(loss, torch.float32, "loss")
    assert logits.dim() == 2 and target.dtype == torch.int64 and target.is_cuda and target.numel() == logits.shape[0]
    tc = target.contiguous()
    rows = logits.shape[0]
    nbytes = lib().mtn_label_smoothing_workspace_bytes(rows)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=logits.device)
    _launch("label_smoothing", 0, rows * V * 8,
            lambda: lib().mtn_label_smoothing_loss_fwd(ptr(logits), logits.stride(0), rows, V, ptr(tc), int(padding_idx),
                                                       float(smoothing), float(scale), 1 if accumulate else 0, ptr(loss),
                                                       ptr(ws), nbytes, stream_ptr()),
            keep=(logits, tc, loss, ws))


def linear_batched(A, W, bias=None, act=ACT_NONE, addend=None, out_f32=None, out_f16=None):
    """`batch` independent problems of one shape in ONE launch (strided batch): A [b, M, K] f16, W [b, N, K] f16,
    bias [b, N] f32, addend / out_f32 [b, M, N] f32, out_f16 [b, M, N] f16.  The last dimension must have unit
    stride; rows and batch may be strided views (e.g. a column block of a packed [b, M, 3d] buffer)."""
    for t, dt, n in ((A, torch.float16, "A"), (W, torch.float16, "W"), (bias, torch.float32, "bias"),
                     (addend, torch.float32, "addend"), (out_f32, torch.float32, "out_f32"), (out_f16, torch.float16, "out_f16")):
        _req(t, dt, n)
    assert A.dim() == 3 and W.dim() == 3 and A.shape[0] == W.shape[0] and A.shape[2] == W.shape[2]
    a = LinearArgs()
    a.batch, a.M, a.N, a.K, a.act = A.shape[0], A.shape[1], W.shape[1], A.shape[2], act
    a.A, a.lda, a.stride_A = A.data_ptr(), A.stride(1), A.stride(0)
    a.W, a.ldw, a.stride_W = W.data_ptr(), W.stride(1), W.stride(0)
    if bias is not None:
        assert tuple(bias.shape) == (a.batch, a.N)
        a.bias, a.stride_bias = bias.data_ptr(), bias.stride(0)
    if addend is not None:
        assert tuple(addend.shape) == (a.batch, a.M, a.N)
        a.addend, a.ld_add, a.stride_add = addend.data_ptr(), addend.stride(1), addend.stride(0)
    if out_f32 is not None:
        assert tuple(out_f32.shape) == (a.batch, a.M, a.N)
        a.out_f32, a.ld32, a.stride_out_f32 = out_f32.data_ptr(), out_f32.stride(1), out_f32.stride(0)
    if out_f16 is not None:
        assert tuple(out_f16.shape) == (a.batch, a.M, a.N)
        a.out_f16, a.ld16, a.stride_out_f16 = out_f16.data_ptr(), out_f16.stride(1), out_f16.stride(0)
    nbytes = a.batch * (2 * (a.M * a.K + a.N * a.K) + a.M * a.N * ((4 if out_f32 is not None else 0) +
                                                                      (2 if out_f16 is not None else 0) +
                                                                      (4 if addend is not None else 0)))
    _launch("linear", 2 * a.batch * a.M * a.N * a.K, nbytes, lambda: lib().mtn_linear_fwd(C.byref(a), stream_ptr()),
            keep=(A, W, bias, addend, out_f32, out_f16))


# ------------------------------------------------------------------------------
# dropout: a `drop` argument is None or a tuple (seed, site, thresh): `seed` a 1-element int64 device tensor read by
# the kernels (so a captured step can bump it), `site` the number of the dropout application inside one forward
# pass, `thresh` = round(p * 65536).  The backward kernels get the SAME tuple and regenerate the decisions.
# ------------------------------------------------------------------------------
def drop_cfg(seed, site, p):
    """(seed, site, thresh) for probability p, or None when p == 0."""
    th = int(round(float(p) * 65536.0))
    if th <= 0:
        return None
    if th >= 65536:
        raise MtnError("dropout probability %g must be < 1" % p)
    assert seed.dtype == torch.int64 and seed.is_cuda and seed.numel() == 1
    return (seed, int(site), th)


def _set_drop(a, drop):
    if drop is not None:
        a.drop_seed, a.drop_site, a.drop_thresh = drop[0].data_ptr(), drop[1], drop[2]


def adam_advance(state, noam=None):
    """state: 8-element f32 device tensor {lr, b1, b2, eps, bc1, bc2, step, -}; noam = (factor, model_size, warmup)."""
    _req(state, torch.float32, "state")
    f, ms, wu = noam if noam is not None else (0.0, 1.0, 1.0)
    _launch("adam_advance", 0, 32, lambda: lib().mtn_adam_advance(ptr(state), float(f), float(ms), float(wu), stream_ptr()),
            keep=(state,))


def adam_step(p, g, m, v, state, p_f16=None, zero_grad=True):
    """Fused Adam over flat f32 buffers (n % 4 == 0); optionally refreshes the f16 copy and zeroes the gradient."""
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v"), (state, "state")):
        _req(t, torch.float32, n)
        assert t.is_contiguous()
    _req(p_f16, torch.float16, "p_f16")
    n = p.numel()
    assert g.numel() == n and m.numel() == n and v.numel() == n and (p_f16 is None or p_f16.numel() == n)
    _launch("adam_step", 0, n * (28 + (4 if zero_grad else 0) + (2 if p_f16 is not None else 0)),
            lambda: lib().mtn_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_f16), n, ptr(state), 1 if zero_grad else 0,
                                        stream_ptr()), keep=(p, g, m, v, state, p_f16))


def seed_bump(seed):
    _launch("seed_bump", 0, 16, lambda: lib().mtn_seed_bump(ptr(seed), stream_ptr()), keep=(seed,))


# ------------------------------------------------------------------------------
# backward wrappers (ABI v3).  `scale2` is a 2-element f32 device tensor {S, 1/S} from grad_scale(); wrappers
# take views of it (scale2[0:1] / scale2[1:2]) or None as `scale` / `alpha` device scalars.
# ------------------------------------------------------------------------------
def gemm(A, B, M, N, K, a_mn=False, b_mn=False, alpha=None, bias=None, act=ACT_NONE, relu_mask=None, addend=None,
         add_period=0, accumulate=False, out_f32=None, out_f16=None, _check_kernel=False, mask_scale=0.0, drop=None,
         drop_after_add=False, colsum_a=None):
    """General tensor-core GEMM (see MtnGemmArgs); 2-D operands, row strides honoured."""
    _req(A, torch.float16, "A"); _req(B, torch.float16, "B"); _req(alpha, torch.float32, "alpha")
    _req(bias, torch.float32, "bias"); _req(relu_mask, torch.float16, "relu_mask"); _req(addend, torch.float32, "addend")
    _req(out_f32, torch.float32, "out_f32"); _req(out_f16, torch.float16, "out_f16")
    a = GemmArgs()
    a.A, a.lda, a.a_mn = A.data_ptr(), A.stride(0), 1 if a_mn else 0
    a.B, a.ldb, a.b_mn = B.data_ptr(), B.stride(0), 1 if b_mn else 0
    a.M, a.N, a.K, a.act = M, N, K, act
    assert tuple(A.shape) == ((K, M) if a_mn else (M, K)) and tuple(B.shape) == ((K, N) if b_mn else (N, K))
    a.alpha = alpha.data_ptr() if alpha is not None else None
    a.bias = bias.data_ptr() if bias is not None else None
    if relu_mask is not None:
        a.relu_mask, a.ld_mask = relu_mask.data_ptr(), relu_mask.stride(0)
    if addend is not None:
        a.addend, a.ld_add, a.add_period = addend.data_ptr(), addend.stride(0), int(add_period)
    a.accumulate = 1 if accumulate else 0
    if out_f32 is not None:
        assert tuple(out_f32.shape) == (M, N)
        a.out_f32, a.ld32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_f16 is not None:
        assert tuple(out_f16.shape) == (M, N)
        a.out_f16, a.ld16 = out_f16.data_ptr(), out_f16.stride(0)
    a.mask_scale = float(mask_scale)
    _set_drop(a, drop)
    a.drop_after_add = 1 if drop_after_add else 0
    if colsum_a is not None:
        _req(colsum_a, torch.float32, "colsum_a")
        assert colsum_a.is_contiguous() and colsum_a.numel() == M
        a.colsum_a = colsum_a.data_ptr()
    fn = lib().mtn_check_gemm_f16 if _check_kernel else lib().mtn_gemm_f16
    _launch("gemm", 2 * M * N * K, 2 * (M * K + N * K) + M * N * 4, lambda: fn(C.byref(a), stream_ptr()),
            keep=(A, B, alpha, bias, relu_mask, addend, out_f32, out_f16))


def linear_dgrad(dY, W, alpha=None, relu_mask=None, addend=None, out_f32=None, out_f16=None, mask_scale=0.0):
    """dX = alpha * dY W (* relu_mask > 0) (+ addend).  dY: [M, N] f16, W: [N, K] f16 (forward layout)."""
    _req(dY, torch.float16, "dY"); _req(W, torch.float16, "W"); _req(alpha, torch.float32, "alpha")
    _req(relu_mask, torch.float16, "relu_mask"); _req(addend, torch.float32, "addend")
    _req(out_f32, torch.float32, "out_f32"); _req(out_f16, torch.float16, "out_f16")
    assert dY.dim() == 2 and W.dim() == 2 and dY.shape[1] == W.shape[0]
    a = LinearDgradArgs()
    a.dY, a.lddy, a.W, a.ldw = dY.data_ptr(), dY.stride(0), W.data_ptr(), W.stride(0)
    a.M, a.N, a.K = dY.shape[0], W.shape[0], W.shape[1]
    a.alpha = alpha.data_ptr() if alpha is not None else None
    if relu_mask is not None:
        assert tuple(relu_mask.shape) == (a.M, a.K)
        a.relu_mask, a.ld_mask = relu_mask.data_ptr(), relu_mask.stride(0)
    a.mask_scale = float(mask_scale)
    if addend is not None:
        assert tuple(addend.shape) == (a.M, a.K)
        a.addend, a.ld_add = addend.data_ptr(), addend.stride(0)
    if out_f32 is not None:
        assert tuple(out_f32.shape) == (a.M, a.K)
        a.dX_f32, a.ld32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_f16 is not None:
        assert tuple(out_f16.shape) == (a.M, a.K)
        a.dX_f16, a.ld16 = out_f16.data_ptr(), out_f16.stride(0)
    nbytes = 2 * (a.M * a.N + a.N * a.K) + a.M * a.K * ((4 if out_f32 is not None else 0) + (2 if out_f16 is not None else 0))
    _launch("linear_dgrad", 2 * a.M * a.N * a.K, nbytes, lambda: lib().mtn_linear_dgrad(C.byref(a), stream_ptr()),
            keep=(dY, W, alpha, relu_mask, addend, out_f32, out_f16))


def linear_wgrad(dY, X, dW, alpha=None, dbias=None, mc=0):
    """dW += alpha * dY^T X (and dbias += alpha * column sums of dY).  dY: [M, N] f16, X: [M, K] f16, dW: [N, K] f32
    (row stride honoured), dbias: [N] f32 contiguous."""
    _req(dbias, torch.float32, "dbias")
    _req(dY, torch.float16, "dY"); _req(X, torch.float16, "X"); _req(dW, torch.float32, "dW")
    _req(alpha, torch.float32, "alpha")
    assert dY.dim() == 2 and X.dim() == 2 and dY.shape[0] == X.shape[0] and tuple(dW.shape) == (dY.shape[1], X.shape[1])
    a = LinearWgradArgs()
    a.dY, a.lddy, a.X, a.ldx = dY.data_ptr(), dY.stride(0), X.data_ptr(), X.stride(0)
    a.M, a.N, a.K = dY.shape[0], dY.shape[1], X.shape[1]
    a.alpha = alpha.data_ptr() if alpha is not None else None
    # mc: byte offset from the gradient buffer's local mapping to its NVLS multicast mapping (0: local accumulation)
    a.dW, a.lddw = dW.data_ptr() + mc, dW.stride(0)
    a.multimem = 1 if mc else 0
    if dbias is not None:
        assert dbias.is_contiguous() and dbias.numel() == a.N
        a.dbias = dbias.data_ptr() + mc
    _launch("linear_wgrad", 2 * a.M * a.N * a.K, 2 * a.M * (a.N + a.K) + 8 * a.N * a.K,
            lambda: lib().mtn_linear_wgrad(C.byref(a), stream_ptr()), keep=(dY, X, dW, alpha, dbias))


def cast_colsum(src, dst_f16=None, colsum=None, scale=None, alpha=None, relu_mask=None, drop=None, mc=0):
    """dst_f16 = f16(src * scale [masked]); colsum += alpha * column sums.  src: 2-D f32 or f16."""
    assert src.dim() == 2 and src.is_cuda and src.dtype in (torch.float32, torch.float16) and src.stride(1) == 1
    _req(dst_f16, torch.float16, "dst_f16"); _req(colsum, torch.float32, "colsum"); _req(scale, torch.float32, "scale")
    _req(alpha, torch.float32, "alpha"); _req(relu_mask, torch.float16, "relu_mask")
    rows, cols = src.shape
    assert dst_f16 is None or tuple(dst_f16.shape) == (rows, cols)
    assert colsum is None or (colsum.is_contiguous() and colsum.numel() == cols)
    esz = 2 if src.dtype == torch.float16 else 4
    _launch("cast_colsum", 0, rows * cols * (esz + (2 if dst_f16 is not None else 0)),
            lambda: lib().mtn_cast_colsum(ptr(src), 1 if src.dtype == torch.float16 else 0, src.stride(0), ptr(dst_f16),
                                          dst_f16.stride(0) if dst_f16 is not None else 0, ptr(relu_mask),
                                          relu_mask.stride(0) if relu_mask is not None else 0, rows, cols, ptr(scale),
                                          ptr(alpha), C.c_void_p(colsum.data_ptr() + mc) if colsum is not None else None,
                                          ptr(drop[0]) if drop else None,
                                          drop[1] if drop else 0, drop[2] if drop else 0, 1 if (mc and colsum is not None) else 0,
                                          stream_ptr()),
            keep=(src, dst_f16, colsum, scale, alpha, relu_mask, drop))


_SCALE_SLOTS = {}


def grad_scale(tensors):
    """{S, 1/S} (2-element f32 device tensor) for the f32 gradient tensors `tensors`: S = 2^k with
    max|g| * S in [128, 256).  No host synchronisation."""
    dev = tensors[0].device
    slot = _SCALE_SLOTS.get(dev)
    if slot is None:
        slot = _SCALE_SLOTS[dev] = torch.zeros(4, dtype=torch.int32, device=dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    for t in tensors:
        _req(t, torch.float32, "grad")
        tc = t if t.is_contiguous() else t.contiguous()
        assert tc.numel() % 4 == 0
        _launch("grad_absmax", 0, tc.numel() * 4,
                lambda tc=tc: lib().mtn_grad_absmax(ptr(tc), tc.numel(), ptr(slot), stream_ptr()), keep=(tc, slot))
    _launch("grad_scale", 0, 16, lambda: lib().mtn_grad_scale(ptr(slot), ptr(out), stream_ptr()), keep=(slot, out))
    return out


def zero(t):
    """Stream-ordered zero fill of a contiguous tensor on the launch stream."""
    assert t.is_contiguous() and t.is_cuda
    n = t.numel() * t.element_size()
    _launch("zero", 0, n, lambda: lib().mtn_zero(ptr(t), n, stream_ptr()), keep=(t,))
    return t


def scale_f32(x, alpha, y, accumulate=False):
    """y = (accumulate ? y : 0) + x * alpha (alpha: 1-element f32 device tensor or None)."""
    _req(x, torch.float32, "x"); _req(y, torch.float32, "y"); _req(alpha, torch.float32, "alpha")
    assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel() and x.numel() % 4 == 0
    _launch("scale_f32", 0, x.numel() * (12 if accumulate else 8),
            lambda: lib().mtn_scale_f32(ptr(x), ptr(alpha), ptr(y), x.numel(), 1 if accumulate else 0, stream_ptr()),
            keep=(x, alpha, y))


def layernorm_bwd(x, a_2, eps, dy, dx, dres=None, da_2=None, db_2=None, dy_scale=None, param_alpha=None, dx_f16=None,
                  dx_colsum=None, drop=None, mc=0):
    """dx = dres + dLN(dy * dy_scale); da_2 / db_2 += param_alpha * (...).  x, dy, dx: [rows, d] f32 contiguous."""
    for t, n in ((x, "x"), (a_2, "a_2"), (dy, "dy"), (dx, "dx"), (dres, "dres"), (da_2, "da_2"), (db_2, "db_2"),
                 (dy_scale, "dy_scale"), (param_alpha, "param_alpha")):
        _req(t, torch.float32, n)
    d = x.shape[-1]
    rows = x.numel() // d
    assert x.is_contiguous() and dy.is_contiguous() and dx.is_contiguous() and (dres is None or dres.is_contiguous())
    assert dy.numel() == x.numel() and dx.numel() == x.numel()
    a = LayerNormBwdArgs()
    a.x, a.a_2, a.eps, a.rows, a.d = x.data_ptr(), a_2.data_ptr(), float(eps), rows, d
    a.dy = dy.data_ptr()
    a.dy_scale = dy_scale.data_ptr() if dy_scale is not None else None
    a.dres = dres.data_ptr() if dres is not None else None
    a.dx = dx.data_ptr()
    a.da_2 = da_2.data_ptr() + mc if da_2 is not None else None
    a.db_2 = db_2.data_ptr() + mc if db_2 is not None else None
    a.multimem = 1 if mc else 0
    a.param_alpha = param_alpha.data_ptr() if param_alpha is not None else None
    _req(dx_f16, torch.float16, "dx_f16"); _req(dx_colsum, torch.float32, "dx_colsum")
    if dx_f16 is not None:
        assert dx_f16.is_contiguous() and dx_f16.numel() == x.numel()
        a.dx_f16 = dx_f16.data_ptr()
    if dx_colsum is not None:
        assert dx_colsum.is_contiguous() and dx_colsum.numel() == d
        a.dx_colsum = dx_colsum.data_ptr() + mc
    _set_drop(a, drop)
    _launch("layernorm_bwd", 0, rows * d * (12 + (4 if dres is not None else 0) + (2 if dx_f16 is not None else 0)),
            lambda: lib().mtn_layernorm_bwd(C.byref(a), stream_ptr()),
            keep=(x, a_2, dy, dx, dres, da_2, db_2, dy_scale, param_alpha, dx_f16, dx_colsum, drop))


def embed_bwd(ids, lut, pe, scale, dy, dlut, ln=None, da_2=None, db_2=None, param_alpha=None, drop=None):
    """Backward of embed(): dlut[ids] += scale * dpre (atomic), LN parameter gradients when ln is given."""
    assert ids.dtype == torch.int64 and ids.dim() == 2 and ids.is_cuda
    _req(lut, torch.float32, "lut"); _req(pe, torch.float32, "pe"); _req(dy, torch.float32, "dy")
    _req(dlut, torch.float32, "dlut")
    B, L = ids.shape
    d = lut.shape[1]
    idc = ids.contiguous()
    assert dy.is_contiguous() and dy.numel() == B * L * d and dlut.is_contiguous() and dlut.shape == lut.shape
    a = EmbedBwdArgs()
    a.ids, a.lut, a.pe = idc.data_ptr(), lut.data_ptr(), pe.data_ptr()
    a.rows, a.L, a.d, a.vocab, a.scale = B * L, L, d, lut.shape[0], float(scale)
    if ln is not None:
        a.a_2, a.eps = ln[0].data_ptr(), float(ln[2])
        a.da_2, a.db_2 = da_2.data_ptr(), db_2.data_ptr()
    a.dy, a.dlut = dy.data_ptr(), dlut.data_ptr()
    a.param_alpha = param_alpha.data_ptr() if param_alpha is not None else None
    _set_drop(a, drop)
    _launch("embed_bwd", 0, B * L * d * 16, lambda: lib().mtn_embed_bwd(C.byref(a), stream_ptr()),
            keep=(idc, lut, pe, dy, dlut, ln, da_2, db_2, param_alpha, drop))


def attn_delta(dO, O, B, Lq, h, d_k, delta):
    _req(dO, torch.float16, "dO"); _req(O, torch.float16, "O"); _req(delta, torch.float32, "delta")
    assert delta.is_contiguous() and delta.numel() == B * h * Lq
    _launch("attn_delta", 0, B * Lq * h * d_k * 4,
            lambda: lib().mtn_attn_delta(ptr(dO), dO.stride(0), ptr(O), O.stride(0), B, Lq, h, d_k, ptr(delta),
                                         stream_ptr()), keep=(dO, O, delta))


def attn_core_bwd(q, k, v, dO, stats, delta, B, h, Lq, Lk, d_k, dq, dk, dv, mask_bits=None, drop=None):
    """q/k/v/dO: f16 2-D views (row stride = leading dimension); stats [B,h,Lq,2], delta [B,h,Lq] f32;
    dq: f32 [B*Lq, >= h*d_k] ACCUMULATED (zero it first) -- or, for Lk <= 128, an f16 tensor written directly;
    dk/dv: f16 [B*Lk, >= h*d_k] written."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (dO, "dO"), (dk, "dk"), (dv, "dv")):
        _req(t, torch.float16, n)
        assert t.dim() == 2
    _req(stats, torch.float32, "stats"); _req(delta, torch.float32, "delta")
    assert dq.is_cuda and dq.dim() == 2 and dq.stride(1) == 1 and dq.dtype in (torch.float32, torch.float16)
    a = AttnCoreBwdArgs()
    a.q, a.ldq, a.k, a.ldk, a.v, a.ldv = q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0)
    a.dO, a.lddo = dO.data_ptr(), dO.stride(0)
    a.stats, a.delta = stats.data_ptr(), delta.data_ptr()
    if mask_bits is not None:
        assert mask_bits.dtype == torch.int32 and mask_bits.is_contiguous() and mask_bits.shape[0] == B
        a.mask_bits, a.mask_rows_q = mask_bits.data_ptr(), mask_bits.shape[1]
    a.B, a.h, a.Lq, a.Lk, a.d_k = B, h, Lq, Lk, d_k
    a.dk, a.lddk, a.dv, a.lddv = dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0)
    if dq.dtype == torch.float16:
        a.dq_f16, a.lddq16 = dq.data_ptr(), dq.stride(0)
    else:
        a.dq, a.lddq = dq.data_ptr(), dq.stride(0)
    _set_drop(a, drop)
    _launch("attn_core_bwd", 10 * B * h * Lq * Lk * d_k, 2 * h * d_k * B * (4 * Lq + 4 * Lk),
            lambda: lib().mtn_attn_core_bwd(C.byref(a), stream_ptr()),
            keep=(q, k, v, dO, stats, delta, dq, dk, dv, mask_bits, drop))


def log_softmax_bwd(y, dy, V, dz):
    """dz[:, :V] = dy - exp(y) * rowsum(dy); dz[:, V:] = 0.  y, dy: [rows, >=V] f32; dz: [rows, ld] f32."""
    _req(y, torch.float32, "y"); _req(dy, torch.float32, "dy"); _req(dz, torch.float32, "dz")
    rows = y.shape[0]
    _launch("log_softmax_bwd", 0, rows * V * 12,
            lambda: lib().mtn_log_softmax_bwd(ptr(y), y.stride(0), ptr(dy), dy.stride(0), rows, V, ptr(dz), dz.stride(0),
                                              stream_ptr()), keep=(y, dy, dz))


def label_smoothing_loss_bwd(z, V, target, padding_idx, smoothing, dz, gscale=1.0, gout=None):
    """dz = gscale * gout * d(label-smoothed KL sum)/dz for logits or log-probs z [rows, ld >= V]."""
    _req(z, torch.float32, "z"); _req(dz, torch.float32, "dz"); _req(gout, torch.float32, "gout")
    tc = target.contiguous()
    rows = z.shape[0]
    ws = torch.empty(256, dtype=torch.uint8, device=z.device)
    _launch("label_smoothing_bwd", 0, rows * V * 8,
            lambda: lib().mtn_label_smoothing_loss_bwd(ptr(z), z.stride(0), rows, V, ptr(tc), int(padding_idx),
                                                       float(smoothing), float(gscale), ptr(gout), ptr(dz), dz.stride(0),
                                                       ptr(ws), 256, stream_ptr()), keep=(z, tc, dz, gout, ws))
