"""ctypes binding of the C-ABI library `libcsbsr_b200.so` (declared in include/csbsr_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a
`CsbsrError` is raised.  PyTorch is used only for device memory and streams.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CSBSR_LIB_PATH: debug override (instrumented builds of the same library, e.g. scripts/trace_conv.py)
LIB_PATH = os.environ.get("CSBSR_LIB_PATH") or os.path.join(_HERE, "lib", "libcsbsr_b200.so")

MAX_TAPS = 64
MAX_PHASES = 16
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2, 3
OUT_BF16_NHWC, OUT_F32_NCHW, OUT_F32_NHWC = 0, 1, 2


class CsbsrError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("x_pitch", C.c_int32), ("x_coff", C.c_int32), ("cin", C.c_int32),
        ("wgt", C.c_void_p),
        ("w_taps", C.c_int32), ("cout_pad", C.c_int32),
        ("nphases", C.c_int32), ("ntaps", C.c_int32), ("stride", C.c_int32),
        ("dh", C.c_int8 * MAX_TAPS), ("dw", C.c_int8 * MAX_TAPS),
        ("widx", C.c_int16 * MAX_TAPS),
        ("oh", C.c_int32), ("ow", C.c_int32),
        ("os", C.c_int32),
        ("ooh", C.c_int8 * MAX_PHASES), ("oow", C.c_int8 * MAX_PHASES),
        ("yh", C.c_int32), ("yw", C.c_int32),
        ("out_mode", C.c_int32),
        ("y", C.c_void_p),
        ("y_pitch", C.c_int32), ("y_coff", C.c_int32), ("cout_store", C.c_int32),
        ("bias", C.c_void_p),
        ("bias_sn", C.c_int32), ("bias_sc", C.c_int32), ("cls_bw", C.c_int32),
        ("act", C.c_int32),
        ("slope", C.c_float),
        ("r0", C.c_void_p), ("r0_pitch", C.c_int32), ("r0_coff", C.c_int32),
        ("rm", C.c_void_p), ("rm_pitch", C.c_int32), ("rm_coff", C.c_int32),
        ("r1", C.c_void_p), ("r1_pitch", C.c_int32), ("r1_coff", C.c_int32),
        ("r1_sign", C.c_float),
        ("r32", C.c_void_p),
        ("block_n", C.c_int32),
        ("r32_pitch", C.c_int32), ("r32_coff", C.c_int32),
        ("nsub", C.c_int32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("g", C.c_void_p),
        ("n", C.c_int32), ("gh", C.c_int32), ("gw", C.c_int32), ("g_pitch", C.c_int32), ("g_coff", C.c_int32),
        ("cg", C.c_int32),
        ("s", C.c_void_p),
        ("sh", C.c_int32), ("sw", C.c_int32), ("s_pitch", C.c_int32), ("s_coff", C.c_int32), ("cs", C.c_int32),
        ("ntaps", C.c_int32), ("stride", C.c_int32),
        ("dh", C.c_int8 * MAX_TAPS), ("dw", C.c_int8 * MAX_TAPS),
        ("wg", C.c_void_p),
        ("ws", C.c_void_p), ("ws_bytes", C.c_size_t),
        ("grad", C.c_void_p),
        ("grad_a", C.c_int32), ("grad_b", C.c_int32), ("grad_btot", C.c_int32), ("grad_b0", C.c_int32), ("grad_cp", C.c_int32),
    ]


_lib = None

# name -> (restype, argtypes); every symbol include/csbsr_b200.h declares
_SIGNATURES = {
    "csbsr_last_error": (C.c_char_p, []),
    "csbsr_version": (C.c_int, []),
    "csbsr_device_ok": (C.c_int, []),
    "csbsr_conv_igemm": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "csbsr_conv_wgrad": (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    "csbsr_conv_wgrad_workspace_bytes": (C.c_size_t, [C.POINTER(WgradDesc)]),
    "csbsr_tap_gather3x3": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "csbsr_blur_ps_bwd_input": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_blur_ps_bwd_kernel_workspace_bytes": (C.c_size_t, [C.c_int] * 6),
    "csbsr_blur_ps_bwd_kernel": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p, C.c_void_p]),
    "csbsr_resize_bicubic_aa_bwd": (C.c_int, [C.c_void_p] * 2 + [C.c_int] * 5 + [C.c_void_p]),
    "csbsr_pack_weights": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]),
    "csbsr_pack_weights_window": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]),
    "csbsr_pack_job_bytes": (C.c_size_t, []),
    "csbsr_pack_job_fill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_ulonglong]),
    "csbsr_pack_job_tiles": (C.c_longlong, [C.c_int] * 7),
    "csbsr_pack_weights_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_ulonglong, C.c_int, C.c_void_p]),
    "csbsr_wgrad_unpack_add": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_wgrad_unpack_add_tapexp": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_tapexp_gather_nhwc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_tapexp_scatter_nhwc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_patch_split": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_patch_join": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_crop_flip_u8": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_float, C.c_void_p]),
    "csbsr_bn_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_float, C.c_float] + [C.c_void_p] * 6),
    "csbsr_bn_apply": (C.c_int, [C.c_void_p] * 7 + [C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    "csbsr_bn_workspace_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_bn_backward": (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_longlong, C.c_int] + [C.c_void_p] * 6),
    "csbsr_psnr_ssim_workspace_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_psnr_ssim": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_size_t, C.c_void_p]),
    "csbsr_prelu_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "csbsr_prelu_bwd_workspace_bytes": (C.c_size_t, []),
    "csbsr_prelu_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_longlong, C.c_void_p, C.c_void_p]),
    "csbsr_adam_step": (C.c_int, [C.c_void_p] * 4 + [C.c_longlong] + [C.c_float] * 4 + [C.c_int, C.c_float, C.c_int,
                                                                                       C.c_void_p]),
    "csbsr_nchw_f32_to_nhwc_bf16": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]),
    "csbsr_patchify": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 12 + [C.c_void_p, C.c_void_p, C.c_int,
                                                                            C.c_void_p]),
    "csbsr_kpred_wpack_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_kpred_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "csbsr_kpred_sr_chain": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_float, C.c_void_p]),
    "csbsr_kpred_cat_chain": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_size_t] + [C.c_int] * 3 + [C.c_float, C.c_void_p]),
    "csbsr_gap_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    "csbsr_kernel_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "csbsr_vec_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "csbsr_broadcast_vec": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_blur_per_sample": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_bicubic_upsample": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    "csbsr_instnorm_workspace_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_clip_instnorm_stats": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_float, C.c_void_p]),
    "csbsr_maxpool3s2_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p]),
    "csbsr_adaptive_avgpool_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]),
    "csbsr_bilinear_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 11 + [C.c_void_p]),
    "csbsr_bilinear_f32": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_bilinear_f32_sigmoid": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_bilinear_add_nhwc": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 14 + [C.c_void_p]),
    "csbsr_softmax_gather": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p]),
    "csbsr_blur_kernel_synth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "csbsr_resize_bicubic_aa": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p]),
    "csbsr_degrade": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 7 + [C.c_void_p]),
    "csbsr_degrade_workspace_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_degrade_fused": (C.c_int, [C.c_void_p] * 5 + [C.c_size_t] + [C.c_int] * 7 + [C.c_void_p]),
    "csbsr_degrade_params_philox": (C.c_int, [C.c_void_p, C.c_int, C.c_ulonglong, C.c_ulonglong] + [C.c_double] * 4 + [C.c_void_p]),
    "csbsr_colsum_workspace_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_bias_grad": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "csbsr_act_bwd": (C.c_int, [C.c_void_p] * 3 + [C.c_longlong, C.c_float, C.c_void_p]),
    "csbsr_axpby": (C.c_int, [C.c_void_p] * 3 + [C.c_longlong, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "csbsr_sft_combine": (C.c_int, [C.c_void_p] * 4 + [C.c_longlong, C.c_void_p]),
    "csbsr_sft_combine_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_longlong, C.c_void_p]),
    "csbsr_window_copy": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p]),
    "csbsr_bilinear_nhwc_bwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 11 + [C.c_void_p]),
    "csbsr_adaptive_avgpool_nhwc_bwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]),
    "csbsr_maxpool3s2_nhwc_bwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 10 + [C.c_void_p]),
    "csbsr_dropout2d_mask": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_ulonglong, C.c_void_p, C.c_uint, C.c_void_p]),
    "csbsr_counter_inc": (C.c_int, [C.c_void_p, C.c_void_p]),
    "csbsr_channel_scale": (C.c_int, [C.c_void_p] * 3 + [C.c_int, C.c_longlong, C.c_int, C.c_void_p]),
    "csbsr_expand_classes": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    "csbsr_expand_classes_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "csbsr_expand_classes_bwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "csbsr_instnorm_apply": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_longlong, C.c_void_p]),
    "csbsr_instnorm_bwd_workspace_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_instnorm_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_longlong, C.c_void_p, C.c_size_t, C.c_void_p]),
    "csbsr_nhwc_bf16_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "csbsr_sdf_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "csbsr_sdf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "csbsr_seg_loss_workspace_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_seg_loss": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_float] * 3 + [C.c_void_p] * 5 +
                       [C.c_size_t, C.c_void_p]),
    "csbsr_seg_loss_wf_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "csbsr_seg_loss_wf_grad": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_void_p] * 4 +
                               [C.c_size_t, C.c_void_p]),
    "csbsr_seg_loss_wf_mean": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_void_p, C.c_void_p,
                                                                                              C.c_size_t, C.c_void_p]),
    "csbsr_sr_loss_workspace_bytes": (C.c_size_t, [C.c_int]),
    "csbsr_sr_loss_bwd": (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 4 + [C.c_float] * 3 + [C.c_void_p] * 4),
    "csbsr_sr_loss": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_float] * 3 + [C.c_void_p, C.c_void_p, C.c_size_t,
                                                                                 C.c_void_p]),
    "csbsr_metrics_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "csbsr_seg_metrics": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_double, C.c_void_p,
                                                                                         C.c_size_t, C.c_void_p]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load the C-ABI library (once). Raises CsbsrError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CsbsrError(
                "csbsr_b200: %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().csbsr_last_error()
        raise CsbsrError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# kernels launched per C-ABI call (our own __global__ functions only; memsets are not counted)
KERNELS_PER_CALL = {"csbsr_conv_igemm": 1, "csbsr_kpred_cat_chain": 2, "csbsr_clip_instnorm_stats": 2, "csbsr_degrade": 3, "csbsr_seg_metrics": 8}
LAUNCHES = 0


def count_launch(name, with_hd=True):
    global LAUNCHES
    n = KERNELS_PER_CALL.get(name, 1)
    if name == "csbsr_seg_metrics" and not with_hd:
        n = 2
    LAUNCHES += n
