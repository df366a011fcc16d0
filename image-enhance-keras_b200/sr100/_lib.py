"""ctypes binding of libsr100.so (C ABI declared in include/sr100.h).

There is no CPU fallback: if the library is missing or the device is not sm_100, the engine raises.
PyTorch is used only for device memory and streams; every compute call goes through this module.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SR100_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libsr100.so")


class SrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sr100 error %d: %s" % (code, msg))
        self.code = code


class ConvDesc(C.Structure):
    _fields_ = [
        ("nsrc", C.c_int),
        ("in_", C.c_void_p * 2),
        ("wpacked", C.c_void_p * 2),
        ("ksize", C.c_int * 2),
        ("NB", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("cin", C.c_int), ("cout", C.c_int),
        ("bias", C.c_void_p),
        ("alpha", C.c_float), ("beta", C.c_float),
        ("relu", C.c_int),
        ("res_f32", C.c_void_p),
        ("res_bf16", C.c_void_p),
        ("out_bf16", C.c_void_p),
        ("out_f32", C.c_void_p),
        ("relu_mask_bf16", C.c_void_p),
        ("a_mode", C.c_int),
        ("nacc", C.c_int),
        ("pair", C.c_int),
        ("out_index", C.c_void_p),
        ("out_h", C.c_int), ("out_w", C.c_int),
        ("shuffle_r", C.c_int), ("shuffle_order", C.c_int),
        ("comp_h", C.c_int), ("comp_w", C.c_int),
        ("leaky_slope", C.c_float),
        ("precision", C.c_int),
        ("out_tf32", C.c_void_p),
        ("colsum_f32", C.c_void_p),
        ("colsum_scale", C.c_float),
        ("stitch_tiles", C.c_void_p),
        ("stitch_u8", C.c_void_p),
        ("stitch_mul", C.c_float),
        ("cin_valid", C.c_int * 2),
        ("relu_mask_slope", C.c_float),
    ]


class StitchTile(C.Structure):
    _fields_ = [("img_offset", C.c_longlong), ("img_h", C.c_int), ("img_w", C.c_int), ("y0", C.c_int), ("x0", C.c_int),
                ("oy0", C.c_int), ("oy1", C.c_int), ("ox0", C.c_int), ("ox1", C.c_int)]


class ConvPlanInfo(C.Structure):
    _fields_ = [
        ("flops", C.c_double),
        ("mma_efficiency", C.c_double),
        ("total_tiles", C.c_int), ("grid", C.c_int), ("smem_bytes", C.c_int),
        ("seg_width", C.c_int), ("nseg", C.c_int), ("strip_rows", C.c_int),
        ("num_wstages", C.c_int), ("tile_positions", C.c_int),
    ]


class PackItem(C.Structure):
    _fields_ = [("hwio", C.c_void_p), ("dst", C.c_void_p), ("ksize", C.c_int), ("cout", C.c_int),
                ("transpose_flip", C.c_int), ("pad_", C.c_int)]


class WgradDesc(C.Structure):
    _fields_ = [
        ("x_bf16", C.c_void_p), ("g_bf16", C.c_void_p),
        ("NB", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("ksize", C.c_int), ("scale", C.c_float), ("accumulate", C.c_int),
        ("dw_hwio", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class WgradPlanInfo(C.Structure):
    _fields_ = [
        ("flops", C.c_double),
        ("grid", C.c_int), ("smem_bytes", C.c_int), ("seg_width", C.c_int), ("nseg", C.c_int),
        ("ring_rows", C.c_int), ("g_slots", C.c_int), ("tap_groups", C.c_int), ("rows_per_unit", C.c_int),
        ("images_per_row", C.c_int),
    ]


class ScoreResult(C.Structure):
    _fields_ = [
        ("sum_sq_y", C.c_double),
        ("ssim_y_sum", C.c_double),
        ("ssim_rgb_sum", C.c_double * 3),
        ("n_pix", C.c_int64),
        ("n_win", C.c_int64),
        ("acc", C.c_uint64 * 6),
        ("ticket", C.c_uint64),
    ]


class ScoreItem(C.Structure):
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("h", C.c_int), ("w", C.c_int)]


class ModelConfig(C.Structure):
    _fields_ = [("precision", C.c_int), ("stream_lr_fp32", C.c_int), ("stream_hr_fp32", C.c_int),
                ("a_mode", C.c_int), ("nacc", C.c_int), ("pair", C.c_int), ("use_graphs", C.c_int),
                ("overlap_heads", C.c_int), ("fused_colsum", C.c_int), ("chain_lr", C.c_int),
                ("overlap_train", C.c_int)]


class ForwardDesc(C.Structure):
    _fields_ = [
        ("NB", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("x", C.c_void_p), ("out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("n_groups", C.c_int),
        ("group_eh", C.POINTER(C.c_int)), ("group_ew", C.POINTER(C.c_int)), ("group_n", C.POINTER(C.c_int)),
        ("group_index", C.POINTER(C.c_int)),
        ("stitch_tiles", C.c_void_p), ("stitch_u8", C.c_void_p), ("stitch_mul", C.c_float),
    ]


class TrainDesc(C.Structure):
    _fields_ = [
        ("NB", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("x", C.c_void_p), ("y", C.c_void_p), ("grads", C.c_void_p), ("loss_sum", C.c_void_p),
        ("pred", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class ModelRunInfo(C.Structure):
    _fields_ = [("conv_flops", C.c_double), ("launches", C.c_int), ("conv_launches", C.c_int),
                ("graph_replay", C.c_int)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol declared in include/sr100.h
SIGNATURES = {
    "sr_last_error_string": (C.c_char_p, []),
    "sr_version": (_i, []),
    "sr_device_supported": (_i, []),
    "sr_dev_switches": (_i, []),
    "sr_dev_set_timeline": (_i, [_vp]),
    "sr_abi_struct_size": (_sz, [_i]),
    "sr_conv_plan_create": (_i, [C.POINTER(ConvDesc), C.POINTER(_vp)]),
    "sr_conv_plan_run": (_i, [_vp, _vp]),
    "sr_conv_plan_destroy": (None, [_vp]),
    "sr_conv_plan_info": (_i, [_vp, C.POINTER(ConvPlanInfo)]),
    "sr_conv_chain_create": (_i, [C.POINTER(ConvDesc), C.POINTER(_i), _i, C.POINTER(_vp)]),
    "sr_conv_chain_run": (_i, [_vp, _vp]),
    "sr_conv_chain_destroy": (None, [_vp]),
    "sr_conv_chain_info": (_i, [_vp, C.POINTER(ConvPlanInfo)]),
    "sr_packed_weight_bytes": (_sz, [_i, _i]),
    "sr_pack_conv_weights": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "sr_pack_conv_weights_batched": (_i, [_vp, _vp, _i, _sz, _vp]),
    "sr_packed_weight_bytes_tf32": (_sz, [_i, _i]),
    "sr_pack_conv_weights_tf32": (_i, [_vp, _i, _i, _vp, _vp]),
    "sr_round_tf32": (_i, [_vp, _sz, _vp, _vp]),
    "sr_conv2d_direct": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "sr_head1x1_fwd": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "sr_bilinear4_fwd": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sr_bilinear2_fwd": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sr_bilinear4_crop_fwd": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sr_bilinear4_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sr_patch_count": (_i, [_i, _i, _i]),
    "sr_canvas_size": (_i, [_i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "sr_patch_gather_u8": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "sr_patch_gather_u8_batched": (_i, [_vp, _i, _sz, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "sr_batch_gather_u8": (_i, [_vp, _sz, _sz, _vp, _i, _f, _vp, _vp]),
    "sr_patch_down4_u8": (_i, [_vp, _i, _i, _i, _i, _i, _i, C.c_longlong, C.c_longlong, _i, _vp, _vp, _i, _f, _vp, _vp]),
    "sr_patch_average_accumulate": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp, _vp, _vp]),
    "sr_patch_average_finalize": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sr_patch_gather_f32": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "sr_patch_stitch": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "sr_patch_stitch_range": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp]),
    "sr_resize_u8": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "sr_sharpen3x3_u8": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "sr_dataprep_patches": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "sr_depth_to_space": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "sr_rgb2y_u8": (_i, [_vp, _sz, _vp, _vp]),
    "sr_score_pair_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "sr_score_batch_u8": (_i, [_vp, _i, _i, _vp, _vp]),
    "sr_sum_sq_diff_f64": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "sr_mse_loss_grad": (_i, [_vp, _vp, _sz, _sz, _vp, _vp, _vp]),
    "sr_adam_step": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _i, _f, _vp]),
    "sr_wgrad_workspace_bytes": (_sz, []),
    "sr_wgrad_plan_create": (_i, [C.POINTER(WgradDesc), C.POINTER(_vp)]),
    "sr_wgrad_plan_run": (_i, [_vp, _vp]),
    "sr_wgrad_plan_destroy": (None, [_vp]),
    "sr_wgrad_plan_info": (_i, [_vp, C.POINTER(WgradPlanInfo)]),
    "sr_mse_tail_grad": (_i, [_vp, _vp, _sz, _i, _sz, _vp, _vp, _vp]),
    "sr_mse_tail_grad_col": (_i, [_vp, _vp, _i, _i, _i, _sz, _vp, _vp, _vp, _vp]),
    "sr_colsum_bf16": (_i, [_vp, _sz, _f, _vp, _vp]),
    "sr_head1x1_bwd": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "sr_axpby_f32": (_i, [_vp, _vp, _f, _f, _sz, _vp, _vp, _vp]),
    "sr_cast_f32_to_bf16": (_i, [_vp, _sz, _vp, _vp]),
    "sr_cast_bf16_to_f32": (_i, [_vp, _sz, _vp, _vp]),
    "sr_model_default_config": (None, [C.POINTER(ModelConfig)]),
    "sr_model_num_layers": (_i, []),
    "sr_model_param_count": (_sz, []),
    "sr_model_layer": (_i, [_i, C.c_char_p, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz),
                            C.POINTER(_sz)]),
    "sr_model_create": (_i, [_vp, C.POINTER(ModelConfig), C.POINTER(_vp)]),
    "sr_model_destroy": (None, [_vp]),
    "sr_model_refresh": (_i, [_vp, _vp]),
    "sr_model_forward_workspace_bytes": (_sz, [_vp, C.POINTER(ForwardDesc)]),
    "sr_model_forward": (_i, [_vp, C.POINTER(ForwardDesc), _vp]),
    "sr_model_forward_info": (_i, [_vp, C.POINTER(ForwardDesc), C.POINTER(ModelRunInfo)]),
    "sr_model_forward_timed": (_i, [_vp, C.POINTER(ForwardDesc), _vp, C.POINTER(C.c_float), C.POINTER(C.c_double),
                                    _i, C.POINTER(_i)]),
    "sr_model_train_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "sr_model_forward_backward": (_i, [_vp, C.POINTER(TrainDesc), _vp]),
    "sr_model_train_info": (_i, [_vp, C.POINTER(TrainDesc), C.POINTER(ModelRunInfo)]),
    "sr_model_apply_gradients": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _f, _f, _f, _vp]),
    "sr_model_train_step": (_i, [_vp, C.POINTER(TrainDesc), _vp, _vp, _i, _f, _f, _f, _f, _vp]),
    "sr_set_pdl": (_i, [_i]),
    "sr_bilinear2_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sr_ipc_export": (_i, [_vp, C.c_char_p, C.POINTER(_sz)]),
    "sr_ipc_open": (_i, [C.c_char_p, _sz, C.POINTER(_vp)]),
    "sr_ipc_close": (_i, [_vp, _sz]),
    "sr_exchange_signal_bytes": (_sz, []),
    "sr_exchange_shard": (_i, [_sz, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "sr_exchange_create": (_i, [_i, _i, _sz, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "sr_exchange_destroy": (None, [_vp]),
    "sr_exchange_set_timeout_ms": (_i, [_vp, C.c_double]),
    "sr_exchange_adam_step": (_i, [_vp, _vp, _vp, _i, _f, _f, _f, _f, _f, _i, _vp]),
    "sr_exchange_status": (_i, [_vp, _vp, C.POINTER(_i)]),
    "sr_peer_barrier_create": (_i, [_i, _i, C.POINTER(_vp), C.POINTER(_vp)]),
    "sr_peer_barrier_destroy": (None, [_vp]),
    "sr_peer_barrier_set_timeout_ms": (_i, [_vp, C.c_double]),
    "sr_peer_barrier_arrive_wait": (_i, [_vp, _vp]),
    "sr_peer_barrier_status": (_i, [_vp, _vp, C.POINTER(_i)]),
    "sr_model_apply_gradients_exchange": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _f, _f, _f, _vp]),
}

_lib = None


def load():
    """Load libsr100.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SrError(-3, "libsr100.so not built (%s); run `python __graft_entry__.py` or "
                          "`make -C image-enhance-keras_b200/csrc`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for which, struct in enumerate((ConvDesc, ConvPlanInfo, PackItem, WgradDesc, WgradPlanInfo, ScoreResult, ModelConfig,
                                    ForwardDesc, TrainDesc, ModelRunInfo, StitchTile, ScoreItem)):
        if lib.sr_abi_struct_size(which) != C.sizeof(struct):
            raise SrError(-1, "%s: ctypes layout (%d bytes) does not match libsr100.so (%d bytes); rebuild the "
                              "library or update sr100/_lib.py" % (struct.__name__, C.sizeof(struct),
                                                                   lib.sr_abi_struct_size(which)))
    if os.environ.get("SR100_PDL", "1") == "0":      # A/B switch: plain stream-ordered launches of the conv kernels
        lib.sr_set_pdl(0)
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SrError(rc, load().sr_last_error_string().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_device():
    """Fail loudly unless a CUDA device of compute capability 10.x is current."""
    import torch
    if not torch.cuda.is_available():
        raise SrError(-3, "no CUDA device: the sr100 engine has no CPU path")
    lib = load()
    if not lib.sr_device_supported():
        raise SrError(-2, "device is not sm_100 (B200); the sr100 kernels are sm_100a-only")
    return lib
