"""ctypes loader for libtensorium_b200.so — the C ABI declared in include/tensorium_b200.h.

There is no fallback: if the shared library is missing or a symbol cannot be bound, importing the
backend raises.  (The .so is built in-tree by tensorium_b200/build.py.)
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TNN_LIB", os.path.join(HERE, "libtensorium_b200.so"))  # TNN_LIB: A/B experiments only

i32, i64, f32, u64 = C.c_int32, C.c_int64, C.c_float, C.c_uint64
vp, cp, sz = C.c_void_p, C.c_char_p, C.c_size_t
ctxp = C.c_void_p
fptr = C.c_void_p  # raw device pointer (float*)


class ConvDesc(C.Structure):
    """tnn_conv_desc (include/tensorium_b200.h)."""
    _fields_ = [(n, i64) for n in ("batch", "c", "h", "w", "filters", "kh", "kw", "pad_h", "pad_w", "stride_y",
                                   "stride_x", "dilation_y", "dilation_x")]


# name -> (restype, [argtypes]) ; every entry point of include/tensorium_b200.h
SIGNATURES = {
    "tnn_device_count": (i32, [C.POINTER(i32)]),
    "tnn_create": (i32, [i32, C.POINTER(ctxp)]),
    "tnn_destroy": (i32, [ctxp]),
    "tnn_last_error": (cp, []),
    "tnn_device_name": (cp, [ctxp]),
    "tnn_device_info": (i32, [ctxp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "tnn_set_stream": (i32, [ctxp, vp]),
    "tnn_get_stream": (vp, [ctxp]),
    "tnn_set_math_mode": (i32, [ctxp, i32]),
    "tnn_get_math_mode": (i32, [ctxp]),
    "tnn_launch_count": (i64, [ctxp]),
    "tnn_version": (cp, []),
    "tnn_create_device_buffer": (i32, [ctxp, i64, C.POINTER(fptr)]),
    "tnn_free_device_buffer": (i32, [ctxp, fptr]),
    "tnn_read_buffer": (i32, [ctxp, fptr, sz, vp]),
    "tnn_write_buffer": (i32, [ctxp, fptr, sz, vp]),
    "tnn_read_buffer_async": (i32, [ctxp, fptr, sz, vp]),
    "tnn_write_buffer_async": (i32, [ctxp, fptr, sz, vp]),
    "tnn_finish": (i32, [ctxp]),
    "tnn_event_create": (i32, [ctxp, C.POINTER(vp)]),
    "tnn_event_record": (i32, [ctxp, vp]),
    "tnn_stream_wait_event": (i32, [ctxp, vp]),
    "tnn_event_destroy": (i32, [vp]),
    "tnn_event_synchronize": (i32, [vp]),
    "tnn_graph_begin_capture": (i32, [ctxp]),
    "tnn_graph_end_capture": (i32, [ctxp, C.POINTER(vp)]),
    "tnn_graph_launch": (i32, [ctxp, vp]),
    "tnn_graph_destroy": (i32, [vp]),
    "tnn_counter_add": (i32, [ctxp, vp, C.c_uint64]),
    "tnn_forward_dropout_dev": (i32, [ctxp, i64, fptr, f32, f32, fptr, fptr, vp, C.c_uint64]),
    "tnn_lstm_gates": (i32, [ctxp, i64, i64, fptr, fptr, fptr, i64, fptr, fptr, fptr, fptr, i64]),
    "tnn_lstm_gates_backward": (i32, [ctxp, i64, i64, fptr, fptr, fptr, i64, fptr, i64, i64, fptr, i64, fptr, fptr, i64]),
    "tnn_host_alloc": (i32, [ctxp, i64, C.POINTER(vp)]),
    "tnn_host_free": (i32, [ctxp, vp]),
    "tnn_activate_array": (i32, [ctxp, i64, fptr, i64, i32]),
    "tnn_activate_array_swish": (i32, [ctxp, i64, fptr, i64, fptr, fptr]),
    "tnn_derive_array": (i32, [ctxp, i64, fptr, i64, i32, fptr]),
    "tnn_forward_bias": (i32, [ctxp, i64, fptr, i64, i64, fptr, i64, i64]),
    "tnn_backward_bias": (i32, [ctxp, i64, fptr, i64, fptr, i64, i64, i64]),
    "tnn_forward_scale": (i32, [ctxp, i64, fptr, i64, i64, fptr, i64, i64]),
    "tnn_forward_scale_add": (i32, [ctxp, i64, fptr, i64, i64, fptr, fptr, i64, i64]),
    "tnn_gemm": (i32, [ctxp, i32, i32, i64, i64, i64, f32, fptr, i64, i64, fptr, i64, i64, f32, fptr, i64, i64]),
    "tnn_gemm_batched": (i32, [ctxp, i32, i32, i64, i64, i64, f32, C.POINTER(fptr), i64, i64, C.POINTER(fptr), i64,
                               i64, f32, C.POINTER(fptr), i64, i64, i64]),
    "tnn_gemm_strided_batched": (i32, [ctxp, i32, i32, i64, i64, i64, f32, fptr, i64, i64, i64, fptr, i64, i64, i64,
                                       f32, fptr, i64, i64, i64, i64]),
    "tnn_addvv": (i32, [ctxp, i64, fptr, i64, i64, fptr, i64, i64, fptr, i64, i64]),
    "tnn_subvv": (i32, [ctxp, i64, fptr, i64, i64, fptr, i64, i64, fptr, i64, i64]),
    "tnn_mulvv": (i32, [ctxp, i64, fptr, i64, i64, fptr, i64, i64, fptr, i64, i64]),
    "tnn_fmavv": (i32, [ctxp, i64, fptr, i64, i64, fptr, i64, i64, fptr, i64, i64, fptr, i64, i64]),
    "tnn_axpy": (i32, [ctxp, i64, f32, fptr, i64, i64, fptr, i64, i64]),
    "tnn_power": (i32, [ctxp, i64, fptr, i64, i64, f32, fptr, i64, i64]),
    "tnn_scale": (i32, [ctxp, i64, f32, fptr, i64]),
    "tnn_fill": (i32, [ctxp, i64, fptr, i64, f32, i64]),
    "tnn_copy": (i32, [ctxp, i64, fptr, i64, i64, fptr, i64, i64]),
    "tnn_fmavss": (i32, [ctxp, i64, fptr, i64, f32, f32, fptr]),
    "tnn_clamp": (i32, [ctxp, i64, f32, fptr, fptr, i64, i64]),
    "tnn_inverse_sqrt": (i32, [ctxp, i64, f32, fptr, fptr, i64, i64]),
    "tnn_cross_entropy_logistic": (i32, [ctxp, i64, fptr, fptr, fptr, fptr]),
    "tnn_cross_entropy_softmax": (i32, [ctxp, i64, fptr, fptr, fptr, fptr]),
    "tnn_cost_l2": (i32, [ctxp, i64, fptr, fptr, fptr, fptr]),
    "tnn_softmax_batch": (i32, [ctxp, i64, fptr, i64, i64, i64, i64, i64, i64, f32, fptr, i64]),
    "tnn_forward_maxpool": (i32, [ctxp, i64, i64, i64, i64, fptr, i64, i64, i64, i64, i64, i64, i64, vp, fptr]),
    "tnn_backward_maxpool": (i32, [ctxp, i64, i64, i64, i64, fptr, vp, fptr]),
    "tnn_upsample": (i32, [ctxp, i64, i64, i64, i64, fptr, i64, i32, f32, fptr, i32]),
    "tnn_forward_dropout": (i32, [ctxp, i64, fptr, f32, f32, fptr, fptr, u64]),
    "tnn_backward_dropout": (i32, [ctxp, i64, fptr, f32, f32, fptr, fptr]),
    "tnn_im2col": (i32, [ctxp] + [i64] * 11 + [fptr, i64, fptr, i64]),
    "tnn_col2im": (i32, [ctxp] + [i64] * 11 + [fptr, i64, fptr, i64]),
    "tnn_means": (i32, [ctxp, i64, i64, i64, fptr, i64, fptr]),
    "tnn_variances": (i32, [ctxp, i64, i64, i64, fptr, i64, fptr, fptr]),
    "tnn_means_and_vars": (i32, [ctxp, i64, i64, i64, fptr, i64, fptr, fptr]),
    "tnn_normalize": (i32, [ctxp, i64, i64, i64, fptr, i64, fptr, i64, fptr, i64]),
    "tnn_means_and_vars_delta": (i32, [ctxp, i64, i64, i64, fptr, fptr, i64, fptr, fptr, fptr, fptr]),
    "tnn_normalize_delta": (i32, [ctxp, i64, i64, i64, fptr, fptr, i64, fptr, fptr, fptr, fptr]),
    "tnn_add_dots": (i32, [ctxp, i64, i64, i64, fptr, fptr, i64, fptr]),
    "tnn_conv2d_forward": (i32, [ctxp, C.POINTER(ConvDesc), fptr, i64, fptr, fptr, fptr, i32, fptr, i64]),
    "tnn_last_conv_kernel": (i32, [ctxp]),
    "tnn_set_weights_static": (i32, [ctxp, fptr, i32]),
    "tnn_conv2d_forward_residual": (i32, [ctxp, C.POINTER(ConvDesc), fptr, i64, fptr, fptr, fptr, i32, fptr, i64, fptr,
                                          i64]),
    "tnn_conv2d_backward_filter": (i32, [ctxp, C.POINTER(ConvDesc), fptr, i64, fptr, i64, fptr]),
    "tnn_conv2d_backward_data": (i32, [ctxp, C.POINTER(ConvDesc), fptr, fptr, i64, fptr, i64]),
    "tnn_batchnorm_forward_training": (i32, [ctxp, i64, i64, i64, fptr, fptr, fptr, fptr, fptr, fptr, fptr, f32, fptr,
                                             fptr, i32]),
    "tnn_batchnorm_forward_inference": (i32, [ctxp, i64, i64, i64, fptr, fptr, fptr, fptr, fptr, i32]),
    "tnn_batchnorm_backward": (i32, [ctxp, i64, i64, i64, fptr, i32, fptr, fptr, fptr, fptr, fptr, fptr, fptr, fptr,
                                     fptr]),
    "tnn_sgd_update": (i32, [ctxp, i64, fptr, fptr, f32, f32, f32]),
    "tnn_sgd_update_multi": (i32, [ctxp, i32, vp, vp, vp, vp, vp, f32]),
    "tnn_yolo_activate": (i32, [ctxp, i64, i64, i64, i64, fptr, f32]),
    "tnn_yolo_candidates": (i32, [ctxp, i64, i64, i64, i64, fptr, f32, i32, i64, fptr]),
    "tnn_concat_channels": (i32, [ctxp, i64, i64, i64, fptr, i64, i64, fptr]),
    "tnn_conv2d_forward_nhwc": (i32, [ctxp, C.POINTER(ConvDesc), fptr, i64, i64, i32, fptr, fptr, fptr, i32, fptr, i64,
                                      i64, i32]),
    "tnn_conv2d_forward_nhwc_residual": (i32, [ctxp, C.POINTER(ConvDesc), fptr, i64, i64, i32, fptr, fptr, fptr, i32, fptr,
                                               i64, i64, fptr, i64, i64, i32]),
    "tnn_conv2d_forward_split": (i32, [ctxp, C.POINTER(ConvDesc), fptr, i64, i64, i32, i64, fptr, fptr, fptr, i32, fptr, i64,
                                       i64, fptr, i64, i64, i32, fptr, i64, i64, i64]),
    "tnn_conv2d_forward_split_residual": (i32, [ctxp, C.POINTER(ConvDesc), fptr, i64, i64, i32, i64, fptr, fptr, fptr, i32,
                                                fptr, i64, i64, i32, i64, fptr, i64, i64, i32, fptr, i64, i64, i64]),
    "tnn_conv2d_forward_yolo_head": (i32, [ctxp, C.POINTER(ConvDesc), fptr, i64, i64, i32, i64, fptr, fptr, fptr, fptr, i64,
                                           i64, f32]),
    "tnn_nhwc_to_split": (i32, [ctxp, i64, i64, fptr, i64, fptr, i64, i64]),
    "tnn_split_to_nhwc": (i32, [ctxp, i64, i64, fptr, i64, i64, fptr, i64]),
    "tnn_nchw_to_nhwc": (i32, [ctxp, i64, i64, i64, fptr, fptr, i64]),
    "tnn_nhwc_to_nchw": (i32, [ctxp, i64, i64, i64, fptr, i64, fptr]),
    "tnn_rows_copy": (i32, [ctxp, i64, i64, fptr, i64, fptr, i64]),
    "tnn_rows_add": (i32, [ctxp, i64, i64, fptr, i64, fptr, i64, fptr, i64]),
    "tnn_upsample_nhwc": (i32, [ctxp, i64, i64, i64, i64, i64, f32, fptr, i64, fptr, i64]),
    "tnn_comm_unique_id": (i32, [vp, i64]),
    "tnn_comm_init_rank": (i32, [ctxp, i32, i32, vp, C.POINTER(vp)]),
    "tnn_comm_destroy": (i32, [vp]),
    "tnn_comm_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "tnn_allreduce_sum": (i32, [ctxp, vp, fptr, i64, i64]),
    "tnn_broadcast": (i32, [ctxp, vp, fptr, i64, i64, i32]),
    "tnn_batchnorm_forward_training_sync": (i32, [ctxp, vp, i64, i64, i64, i64, fptr, fptr, fptr, fptr, fptr, fptr, fptr, f32,
                                                  fptr, fptr, i32]),
    "tnn_batchnorm_backward_sync": (i32, [ctxp, vp, i64, i64, i64, i64, fptr, i32, fptr, fptr, fptr, fptr, fptr, fptr, fptr,
                                          fptr, fptr]),
}

_lib = None


class BackendError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the backend and bind every symbol; raises if anything is missing (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BackendError(
            f"{LIB_PATH} is missing: build it with `python -m tensorium_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise BackendError(f"libtensorium_b200.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
