"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tensorium_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"TNN_API\s+[\w\s\*]+?\b(tnn_\w+)\s*\(", src)))


def test_header_declares_the_tnncuda_surface():
    syms = declared_symbols()
    # one entry point per public TNNCuda<T> method of source/nncuda.pas:100-158
    for m in ["device_count", "create", "destroy", "create_device_buffer", "free_device_buffer", "read_buffer",
              "write_buffer", "finish", "activate_array", "activate_array_swish", "derive_array", "forward_bias",
              "backward_bias", "gemm", "gemm_batched", "gemm_strided_batched", "addvv", "subvv", "mulvv", "fmavv",
              "axpy", "power", "scale", "fill", "copy", "cross_entropy_logistic", "cross_entropy_softmax",
              "softmax_batch", "forward_maxpool", "backward_maxpool", "im2col", "col2im", "upsample", "fmavss",
              "means_and_vars", "means", "variances", "normalize", "means_and_vars_delta", "normalize_delta",
              "add_dots", "forward_scale", "forward_scale_add", "forward_dropout", "backward_dropout", "cost_l2",
              "clamp", "inverse_sqrt"]:
        assert "tnn_" + m in syms, m
    for m in ["conv2d_forward", "conv2d_forward_residual", "last_conv_kernel", "set_weights_static", "event_synchronize", "host_alloc", "host_free", "conv2d_backward_filter", "conv2d_backward_data", "sgd_update",
              "batchnorm_forward_training", "batchnorm_forward_inference", "batchnorm_backward"]:
        assert "tnn_" + m in syms, m


def test_library_exports_every_declared_symbol():
    from tensorium_b200 import _lib
    lib = _lib.load()  # raises if the .so is missing: there is no CPU fallback
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared_symbols())


def test_version_and_error_string_without_gpu():
    from tensorium_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.tnn_version()
    # null-context calls fail cleanly with a message (error convention of the header)
    rc = lib.tnn_finish(None)
    assert rc != 0 and b"null context" in lib.tnn_last_error()


def test_activation_enum_matches_reference_ordinals():
    from tensorium_b200 import nncuda as nn
    # source/ntypes.pas:66-71
    assert (nn.acLOGISTIC, nn.acRELU, nn.acLINEAR, nn.acLEAKY, nn.acSWISH) == (0, 1, 4, 9, 18)
    src = open(HEADER).read()
    assert "TNN_ACT_LEAKY = 9" in src and "TNN_ACT_LINEAR = 4" in src and "TNN_ACT_LOGISTIC = 0" in src
    # every ordinal of TActivationType (source/ntypes.pas:66-71), in the reference's order
    import re
    names = ["LOGISTIC", "RELU", "RELU6", "RELIE", "LINEAR", "RAMP", "TANH", "PLSE", "REVLEAKY", "LEAKY", "ELU", "LOGGY",
             "STAIR", "HARDTAN", "LHTAN", "SELU", "SOFTMAX", "GELU", "SWISH", "MISH", "HARD_MISH", "NORM_CHAN",
             "NORM_CHAN_SOFTMAX", "NORM_CHAN_SOFTMAX_MAXVAL"]
    enum = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"TNN_ACT_(\w+) = (\d+)", src))
    for i, n in enumerate(names):
        assert enum.get(n) == i, (n, enum.get(n), i)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tensorium_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "tensorium_oracle" not in txt, fn
