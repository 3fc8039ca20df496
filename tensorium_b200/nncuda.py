"""TNNCuda — host-side mirror of the reference's `TNNCuda<T>` class (source/nncuda.pas:35-211).

Same method names, argument order and meaning as the Pascal methods, so that code (and tests) written
against the reference's `cuda.<op>(…)` calls reads the same here.  Every method forwards to the C ABI in
libtensorium_b200.so (include/tensorium_b200.h); failures raise CudaError carrying tnn_last_error()
— the reference's SAFE_CALL asserts with the CUDA error string (nncuda.pas:237-274).

Device memory arguments (`TCUMem`) are raw device addresses (Python ints), e.g. from createDeviceBuffer or
`torch.Tensor.data_ptr()`.  Offsets and sizes are in elements, as in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ConvDesc

# TActivationType ordinals (source/ntypes.pas:66-71)
acLOGISTIC, acRELU, acRELU6, acRELIE, acLINEAR, acRAMP, acTANH, acPLSE, acREVLEAKY, acLEAKY, acELU, acLOGGY, \
    acSTAIR, acHARDTAN, acLHTAN, acSELU, acSOFTMAX, acGELU, acSWISH, acMISH, acHARD_MISH, acNORM_CHAN, \
    acNORM_CHAN_SOFTMAX, acNORM_CHAN_SOFTMAX_MAXVAL = range(24)

MATH_FP32, MATH_3XTF32, MATH_TF32, MATH_BF16X3 = 0, 1, 2, 3


class CudaError(RuntimeError):
    pass


def _addr(x) -> int:
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if hasattr(x, "ptr"):
        return int(x.ptr)
    raise TypeError(f"not a device pointer: {type(x)}")


class TNNCuda:
    """One instance per device, like the reference's global `cuda` (ntensors.pas:1053-1056, 6191-6209)."""

    def __init__(self, deviceIndex: int = 0):
        self._l = _lib.load()
        h = C.c_void_p()
        rc = self._l.tnn_create(deviceIndex, C.byref(h))
        if rc != 0:
            raise CudaError(self._l.tnn_last_error().decode())
        self._h = h
        self.deviceIndex = deviceIndex
        self.useBLAS = 1  # accepted for source compatibility (CIFAR10Conv.pas:139); see set_math_mode

    # ---- plumbing
    def _ck(self, rc: int):
        if rc != 0:
            raise CudaError(self._l.tnn_last_error().decode())

    @staticmethod
    def deviceCount() -> int:
        n = C.c_int32(0)
        lib = _lib.load()
        if lib.tnn_device_count(C.byref(n)) != 0:
            raise CudaError(lib.tnn_last_error().decode())
        return n.value

    def Destroy(self):
        if self._h:
            self._l.tnn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.Destroy()
        except Exception:
            pass

    @property
    def name(self) -> str:  # properties.name
        return self._l.tnn_device_name(self._h).decode()

    def device_info(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self._l.tnn_device_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(sm_count=a.value, cc=(b.value, c.value))

    def set_stream(self, stream_handle: int):
        self._ck(self._l.tnn_set_stream(self._h, C.c_void_p(stream_handle)))

    def get_stream(self) -> int:
        return int(self._l.tnn_get_stream(self._h) or 0)

    def set_math_mode(self, mode: int):
        self._ck(self._l.tnn_set_math_mode(self._h, mode))

    def get_math_mode(self) -> int:
        return self._l.tnn_get_math_mode(self._h)

    def launch_count(self) -> int:
        return int(self._l.tnn_launch_count(self._h))

    # ---- memory (nncuda.pas:488-522)
    def createDeviceBuffer(self, N: int) -> int:
        p = C.c_void_p()
        self._ck(self._l.tnn_create_device_buffer(self._h, N, C.byref(p)))
        return int(p.value or 0)

    def freeDeviceBuffer(self, cudaMem):
        self._ck(self._l.tnn_free_device_buffer(self._h, _addr(cudaMem)))

    def readBuffer(self, cudaMem, bufferSize: int, buffer: np.ndarray):
        assert buffer.flags["C_CONTIGUOUS"] and buffer.nbytes >= bufferSize
        self._ck(self._l.tnn_read_buffer(self._h, _addr(cudaMem), bufferSize, buffer.ctypes.data))

    def writeBuffer(self, cudaMem, bufferSize: int, buffer: np.ndarray):
        assert buffer.flags["C_CONTIGUOUS"] and buffer.nbytes >= bufferSize
        self._ck(self._l.tnn_write_buffer(self._h, _addr(cudaMem), bufferSize, buffer.ctypes.data))

    def readBufferAsync(self, cudaMem, bufferSize: int, host_ptr: int):
        self._ck(self._l.tnn_read_buffer_async(self._h, _addr(cudaMem), bufferSize, host_ptr))

    def writeBufferAsync(self, cudaMem, bufferSize: int, host_ptr: int):
        self._ck(self._l.tnn_write_buffer_async(self._h, _addr(cudaMem), bufferSize, host_ptr))

    def finish(self):
        self._ck(self._l.tnn_finish(self._h))

    # ---- events (stream-to-stream ordering between two TNNCuda instances of one device)
    def createEvent(self) -> int:
        e = C.c_void_p()
        self._ck(self._l.tnn_event_create(self._h, C.byref(e)))
        return int(e.value)

    def recordEvent(self, ev: int):
        self._ck(self._l.tnn_event_record(self._h, C.c_void_p(ev)))

    def waitEvent(self, ev: int):
        self._ck(self._l.tnn_stream_wait_event(self._h, C.c_void_p(ev)))

    # ---- CUDA graphs (launch-bound inner loops)
    def beginCapture(self):
        self._ck(self._l.tnn_graph_begin_capture(self._h))

    def endCapture(self) -> int:
        g = C.c_void_p()
        self._ck(self._l.tnn_graph_end_capture(self._h, C.byref(g)))
        return g.value

    def launchGraph(self, graph: int):
        self._ck(self._l.tnn_graph_launch(self._h, graph))

    def destroyGraph(self, graph: int):
        self._l.tnn_graph_destroy(graph)

    def counterAdd(self, ctr, inc: int):
        self._ck(self._l.tnn_counter_add(self._h, _addr(ctr), inc & 0xFFFFFFFFFFFFFFFF))

    def forwardDropoutDev(self, N, src, probability, scale, rnd, dst, seed_dev, seed_offset):
        self._ck(self._l.tnn_forward_dropout_dev(self._h, N, _addr(src), probability, scale, _addr(rnd), _addr(dst),
                                                 _addr(seed_dev), seed_offset & 0xFFFFFFFFFFFFFFFF))

    def syncEvent(self, ev: int):
        self._ck(self._l.tnn_event_synchronize(ev))

    def allocHost(self, shape, dtype=np.float32) -> np.ndarray:
        """Page-locked host array (for writeBufferAsync / readBufferAsync); release with freeHost."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._ck(self._l.tnn_host_alloc(self._h, n, C.byref(p)))
        buf = (C.c_char * n).from_address(p.value)
        a = np.frombuffer(buf, dtype=dtype).reshape(shape)
        a.flags.writeable = True
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[a.ctypes.data] = p.value
        return a

    def freeHost(self, a: np.ndarray):
        p = getattr(self, "_pinned", {}).pop(a.ctypes.data, None)
        if p:
            self._ck(self._l.tnn_host_free(self._h, p))

    def destroyEvent(self, ev: int):
        self._l.tnn_event_destroy(C.c_void_p(ev))

    # ---- activations (nncuda.pas:524-580)
    def ActivateArray(self, N, x, offset, activation):
        self._ck(self._l.tnn_activate_array(self._h, N, _addr(x), offset, int(activation)))

    def activateArraySWISH(self, N, x, offset, output_sigmoid, output):
        self._ck(self._l.tnn_activate_array_swish(self._h, N, _addr(x), offset, _addr(output_sigmoid), _addr(output)))

    def DeriveArray(self, N, x, offset, activation, delta):
        self._ck(self._l.tnn_derive_array(self._h, N, _addr(x), offset, int(activation), _addr(delta)))

    # ---- bias / scale (nncuda.pas:582-622, 1468-1510)
    def forwardBias(self, dstSize, dst, offset, srcSize, src, incb, batch):
        self._ck(self._l.tnn_forward_bias(self._h, dstSize, _addr(dst), offset, srcSize, _addr(src), incb, batch))

    def backwardBias(self, dstSize, dst, srcSize, src, srcOffset, incb, batch):
        self._ck(self._l.tnn_backward_bias(self._h, dstSize, _addr(dst), srcSize, _addr(src), srcOffset, incb, batch))

    def forwardScale(self, dstSize, dst, offset, scaleSize, scale, incb, batch):
        self._ck(self._l.tnn_forward_scale(self._h, dstSize, _addr(dst), offset, scaleSize, _addr(scale), incb, batch))

    def forwardScaleAdd(self, dstSize, dst, offset, scaleSize, scales, biases, incb, batch):
        self._ck(self._l.tnn_forward_scale_add(self._h, dstSize, _addr(dst), offset, scaleSize, _addr(scales),
                                               _addr(biases), incb, batch))

    # ---- SGEMM (nncuda.pas:624-895)
    def gemm(self, transA, transB, M, N, K, ALPHA, A, aOffset, lda, B, bOffset, ldb, BETA, Cm, cOffset, ldc):
        self._ck(self._l.tnn_gemm(self._h, int(bool(transA)), int(bool(transB)), M, N, K, ALPHA, _addr(A), aOffset,
                                  lda, _addr(B), bOffset, ldb, BETA, _addr(Cm), cOffset, ldc))

    def gemmBatched(self, transA, transB, M, N, K, ALPHA, A, aOffset, lda, B, bOffset, ldb, BETA, Cm, cOffset, ldc,
                    batchCount):
        arr = lambda ps: (C.c_void_p * batchCount)(*[_addr(p) for p in ps])
        self._ck(self._l.tnn_gemm_batched(self._h, int(bool(transA)), int(bool(transB)), M, N, K, ALPHA, arr(A),
                                          aOffset, lda, arr(B), bOffset, ldb, BETA, arr(Cm), cOffset, ldc,
                                          batchCount))

    def gemmStridedBatched(self, transA, transB, M, N, K, ALPHA, A, aOffset, lda, strideA, B, bOffset, ldb, strideB,
                           BETA, Cm, cOffset, ldc, strideC, batchCount):
        self._ck(self._l.tnn_gemm_strided_batched(self._h, int(bool(transA)), int(bool(transB)), M, N, K, ALPHA,
                                                  _addr(A), aOffset, lda, strideA, _addr(B), bOffset, ldb, strideB,
                                                  BETA, _addr(Cm), cOffset, ldc, strideC, batchCount))

    # ---- BLAS-1 style (nncuda.pas:897-1054)
    def addvv(self, N, src1, src1Offset, inca, src2, src2Offset, incb, dst, dstOffset, incc):
        self._ck(self._l.tnn_addvv(self._h, N, _addr(src1), src1Offset, inca, _addr(src2), src2Offset, incb,
                                   _addr(dst), dstOffset, incc))

    def subvv(self, N, src1, src1Offset, inca, src2, src2Offset, incb, dst, dstOffset, incc):
        self._ck(self._l.tnn_subvv(self._h, N, _addr(src1), src1Offset, inca, _addr(src2), src2Offset, incb,
                                   _addr(dst), dstOffset, incc))

    def mulvv(self, N, src1, src1Offset, inca, src2, src2Offset, incb, dst, dstOffset, incc):
        self._ck(self._l.tnn_mulvv(self._h, N, _addr(src1), src1Offset, inca, _addr(src2), src2Offset, incb,
                                   _addr(dst), dstOffset, incc))

    def fmavv(self, N, src1, src1Offset, inca, src2, src2Offset, incb, src3, src3Offset, incc, dst, dstOffset, incd):
        self._ck(self._l.tnn_fmavv(self._h, N, _addr(src1), src1Offset, inca, _addr(src2), src2Offset, incb,
                                   _addr(src3), src3Offset, incc, _addr(dst), dstOffset, incd))

    def axpy(self, N, a, x, xOffset, incx, y, yOffset, incy):
        self._ck(self._l.tnn_axpy(self._h, N, a, _addr(x), xOffset, incx, _addr(y), yOffset, incy))

    def power(self, N, x, xOffset, incx, a, y, yOffset, incy):
        self._ck(self._l.tnn_power(self._h, N, _addr(x), xOffset, incx, a, _addr(y), yOffset, incy))

    def scale(self, N, a, x, stride):
        self._ck(self._l.tnn_scale(self._h, N, a, _addr(x), stride))

    def fill(self, N, x, offset, val, stride):
        self._ck(self._l.tnn_fill(self._h, N, _addr(x), offset, val, stride))

    def copy(self, N, src, srcOffset, inca, dst, dstOffset, incb):
        self._ck(self._l.tnn_copy(self._h, N, _addr(src), srcOffset, inca, _addr(dst), dstOffset, incb))

    def lstmGates(self, batch, hidden, wz, bw, uz, uzOffset, c, h, cell_out, h_out, outOffset):
        self._ck(self._l.tnn_lstm_gates(self._h, batch, hidden, _addr(wz), _addr(bw), _addr(uz), uzOffset, _addr(c),
                                        _addr(h), _addr(cell_out), _addr(h_out), outOffset))

    def fmavss(self, N, src, offset, scalar, bias, dst):
        self._ck(self._l.tnn_fmavss(self._h, N, _addr(src), offset, scalar, bias, _addr(dst)))

    def clamp(self, N, alpha, src, dst, stride=1, offset=0):
        self._ck(self._l.tnn_clamp(self._h, N, alpha, _addr(src), _addr(dst), stride, offset))

    def inverseSqrt(self, N, alpha, src, dst, stride=1, offset=0):
        self._ck(self._l.tnn_inverse_sqrt(self._h, N, alpha, _addr(src), _addr(dst), stride, offset))

    # ---- losses / softmax / pooling / upsample / dropout
    def crossEntropyLogistic(self, N, pred, truth, delta, error):
        self._ck(self._l.tnn_cross_entropy_logistic(self._h, N, _addr(pred), _addr(truth), _addr(delta), _addr(error)))

    def crossEntropySoftmax(self, N, pred, truth, delta, error):
        self._ck(self._l.tnn_cross_entropy_softmax(self._h, N, _addr(pred), _addr(truth), _addr(delta), _addr(error)))

    def costL2(self, N, pred, truth, delta, error):
        self._ck(self._l.tnn_cost_l2(self._h, N, _addr(pred), _addr(truth), _addr(delta), _addr(error)))

    def softmaxBatch(self, N, input, iOffset, batch, batch_size, groups, group_size, stride, temp, output, oOffset):
        self._ck(self._l.tnn_softmax_batch(self._h, N, _addr(input), iOffset, batch, batch_size, groups, group_size,
                                           stride, temp, _addr(output), oOffset))

    def forwardMaxPool(self, aBatch, outC, outH, outW, input, c, h, w, stride_x, stride_y, padding, kernelSize,
                       indexes, output):
        self._ck(self._l.tnn_forward_maxpool(self._h, aBatch, outC, outH, outW, _addr(input), c, h, w, stride_x,
                                             stride_y, padding, kernelSize, _addr(indexes), _addr(output)))

    def backwardMaxPool(self, aBatch, outC, outH, outW, output, indexes, delta):
        self._ck(self._l.tnn_backward_maxpool(self._h, aBatch, outC, outH, outW, _addr(output), _addr(indexes),
                                              _addr(delta)))

    def upSample(self, aBatch, aChannels, outHeight, outWidth, in_, stride, isForward, scale, out, zeroIn=0):
        self._ck(self._l.tnn_upsample(self._h, aBatch, aChannels, outHeight, outWidth, _addr(in_), stride,
                                      int(isForward), scale, _addr(out), int(zeroIn)))

    def forwardDropout(self, N, src, probability, scale, rnd, dst, seed=0):
        self._ck(self._l.tnn_forward_dropout(self._h, N, _addr(src), probability, scale, _addr(rnd), _addr(dst), seed))

    def backwardDropout(self, N, src, probability, scale, rnd, dst):
        self._ck(self._l.tnn_backward_dropout(self._h, N, _addr(src), probability, scale, _addr(rnd), _addr(dst)))

    # ---- im2col / col2im (nncuda.pas:1144-1265)
    def im2col(self, aChannels, aHeight, aWidth, kernelHeight, kernelWidth, padHeight, padWidth, strideY, strideX,
               dilationY, dilationX, im, imOffset, col, colOffset):
        self._ck(self._l.tnn_im2col(self._h, aChannels, aHeight, aWidth, kernelHeight, kernelWidth, padHeight,
                                    padWidth, strideY, strideX, dilationY, dilationX, _addr(im), imOffset, _addr(col),
                                    colOffset))

    def col2im(self, aChannels, aHeight, aWidth, kernelHeight, kernelWidth, padHeight, padWidth, strideY, strideX,
               dilationY, dilationX, col, colOffset, im, imOffset):
        self._ck(self._l.tnn_col2im(self._h, aChannels, aHeight, aWidth, kernelHeight, kernelWidth, padHeight,
                                    padWidth, strideY, strideX, dilationY, dilationX, _addr(col), colOffset, _addr(im),
                                    imOffset))

    # ---- batch-norm pieces (nncuda.pas:1302-1466)
    def means(self, srcSize, dstSize, groups, src, offset, means):
        self._ck(self._l.tnn_means(self._h, srcSize, dstSize, groups, _addr(src), offset, _addr(means)))

    def variances(self, srcSize, dstSize, groups, src, offset, means, vars):
        self._ck(self._l.tnn_variances(self._h, srcSize, dstSize, groups, _addr(src), offset, _addr(means),
                                       _addr(vars)))

    def meansAndVars(self, srcSize, dstSize, groups, src, offset, means, vars):
        self._ck(self._l.tnn_means_and_vars(self._h, srcSize, dstSize, groups, _addr(src), offset, _addr(means),
                                            _addr(vars)))

    def normalize(self, srcSize, dstSize, groups, means, meansStride, vars, varsStride, dst, dstOffset):
        self._ck(self._l.tnn_normalize(self._h, srcSize, dstSize, groups, _addr(means), meansStride, _addr(vars),
                                       varsStride, _addr(dst), dstOffset))

    def meansAndVarsDelta(self, srcSize, dstSize, groups, delta, x, offset, mean, variance, mean_delta,
                          variance_delta):
        self._ck(self._l.tnn_means_and_vars_delta(self._h, srcSize, dstSize, groups, _addr(delta), _addr(x), offset,
                                                  _addr(mean), _addr(variance), _addr(mean_delta),
                                                  _addr(variance_delta)))

    def normalizeDelta(self, deltaSize, meanSize, groups, delta, x, offset, mean, variance, mean_delta,
                       variance_delta):
        self._ck(self._l.tnn_normalize_delta(self._h, deltaSize, meanSize, groups, _addr(delta), _addr(x), offset,
                                             _addr(mean), _addr(variance), _addr(mean_delta), _addr(variance_delta)))

    def addDots(self, N, dstSize, groups, src1, src2, srcOffset, dst):
        self._ck(self._l.tnn_add_dots(self._h, N, dstSize, groups, _addr(src1), _addr(src2), srcOffset, _addr(dst)))

    # ---- coarse fused entry points (new with this backend; see include/tensorium_b200.h)
    @staticmethod
    def conv_desc(batch, c, h, w, filters, kh, kw, pad_h, pad_w, stride_y, stride_x, dilation_y=1, dilation_x=1):
        return ConvDesc(batch, c, h, w, filters, kh, kw, pad_h, pad_w, stride_y, stride_x, dilation_y, dilation_x)

    def conv2dForward(self, desc, in_, inOffset, weights, scale, bias, activation, out, outOffset):
        self._ck(self._l.tnn_conv2d_forward(self._h, C.byref(desc), _addr(in_), inOffset, _addr(weights),
                                            _addr(scale), _addr(bias), int(activation), _addr(out), outOffset))

    CONV_KERNELS = {0: "none", 1: "direct::k_conv_direct", 2: "umma::k_umma<Im2colPixelLoader> (gather)",
                    3: "umma::k_umma<TmaPatchLoader> (TMA activations)", 4: "umma::k_umma_halo", 5: "simt::k_gemm",
                    6: "umma::k_umma_pair (gather, CTA pairs)",
                    7: "nhwc::k_conv_nhwc (TMA im2col activations, TMA store)", 8: "direct::k_conv_direct (NHWC out)",
                    9: "nhwc::k_conv_image (planar image in, tensor cores, TMA store)",
                    11: "nhwc::k_conv_nhwc_pair (CTA pairs, TMA im2col activations, TMA store)"}

    def setWeightsStatic(self, weights, is_static: bool = True):
        """The backend may keep its re-ordered / TF32-split copy of `weights` across calls (inference)."""
        self._ck(self._l.tnn_set_weights_static(self._h, _addr(weights), int(bool(is_static))))

    def lastConvKernel(self) -> str:
        return self.CONV_KERNELS.get(int(self._l.tnn_last_conv_kernel(self._h)), "?")

    def conv2dForwardResidual(self, desc, in_, inOffset, weights, scale, bias, activation, residual, residualOffset, out,
                              outOffset):
        self._ck(self._l.tnn_conv2d_forward_residual(self._h, C.byref(desc), _addr(in_), inOffset, _addr(weights),
                                                     _addr(scale), _addr(bias), int(activation), _addr(residual),
                                                     residualOffset, _addr(out), outOffset))

    def conv2dBackwardFilter(self, desc, in_, inOffset, delta, deltaOffset, weight_updates):
        self._ck(self._l.tnn_conv2d_backward_filter(self._h, C.byref(desc), _addr(in_), inOffset, _addr(delta),
                                                    deltaOffset, _addr(weight_updates)))

    def conv2dBackwardData(self, desc, weights, delta, deltaOffset, in_delta, inDeltaOffset):
        self._ck(self._l.tnn_conv2d_backward_data(self._h, C.byref(desc), _addr(weights), _addr(delta), deltaOffset,
                                                  _addr(in_delta), inDeltaOffset))

    def batchNormForwardTraining(self, batch, channels, blockSize, out, x, x_norm, mean, variance, rolling_mean,
                                 rolling_variance, bn_momentum, scales, biases, activation):
        self._ck(self._l.tnn_batchnorm_forward_training(self._h, batch, channels, blockSize, _addr(out), _addr(x),
                                                        _addr(x_norm), _addr(mean), _addr(variance),
                                                        _addr(rolling_mean), _addr(rolling_variance), bn_momentum,
                                                        _addr(scales), _addr(biases), int(activation)))

    def batchNormForwardInference(self, batch, channels, blockSize, out, rolling_mean, rolling_variance, scales,
                                  biases, activation):
        self._ck(self._l.tnn_batchnorm_forward_inference(self._h, batch, channels, blockSize, _addr(out),
                                                         _addr(rolling_mean), _addr(rolling_variance), _addr(scales),
                                                         _addr(biases), int(activation)))

    def batchNormBackward(self, batch, channels, blockSize, out, activation, delta, x, x_norm, mean, variance, scales,
                          scale_updates, mean_delta, variance_delta):
        self._ck(self._l.tnn_batchnorm_backward(self._h, batch, channels, blockSize, _addr(out), int(activation),
                                                _addr(delta), _addr(x), _addr(x_norm), _addr(mean), _addr(variance),
                                                _addr(scales), _addr(scale_updates), _addr(mean_delta),
                                                _addr(variance_delta)))

    def sgdUpdate(self, N, params, updates, rate, momentum, decay_term):
        self._ck(self._l.tnn_sgd_update(self._h, N, _addr(params), _addr(updates), rate, momentum, decay_term))

    def sgdUpdateMulti(self, items, momentum):
        """items: [(params devptr, updates devptr, n, rate, decay_term)] — one launch per 32 tensors."""
        n = len(items)
        if not n:
            return
        P = (C.c_void_p * n)(*[int(i[0]) for i in items])
        U = (C.c_void_p * n)(*[int(i[1]) for i in items])
        S = (C.c_uint64 * n)(*[int(i[2]) for i in items])
        R = (C.c_float * n)(*[float(i[3]) for i in items])
        D = (C.c_float * n)(*[float(i[4]) for i in items])
        self._ck(self._l.tnn_sgd_update_multi(self._h, n, P, U, S, R, D, float(momentum)))

    def yoloActivate(self, batch, n, classes, wh, out, scale_xy=1.0):
        self._ck(self._l.tnn_yolo_activate(self._h, batch, n, classes, wh, _addr(out), scale_xy))

    def yoloCandidates(self, batch, n, classes, wh, out, thresh, tag, maxCand, cand):
        """The `objectness > thresh` test of TYoloLayer.getDetections on the device: appends compact records to `cand`."""
        self._ck(self._l.tnn_yolo_candidates(self._h, batch, n, classes, wh, _addr(out), float(thresh), int(tag), maxCand,
                                             _addr(cand)))

    def concatChannels(self, batch, c, blockSize, in_, outChannels, cOffset, out):
        self._ck(self._l.tnn_concat_channels(self._h, batch, c, blockSize, _addr(in_), outChannels, cOffset,
                                             _addr(out)))

    # ---- channels-innermost (NHWC) inference graph (include/tensorium_b200.h, last section)
    LAYOUT_NCHW, LAYOUT_NHWC, LAYOUT_NHWC_SPLIT = 0, 1, 2

    def conv2dForwardNhwc(self, desc, in_, inOffset, inPitch, inLayout, weights, scale, bias, activation, out,
                          outOffset, outPitch, outLayout):
        self._ck(self._l.tnn_conv2d_forward_nhwc(self._h, C.byref(desc), _addr(in_), inOffset, inPitch, int(inLayout),
                                                 _addr(weights), _addr(scale), _addr(bias), int(activation),
                                                 _addr(out), outOffset, outPitch, int(outLayout)))

    def conv2dForwardNhwcResidual(self, desc, in_, inOffset, inPitch, inLayout, weights, scale, bias, activation, residual,
                                  residualOffset, residualPitch, out, outOffset, outPitch, outLayout):
        self._ck(self._l.tnn_conv2d_forward_nhwc_residual(
            self._h, C.byref(desc), _addr(in_), inOffset, inPitch, int(inLayout), _addr(weights), _addr(scale),
            _addr(bias), int(activation), _addr(residual), residualOffset, residualPitch, _addr(out), outOffset, outPitch,
            int(outLayout)))

    def conv2dForwardSplit(self, desc, in_, inOffset, inPitch, inLayout, inPlane, weights, scale, bias, activation, residual,
                           residualOffset, residualPitch, out, outOffset, outPitch, outLayout, outSplit, outSplitOffset,
                           outSplitPitch, outSplitPlane):
        """Pre-split (bf16 hi/lo planes) activations in and / or out — BF16x3 mode (include/tensorium_b200.h)."""
        self._ck(self._l.tnn_conv2d_forward_split(
            self._h, C.byref(desc), _addr(in_), inOffset, inPitch, int(inLayout), inPlane, _addr(weights), _addr(scale),
            _addr(bias), int(activation), _addr(residual), residualOffset, residualPitch, _addr(out), outOffset, outPitch,
            int(outLayout), _addr(outSplit), outSplitOffset, outSplitPitch, outSplitPlane))

    def conv2dForwardSplitResidual(self, desc, in_, inOffset, inPitch, inLayout, inPlane, weights, scale, bias, activation,
                                   residual, residualOffset, residualPitch, residualLayout, residualPlane, out, outOffset,
                                   outPitch, outLayout, outSplit, outSplitOffset, outSplitPitch, outSplitPlane):
        """conv2dForwardSplit with the folded shortcut's other operand as fp32 rows (LAYOUT_NHWC) or as pre-split planes
        (LAYOUT_NHWC_SPLIT: residual = hi + lo)."""
        self._ck(self._l.tnn_conv2d_forward_split_residual(
            self._h, C.byref(desc), _addr(in_), inOffset, inPitch, int(inLayout), inPlane, _addr(weights), _addr(scale),
            _addr(bias), int(activation), _addr(residual), residualOffset, residualPitch, int(residualLayout), residualPlane,
            _addr(out), outOffset, outPitch, int(outLayout), _addr(outSplit), outSplitOffset, outSplitPitch, outSplitPlane))

    def conv2dForwardYoloHead(self, desc, in_, inOffset, inPitch, inLayout, inPlane, weights, scale, bias, out, outOffset,
                              classes, scaleXY):
        """The linear 1x1 head convolution with the following yolo layer's activation folded into its epilogue; `out` is
        the yolo layer's (planar) output."""
        self._ck(self._l.tnn_conv2d_forward_yolo_head(
            self._h, C.byref(desc), _addr(in_), inOffset, inPitch, int(inLayout), inPlane, _addr(weights), _addr(scale),
            _addr(bias), _addr(out), outOffset, classes, float(scaleXY)))

    def nhwcToSplit(self, rows, channels, in_, inPitch, outSplit, outPitch, outPlane):
        self._ck(self._l.tnn_nhwc_to_split(self._h, rows, channels, _addr(in_), inPitch, _addr(outSplit), outPitch, outPlane))

    def splitToNhwc(self, rows, channels, inSplit, inPitch, inPlane, out, outPitch):
        self._ck(self._l.tnn_split_to_nhwc(self._h, rows, channels, _addr(inSplit), inPitch, inPlane, _addr(out), outPitch))

    def nchwToNhwc(self, batch, channels, blockSize, in_, out, outPitch):
        self._ck(self._l.tnn_nchw_to_nhwc(self._h, batch, channels, blockSize, _addr(in_), _addr(out), outPitch))

    def nhwcToNchw(self, batch, channels, blockSize, in_, inPitch, out):
        self._ck(self._l.tnn_nhwc_to_nchw(self._h, batch, channels, blockSize, _addr(in_), inPitch, _addr(out)))

    def rowsCopy(self, rows, channels, in_, inPitch, out, outPitch):
        self._ck(self._l.tnn_rows_copy(self._h, rows, channels, _addr(in_), inPitch, _addr(out), outPitch))

    def rowsAdd(self, rows, channels, a, aPitch, b, bPitch, out, outPitch):
        self._ck(self._l.tnn_rows_add(self._h, rows, channels, _addr(a), aPitch, _addr(b), bPitch, _addr(out),
                                      outPitch))

    def upSampleNhwc(self, batch, channels, inHeight, inWidth, stride, scale, in_, inPitch, out, outPitch):
        self._ck(self._l.tnn_upsample_nhwc(self._h, batch, channels, inHeight, inWidth, stride, scale, _addr(in_),
                                           inPitch, _addr(out), outPitch))

    # ---- multi-GPU exchange (include/tensorium_b200.h, last section): NCCL behind the C ABI
    @staticmethod
    def commUniqueId() -> bytes:
        buf = C.create_string_buffer(128)
        lib = _lib.load()
        if lib.tnn_comm_unique_id(buf, 128):
            raise CudaError(lib.tnn_last_error().decode())
        return buf.raw

    def commInitRank(self, nranks: int, rank: int, uid: bytes):
        comm = C.c_void_p()
        self._ck(self._l.tnn_comm_init_rank(self._h, nranks, rank, C.c_char_p(uid), C.byref(comm)))
        return comm

    def commDestroy(self, comm):
        self._ck(self._l.tnn_comm_destroy(comm))

    def commInfo(self, comm):
        n, r, v = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self._l.tnn_comm_info(comm, C.byref(n), C.byref(r), C.byref(v)))
        return n.value, r.value, v.value

    def allreduceSum(self, comm, buf, offset, n):
        self._ck(self._l.tnn_allreduce_sum(self._h, comm, _addr(buf), offset, n))

    def broadcast(self, comm, buf, offset, n, root=0):
        self._ck(self._l.tnn_broadcast(self._h, comm, _addr(buf), offset, n, root))

    def batchNormForwardTrainingSync(self, comm, globalBatch, batch, channels, blockSize, out, x, x_norm, mean, variance,
                                     rolling_mean, rolling_variance, bn_momentum, scales, biases, activation):
        self._ck(self._l.tnn_batchnorm_forward_training_sync(
            self._h, comm, globalBatch, batch, channels, blockSize, _addr(out), _addr(x), _addr(x_norm), _addr(mean),
            _addr(variance), _addr(rolling_mean), _addr(rolling_variance), bn_momentum, _addr(scales), _addr(biases),
            int(activation)))

    def batchNormBackwardSync(self, comm, globalBatch, batch, channels, blockSize, out, activation, delta, x, x_norm, mean,
                              variance, scales, scale_updates, mean_delta, variance_delta):
        self._ck(self._l.tnn_batchnorm_backward_sync(
            self._h, comm, globalBatch, batch, channels, blockSize, _addr(out), int(activation), _addr(delta), _addr(x),
            _addr(x_norm), _addr(mean), _addr(variance), _addr(scales), _addr(scale_updates), _addr(mean_delta),
            _addr(variance_delta)))

    def lstmGatesBackward(self, batch, hidden, wz, bw, uz, zOffset, cell, cellOffset, prevCellOffset, dh, dhOffset, dc,
                          dz, dzOffset):
        self._ck(self._l.tnn_lstm_gates_backward(self._h, batch, hidden, _addr(wz), _addr(bw), _addr(uz), zOffset,
                                                 _addr(cell), cellOffset, prevCellOffset, _addr(dh), dhOffset, _addr(dc),
                                                 _addr(dz), dzOffset))
