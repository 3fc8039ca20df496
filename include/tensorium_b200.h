/*
 * tensorium_b200.h — C ABI of libtensorium_b200.so
 *
 * A B200 (sm_100a) native replacement for the reference's CUDA backend unit
 * `nnCuda` (reference: source/nncuda.pas, class TNNCuda<T>, lines 35-211; the
 * global instance `cuda : TNNCuda<Single>`, source/ntensors.pas:1053-1056).
 * Every entry point below is what a FreePascal `external 'tensorium_b200'`
 * declaration binds; INTEGRATION.md shows the replacement unit.
 *
 * Conventions (same as the reference methods they replace):
 *   - all sizes, strides and offsets are int64_t (= Pascal SizeInt on x86-64);
 *     offsets are in ELEMENTS (floats), not bytes;
 *   - `float*` arguments are raw device pointers obtained from
 *     tnn_create_device_buffer (plain cudaMalloc memory: TTensor.Free calls
 *     cudaFree on them directly, ntensors.pas:6126-6129) or any other CUDA
 *     allocation of the same device;
 *   - tensors are fp32, NCHW, weights [filters][C*kh*kw] (nConvolutionLayer.pas:115-119);
 *   - calls are asynchronous on the context's stream (nncuda.pas:455) except the
 *     host<->device copies and tnn_finish;
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from tnn_last_error() (thread local).  The reference's SAFE_CALL
 *     (nncuda.pas:237-274) asserts on the same condition.
 *   - arithmetic follows the reference's CPU path (ntensors.pas / nactivation.pas),
 *     not its NVRTC kernels, where the two differ: variance is
 *     divided by N-1, normalisation clamps max(sqrt(var),1e-6), tanh = (e^x-e^-x)/(e^x+e^-x).
 *
 * No entry point falls back to the CPU: if no sm_100 device is present,
 * tnn_create fails.
 */
#ifndef TENSORIUM_B200_H
#define TENSORIUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TNN_API __declspec(dllexport)
#else
#define TNN_API __attribute__((visibility("default")))
#endif

typedef struct tnn_context tnn_context; /* one TNNCuda<Single> instance */

/* TActivationType ordinals — source/ntypes.pas:66-71 (must match; cf. cuda_sgemm.cu:93-98) */
typedef enum {
  TNN_ACT_LOGISTIC = 0, TNN_ACT_RELU = 1, TNN_ACT_RELU6 = 2, TNN_ACT_RELIE = 3, TNN_ACT_LINEAR = 4,
  TNN_ACT_RAMP = 5, TNN_ACT_TANH = 6, TNN_ACT_PLSE = 7, TNN_ACT_REVLEAKY = 8, TNN_ACT_LEAKY = 9,
  TNN_ACT_ELU = 10, TNN_ACT_LOGGY = 11, TNN_ACT_STAIR = 12, TNN_ACT_HARDTAN = 13, TNN_ACT_LHTAN = 14,
  TNN_ACT_SELU = 15, TNN_ACT_SOFTMAX = 16, TNN_ACT_GELU = 17, TNN_ACT_SWISH = 18, TNN_ACT_MISH = 19,
  TNN_ACT_HARD_MISH = 20, TNN_ACT_NORM_CHAN = 21, TNN_ACT_NORM_CHAN_SOFTMAX = 22,
  TNN_ACT_NORM_CHAN_SOFTMAX_MAXVAL = 23
} tnn_activation;

/* Contraction arithmetic of the conv / SGEMM kernels (replaces TNNCuda.useBLAS, nncuda.pas:91) */
typedef enum {
  TNN_MATH_FP32 = 0,   /* CUDA-core FMA, fp32 exact products (reference: cublasSgemm SIMT / sgemm1_nn) */
  TNN_MATH_3XTF32 = 1, /* tcgen05 kind::tf32, big/small split, 3 MMAs: fp32-level accuracy (default)   */
  TNN_MATH_TF32 = 2,   /* tcgen05 kind::tf32, single MMA (optional fast mode, 5e-3 tolerance)           */
  TNN_MATH_BF16X3 = 3  /* tcgen05 kind::f16 on bf16 big/small halves, 3 MMAs at twice the TF32 rate: the same 1e-4
                          gate as 3xTF32 (16+ mantissa bits per operand, fp32 accumulation).  Taken by the
                          channels-innermost convolution; every other contraction runs 3xTF32 in this mode.     */
} tnn_math_mode;

/* ------------------------------------------------------------------ lifetime / device */
/* TNNCuda.deviceCount nncuda.pas:443-446 */
TNN_API int32_t tnn_device_count(int32_t *count);
/* TNNCuda.Create(deviceIndex) nncuda.pas:448-467 (cudaSetDevice + stream; NVRTC/cuBLAS not needed) */
TNN_API int32_t tnn_create(int32_t device_index, tnn_context **ctx);
/* TNNCuda.Destroy nncuda.pas:469-476 */
TNN_API int32_t tnn_destroy(tnn_context *ctx);
/* message of the last failure on this thread (SAFE_CALL's assert text, nncuda.pas:237-274) */
TNN_API const char *tnn_last_error(void);
/* properties.name (read by CIFAR10Conv.pas:138) */
TNN_API const char *tnn_device_name(tnn_context *ctx);
/* properties.multiProcessorCount / major / minor */
TNN_API int32_t tnn_device_info(tnn_context *ctx, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor);
/* TNNCuda.stream field (nncuda.pas:78): adopt / read the CUstream all ops are enqueued on */
TNN_API int32_t tnn_set_stream(tnn_context *ctx, void *cuda_stream);
TNN_API void *tnn_get_stream(tnn_context *ctx);
/* TNNCuda.useBLAS (nncuda.pas:91) becomes the math mode of the contraction kernels */
TNN_API int32_t tnn_set_math_mode(tnn_context *ctx, int32_t mode);
TNN_API int32_t tnn_get_math_mode(tnn_context *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
TNN_API int64_t tnn_launch_count(tnn_context *ctx);
/* library build id string (arch, git-less) */
TNN_API const char *tnn_version(void);

/* ------------------------------------------------------------------ memory */
/* createDeviceBuffer(N bytes) nncuda.pas:488-491 */
TNN_API int32_t tnn_create_device_buffer(tnn_context *ctx, int64_t bytes, float **mem);
/* freeDeviceBuffer nncuda.pas:493-497 */
TNN_API int32_t tnn_free_device_buffer(tnn_context *ctx, float *mem);
/* readBuffer (D2H, synchronous) nncuda.pas:499-510 */
TNN_API int32_t tnn_read_buffer(tnn_context *ctx, const float *mem, size_t bytes, void *host);
/* writeBuffer (H2D, synchronous) nncuda.pas:512-522 */
TNN_API int32_t tnn_write_buffer(tnn_context *ctx, float *mem, size_t bytes, const void *host);
/* asynchronous variants for pinned host memory (no reference counterpart; used by the harness) */
TNN_API int32_t tnn_read_buffer_async(tnn_context *ctx, const float *mem, size_t bytes, void *host);
TNN_API int32_t tnn_write_buffer_async(tnn_context *ctx, float *mem, size_t bytes, const void *host);
/* finish nncuda.pas:1575-1584 */
TNN_API int32_t tnn_finish(tnn_context *ctx);
/* stream-to-stream ordering between two contexts of one device (no reference counterpart: the reference has a single
 * stream; used to overlap host<->device copies on a second TNNCuda instance with compute) */
typedef struct tnn_event tnn_event;
TNN_API int32_t tnn_event_create(tnn_context *ctx, tnn_event **ev);
TNN_API int32_t tnn_event_record(tnn_context *ctx, tnn_event *ev);       /* after all work queued on ctx so far */
TNN_API int32_t tnn_stream_wait_event(tnn_context *ctx, tnn_event *ev);  /* later work on ctx waits for ev    */
TNN_API int32_t tnn_event_destroy(tnn_event *ev);
TNN_API int32_t tnn_event_synchronize(tnn_event *ev);                    /* the HOST waits for ev            */
/* CUDA-graph capture of everything queued on the context's stream between begin and end (launch-bound inner loops:
 * a CIFAR training step is ~110 small kernels).  No allocation may happen while capturing: run the step once eagerly
 * first so the scratch buffers have their final size.  tnn_graph_launch replays the captured launches. */
typedef struct tnn_graph tnn_graph;
TNN_API int32_t tnn_graph_begin_capture(tnn_context *ctx);
TNN_API int32_t tnn_graph_end_capture(tnn_context *ctx, tnn_graph **graph);
TNN_API int32_t tnn_graph_launch(tnn_context *ctx, tnn_graph *graph);
TNN_API int32_t tnn_graph_destroy(tnn_graph *graph);
/* a 64-bit device counter for values that must change between replays of a graph (dropout seeds): *ctr += inc */
TNN_API int32_t tnn_counter_add(tnn_context *ctx, uint64_t *ctr, uint64_t inc);
/* page-locked host memory for tnn_write_buffer_async / tnn_read_buffer_async (a pageable source makes the "async"
 * copy synchronous); freed with tnn_host_free */
TNN_API int32_t tnn_host_alloc(tnn_context *ctx, int64_t bytes, void **host);
TNN_API int32_t tnn_host_free(tnn_context *ctx, void *host);

/* ------------------------------------------------------------------ activations (nncuda.pas:524-580) */
/* ActivateArray(N, x, offset, activation): in place; no-op for acLINEAR.  CPU twin: activate_array nactivation.pas:508-621 */
TNN_API int32_t tnn_activate_array(tnn_context *ctx, int64_t N, float *x, int64_t offset, int32_t activation);
/* activateArraySWISH(N, x, offset, output_sigmoid, output) nncuda.pas:544-560 */
TNN_API int32_t tnn_activate_array_swish(tnn_context *ctx, int64_t N, const float *x, int64_t offset,
                                         float *output_sigmoid, float *output);
/* DeriveArray(N, x(outputs), offset, activation, delta): delta *= f'(x).  CPU twin: gradient_array nactivation.pas:624-717 */
TNN_API int32_t tnn_derive_array(tnn_context *ctx, int64_t N, const float *x, int64_t offset, int32_t activation,
                                 float *delta);

/* ------------------------------------------------------------------ bias / scale (nncuda.pas:582-622,1468-1510) */
/* forwardBias(dstSize, dst, offset, srcSize, src, incb, batch): dst[i+offset] += src[(i/blockSize)%srcSize].  CPU twin ntensors.pas:7709-7727 */
TNN_API int32_t tnn_forward_bias(tnn_context *ctx, int64_t dstSize, float *dst, int64_t offset, int64_t srcSize,
                                 const float *src, int64_t incb, int64_t batch);
/* backwardBias(dstSize, dst, srcSize, src, srcOffset, incb, batch): dst[f] += sum over batch,block.  CPU twin addSums ntensors.pas:7729-7781 */
TNN_API int32_t tnn_backward_bias(tnn_context *ctx, int64_t dstSize, float *dst, int64_t srcSize, const float *src,
                                  int64_t srcOffset, int64_t incb, int64_t batch);
/* forwardScale nncuda.pas:1468-1488.  CPU twin ntensors.pas:7687-7707 */
TNN_API int32_t tnn_forward_scale(tnn_context *ctx, int64_t dstSize, float *dst, int64_t offset, int64_t scaleSize,
                                  const float *scale, int64_t incb, int64_t batch);
/* forwardScaleAdd nncuda.pas:1490-1510: dst = dst*scale[f] + bias[f] */
TNN_API int32_t tnn_forward_scale_add(tnn_context *ctx, int64_t dstSize, float *dst, int64_t offset,
                                      int64_t scaleSize, const float *scales, const float *biases, int64_t incb,
                                      int64_t batch);

/* ------------------------------------------------------------------ SGEMM (nncuda.pas:624-895) */
/* gemm: row-major C = beta*C + alpha*op(A)*op(B).  CPU twin cblas_sgemm ntensors.pas:2231-2286 */
TNN_API int32_t tnn_gemm(tnn_context *ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                         float alpha, const float *A, int64_t aOffset, int64_t lda, const float *B, int64_t bOffset,
                         int64_t ldb, float beta, float *C, int64_t cOffset, int64_t ldc);
/* gemmBatched (arrays of device pointers live in HOST memory, as PCUMem in nncuda.pas:727-807) */
TNN_API int32_t tnn_gemm_batched(tnn_context *ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                                 float alpha, const float *const *A, int64_t aOffset, int64_t lda,
                                 const float *const *B, int64_t bOffset, int64_t ldb, float beta, float *const *C,
                                 int64_t cOffset, int64_t ldc, int64_t batchCount);
/* gemmStridedBatched nncuda.pas:809-895.  strideC = 0 ACCUMULATES the batch sequentially into one C
 * (the wgrad call, nConvolutionLayer.pas:1228-1232), as the reference's non-BLAS host loop does. */
TNN_API int32_t tnn_gemm_strided_batched(tnn_context *ctx, int32_t transA, int32_t transB, int64_t M, int64_t N,
                                         int64_t K, float alpha, const float *A, int64_t aOffset, int64_t lda,
                                         int64_t strideA, const float *B, int64_t bOffset, int64_t ldb,
                                         int64_t strideB, float beta, float *C, int64_t cOffset, int64_t ldc,
                                         int64_t strideC, int64_t batchCount);

/* ------------------------------------------------------------------ BLAS-1 style (nncuda.pas:897-1054) */
TNN_API int32_t tnn_addvv(tnn_context *ctx, int64_t N, const float *src1, int64_t src1Offset, int64_t inca,
                          const float *src2, int64_t src2Offset, int64_t incb, float *dst, int64_t dstOffset,
                          int64_t incc);
TNN_API int32_t tnn_subvv(tnn_context *ctx, int64_t N, const float *src1, int64_t src1Offset, int64_t inca,
                          const float *src2, int64_t src2Offset, int64_t incb, float *dst, int64_t dstOffset,
                          int64_t incc);
TNN_API int32_t tnn_mulvv(tnn_context *ctx, int64_t N, const float *src1, int64_t src1Offset, int64_t inca,
                          const float *src2, int64_t src2Offset, int64_t incb, float *dst, int64_t dstOffset,
                          int64_t incc);
/* fmavv: dst = src1*src2 + src3 (the reference wrapper is an empty stub, nncuda.pas:937-943; kernel cuda_sgemm.cu:1671) */
TNN_API int32_t tnn_fmavv(tnn_context *ctx, int64_t N, const float *src1, int64_t src1Offset, int64_t inca,
                          const float *src2, int64_t src2Offset, int64_t incb, const float *src3,
                          int64_t src3Offset, int64_t incc, float *dst, int64_t dstOffset, int64_t incd);
/* axpy: y += a*x.  CPU twin ntensors.pas:7603-7671 */
TNN_API int32_t tnn_axpy(tnn_context *ctx, int64_t N, float a, const float *x, int64_t xOffset, int64_t incx,
                         float *y, int64_t yOffset, int64_t incy);
TNN_API int32_t tnn_power(tnn_context *ctx, int64_t N, const float *x, int64_t xOffset, int64_t incx, float a,
                          float *y, int64_t yOffset, int64_t incy);
/* scale: x *= a (a = 0 zeroes deltas, nnet.pas:291) */
TNN_API int32_t tnn_scale(tnn_context *ctx, int64_t N, float a, float *x, int64_t stride);
TNN_API int32_t tnn_fill(tnn_context *ctx, int64_t N, float *x, int64_t offset, float val, int64_t stride);
TNN_API int32_t tnn_copy(tnn_context *ctx, int64_t N, const float *src, int64_t srcOffset, int64_t inca, float *dst,
                         int64_t dstOffset, int64_t incb);
/* fmavss: dst[offset+i] = src[offset+i]*scalar + bias  nncuda.pas:1288-1300 */
TNN_API int32_t tnn_fmavss(tnn_context *ctx, int64_t N, const float *src, int64_t offset, float scalar, float bias,
                           float *dst);
TNN_API int32_t tnn_clamp(tnn_context *ctx, int64_t N, float alpha, const float *src, float *dst, int64_t stride,
                          int64_t offset);
/* inverseSqrt: dst = 1/sqrt(max(src,1e-6)) (alpha is ignored by the reference too, nncuda.pas:1563-1573) */
TNN_API int32_t tnn_inverse_sqrt(tnn_context *ctx, int64_t N, float alpha, const float *src, float *dst,
                                 int64_t stride, int64_t offset);

/* ------------------------------------------------------------------ losses / softmax / pooling / misc layers */
TNN_API int32_t tnn_cross_entropy_logistic(tnn_context *ctx, int64_t N, const float *pred, const float *truth,
                                           float *delta, float *error);
TNN_API int32_t tnn_cross_entropy_softmax(tnn_context *ctx, int64_t N, const float *pred, const float *truth,
                                          float *delta, float *error);
TNN_API int32_t tnn_cost_l2(tnn_context *ctx, int64_t N, const float *pred, const float *truth, float *delta,
                            float *error);
/* softmaxBatch nncuda.pas:1056-1073 */
TNN_API int32_t tnn_softmax_batch(tnn_context *ctx, int64_t N, const float *input, int64_t iOffset, int64_t batch,
                                  int64_t batch_size, int64_t groups, int64_t group_size, int64_t stride, float temp,
                                  float *output, int64_t oOffset);
/* forwardMaxPool / backwardMaxPool nncuda.pas:1088-1142; indexes are 64-bit (size_t in cuda_sgemm.cu:868) */
TNN_API int32_t tnn_forward_maxpool(tnn_context *ctx, int64_t aBatch, int64_t outC, int64_t outH, int64_t outW,
                                    const float *input, int64_t c, int64_t h, int64_t w, int64_t stride_x,
                                    int64_t stride_y, int64_t padding, int64_t kernelSize, uint64_t *indexes,
                                    float *output);
TNN_API int32_t tnn_backward_maxpool(tnn_context *ctx, int64_t aBatch, int64_t outC, int64_t outH, int64_t outW,
                                     float *output, const uint64_t *indexes, const float *delta);
/* upSample nncuda.pas:1267-1286 (in = low-res, out = high-res; isForward=0 accumulates into in) */
TNN_API int32_t tnn_upsample(tnn_context *ctx, int64_t aBatch, int64_t aChannels, int64_t outHeight,
                             int64_t outWidth, float *in, int64_t stride, int32_t isForward, float scale, float *out,
                             int32_t zeroIn);
/* forwardDropout / backwardDropout nncuda.pas:1512-1538; `seed` replaces the wrapper's random($10000000) */
TNN_API int32_t tnn_forward_dropout(tnn_context *ctx, int64_t N, const float *src, float probability, float scale,
                                    float *rnd, float *dst, uint64_t seed);
/* the same with the seed read on the device: seed = *seed_dev + seed_offset (graph replays) */
TNN_API int32_t tnn_forward_dropout_dev(tnn_context *ctx, int64_t N, const float *src, float probability, float scale,
                                        float *rnd, float *dst, const uint64_t *seed_dev, uint64_t seed_offset);

/* One launch for the element-wise part of an LSTM step (replaces the 15 operator calls of TLSTMLayer.forwardGPU,
 * source/nlstmlayer.pas:934-973).  wz = h(t-1) W^T, [batch][4*hidden] in gate order f|i|g|o without bias; bw[4*hidden]
 * the hidden-side biases; uz + uzOffset = x(t) U^T + bu in the same layout.  Updates c and h [batch][hidden] in place
 * and stores them at cell_out + outOffset / h_out + outOffset (offsets in elements). */
TNN_API int32_t tnn_lstm_gates(tnn_context *ctx, int64_t batch, int64_t hidden, const float *wz, const float *bw,
                               const float *uz, int64_t uzOffset, float *c, float *h, float *cell_out, float *h_out,
                               int64_t outOffset);
/* One launch for the element-wise part of an LSTM step of back-propagation through time (TLSTMLayer.backward,
 * source/nlstmlayer.pas:661-812, plus the sub-layers' delta.Clamp(-1,1), nconnectedlayer.pas:262).  wz / uz hold the gate
 * pre-activations of ALL steps ([steps*batch][4*hidden], gate order f|i|g|o; wz without bias, uz with the input-side
 * bias) and zOffset selects the step; cell holds c of all steps (cellOffset = this step, prevCellOffset = the step
 * before, or -1 at the first step); dh + dhOffset = dL/dh of this step; dc [batch][hidden] is the carried dL/dc
 * (updated in place); dz + dzOffset receives the four pre-activation deltas [batch][4*hidden]. */
TNN_API int32_t tnn_lstm_gates_backward(tnn_context *ctx, int64_t batch, int64_t hidden, const float *wz, const float *bw,
                                        const float *uz, int64_t zOffset, const float *cell, int64_t cellOffset,
                                        int64_t prevCellOffset, const float *dh, int64_t dhOffset, float *dc, float *dz,
                                        int64_t dzOffset);
TNN_API int32_t tnn_backward_dropout(tnn_context *ctx, int64_t N, const float *src, float probability, float scale,
                                     const float *rnd, float *dst);

/* ------------------------------------------------------------------ im2col / col2im (nncuda.pas:1144-1265) */
/* im2col: col[c][kh][kw][oh][ow] (the CPU layout, ntensors.pas:11415-11491), one image */
TNN_API int32_t tnn_im2col(tnn_context *ctx, int64_t aChannels, int64_t aHeight, int64_t aWidth, int64_t kernelHeight,
                           int64_t kernelWidth, int64_t padHeight, int64_t padWidth, int64_t strideY, int64_t strideX,
                           int64_t dilationY, int64_t dilationX, const float *im, int64_t imOffset, float *col,
                           int64_t colOffset);
/* col2im: scatter-ADD of col into im (ntensors.pas:11650-11763), one image */
TNN_API int32_t tnn_col2im(tnn_context *ctx, int64_t aChannels, int64_t aHeight, int64_t aWidth, int64_t kernelHeight,
                           int64_t kernelWidth, int64_t padHeight, int64_t padWidth, int64_t strideY, int64_t strideX,
                           int64_t dilationY, int64_t dilationX, const float *col, int64_t colOffset, float *im,
                           int64_t imOffset);

/* ------------------------------------------------------------------ batch-norm pieces (nncuda.pas:1302-1466) */
/* means / variances / meansAndVars(srcSize, dstSize, groups, src, offset, means[, vars]).
 * blockSize = srcSize/(dstSize*groups); variance is UNBIASED (/(S-1)) like ntensors.pas:9102-9177 */
TNN_API int32_t tnn_means(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups, const float *src,
                          int64_t offset, float *means);
TNN_API int32_t tnn_variances(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups, const float *src,
                              int64_t offset, const float *means, float *vars);
TNN_API int32_t tnn_means_and_vars(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups,
                                   const float *src, int64_t offset, float *means, float *vars);
/* normalize(srcSize(=channels), dstSize, groups, means, meansStride, vars, varsStride, dst, dstOffset):
 * dst = (dst-mean[c]) / max(sqrt(var[c]),1e-6)  (ntensors.pas:4357-4382; sqrt(max(var,1e-6)) when blockSize=1) */
TNN_API int32_t tnn_normalize(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups, const float *means,
                              int64_t meansStride, const float *vars, int64_t varsStride, float *dst,
                              int64_t dstOffset);
/* meansAndVarsDelta nncuda.pas:1398-1419.  CPU twin sMeanAndVarianceDelta ntensors.pas:8831-8899 */
TNN_API int32_t tnn_means_and_vars_delta(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups,
                                         const float *delta, const float *x, int64_t offset, const float *mean,
                                         const float *variance, float *mean_delta, float *variance_delta);
/* normalizeDelta nncuda.pas:1421-1446.  CPU twin sNormalizeDelta ntensors.pas:8902-8951 */
TNN_API int32_t tnn_normalize_delta(tnn_context *ctx, int64_t deltaSize, int64_t meanSize, int64_t groups,
                                    float *delta, const float *x, int64_t offset, const float *mean,
                                    const float *variance, const float *mean_delta, const float *variance_delta);
/* addDots(N, dstSize, groups, src1, src2, srcOffset, dst): dst[c] += sum src1*src2.  CPU twin ntensors.pas:7783-7837 */
TNN_API int32_t tnn_add_dots(tnn_context *ctx, int64_t N, int64_t dstSize, int64_t groups, const float *src1,
                             const float *src2, int64_t srcOffset, float *dst);

/* ------------------------------------------------------------------ coarse, fused entry points
 * New with this backend: the bodies of TConvolutionalLayer.forwardGPU / backwardGPU / updateGPU
 * (nConvolutionLayer.pas:1022-1330) and TBaseLayer.batchNormGPU (nbaselayer.pas:550-660) collapse into
 * these calls (INTEGRATION.md shows the edited method bodies).  No column buffer is materialised. */

typedef struct {
  int64_t batch, c, h, w;          /* input  [batch][c][h][w]                                   */
  int64_t filters, kh, kw;         /* weights [filters][c*kh*kw]                                */
  int64_t pad_h, pad_w, stride_y, stride_x, dilation_y, dilation_x;
} tnn_conv_desc;

/* forwardGPU steps 2-5 (nConvolutionLayer.pas:1040-1122) in one kernel:
 *   out = act( conv(in, weights) * scale[f] + bias[f] )
 * `scale` may be NULL (1.0); pass scale[f] = scales[f]/max(sqrt(rolling_variance[f]),1e-6) and the matching
 * folded bias for inference-mode batch-norm, or the fuseBatchNorm'ed weights (nConvolutionLayer.pas:102-126)
 * with scale = NULL.  `bias` may be NULL.  activation: TNN_ACT_LINEAR/LEAKY/LOGISTIC/RELU/... (elementwise ones). */
TNN_API int32_t tnn_conv2d_forward(tnn_context *ctx, const tnn_conv_desc *d, const float *in, int64_t inOffset,
                                   const float *weights, const float *scale, const float *bias,
                                   int32_t activation, float *out, int64_t outOffset);
/* The same with the following shortcut layer folded in (TAddLayer.forwardGPU naddlayer.pas:885-910 with a linear
 * activation: output = conv_output + layers[from].output):
 *   out = act( conv(in, weights) * scale[f] + bias[f] ) + residual        (residual has the shape of out) */
TNN_API int32_t tnn_conv2d_forward_residual(tnn_context *ctx, const tnn_conv_desc *d, const float *in, int64_t inOffset,
                                            const float *weights, const float *scale, const float *bias,
                                            int32_t activation, const float *residual, int64_t residualOffset,
                                            float *out, int64_t outOffset);
/* Declare a weight tensor immutable (is_static = 1) or mutable again (0).  While static, the convolution entry points
 * keep their re-ordered / TF32-split copy of it across calls instead of re-deriving it per call (inference: the
 * 75 prepare launches of a YOLOv3 step disappear).  The caller must not modify the buffer while it is static;
 * declaring it again (either value) drops the cached copies — do that after changing the contents. */
TNN_API int32_t tnn_set_weights_static(tnn_context *ctx, const float *weights, int32_t is_static);
/* Which kernel the last tnn_conv2d_forward* call on this context dispatched to (measurement tools attribute layer
 * times to kernels with it): 0 none yet, 1 direct fp32 (k_conv_direct), 2 tcgen05 gathering implicit GEMM (k_umma
 * with Im2colPixelLoader), 3 tcgen05 with TMA-fed activations (k_umma with TmaPatchLoader), 4 tcgen05 halo-tile
 * (k_umma_halo), 5 CUDA-core implicit GEMM (simt::k_gemm, TNN_MATH_FP32), 7 tcgen05 with TMA im2col-mode
 * activations on channels-innermost tensors (nhwc::k_conv_nhwc, one CTA per tile), 8 direct fp32 writing
 * channels-innermost rows, 9 the image layer on the tensor cores (nhwc::k_conv_image), 11 the same as 7 on CTA pairs
 * (nhwc::k_conv_nhwc_pair, tcgen05.mma.cta_group::2). */
TNN_API int32_t tnn_last_conv_kernel(tnn_context *ctx);
/* backwardGPU wgrad (nConvolutionLayer.pas:1203-1232): weight_updates += sum_b delta_b * im2col(in_b)^T */
TNN_API int32_t tnn_conv2d_backward_filter(tnn_context *ctx, const tnn_conv_desc *d, const float *in,
                                           int64_t inOffset, const float *delta, int64_t deltaOffset,
                                           float *weight_updates);
/* backwardGPU dgrad (nConvolutionLayer.pas:1250-1275): in_delta += col2im(weights^T * delta_b) */
TNN_API int32_t tnn_conv2d_backward_data(tnn_context *ctx, const tnn_conv_desc *d, const float *weights,
                                         const float *delta, int64_t deltaOffset, float *in_delta,
                                         int64_t inDeltaOffset);
/* training-mode batchNormGPU (nbaselayer.pas:583-643) fused: statistics, rolling update
 * (r = (1-m) r + m stat), x := out, out := normalised, x_norm := out, then out = out*scale+bias and
 * the layer activation.  All per-channel vectors have `channels` entries. */
TNN_API int32_t tnn_batchnorm_forward_training(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize,
                                               float *out, float *x, float *x_norm, float *mean, float *variance,
                                               float *rolling_mean, float *rolling_variance, float bn_momentum,
                                               const float *scales, const float *biases, int32_t activation);
/* inference-mode batchNormGPU (nbaselayer.pas:628-643) + activation in one pass */
TNN_API int32_t tnn_batchnorm_forward_inference(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize,
                                                float *out, const float *rolling_mean,
                                                const float *rolling_variance, const float *scales,
                                                const float *biases, int32_t activation);
/* DeriveArray + batchNormBackGPU (nConvolutionLayer.pas:1186, nbaselayer.pas:662-714) fused:
 * delta *= f'(out); scale_updates += sum x_norm*delta; delta *= scale; mean_delta, variance_delta; delta normalised */
TNN_API int32_t tnn_batchnorm_backward(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize,
                                       const float *out, int32_t activation, float *delta, const float *x,
                                       const float *x_norm, const float *mean, const float *variance,
                                       const float *scales, float *scale_updates, float *mean_delta,
                                       float *variance_delta);
/* updateGPU (nConvolutionLayer.pas:1309-1322) for one parameter tensor in one kernel:
 *   updates += decay_term*params ; params += rate*updates ; updates *= momentum
 * with decay_term = -decay*batch for weights, 0 for biases/scales; rate = learning_rate/batch. */
TNN_API int32_t tnn_sgd_update(tnn_context *ctx, int64_t N, float *params, float *updates, float rate,
                               float momentum, float decay_term);
/* The same update for `count` parameter tensors in one launch per 32 tensors (TNNet.update, nnet.pas:378-403, walks the
 * layers and each updateGPU touches two or three tensors): params[i], updates[i], sizes[i], rates[i], decay_terms[i] are
 * HOST arrays of device pointers / element counts / per-tensor rate and decay term. */
TNN_API int32_t tnn_sgd_update_multi(tnn_context *ctx, int32_t count, float *const *params, float *const *updates,
                                     const uint64_t *sizes, const float *rates, const float *decay_terms,
                                     float momentum);
/* TYoloLayer.forwardGPU's per-(image, anchor) ActivateArray / fmavss loop (nyololayer.pas:1081-1099) in one
 * launch: logistic on the x, y, objectness and class planes of out[batch][n*(5+classes)][wh]; x,y are then
 * mapped v*scale_xy + 0.5*(1-scale_xy) when scale_xy != 1. */
TNN_API int32_t tnn_yolo_activate(tnn_context *ctx, int64_t batch, int64_t n, int64_t classes, int64_t wh,
                                  float *out, float scale_xy);
/* The candidate test of TYoloLayer.getDetections (nyololayer.pas:414-418, `if objectness > thresh`) on the device, so
 * that the host reads a short list instead of the whole output tensor: every (image, anchor, cell) of the ACTIVATED
 * output out[batch][n*(5+classes)][wh] whose objectness exceeds thresh appends one record to `cand`:
 *   cand[0]                       uint32 count of candidates found (zero it before the first call of a step; it keeps
 *                                 counting past maxCand, the list then holds the first maxCand to arrive -> read the
 *                                 full output instead)
 *   cand + 4 + r*(classes + 9)    record r: int32 {tag, image, anchor, cell}, then the classes + 5 entries of the cell
 * Several layers may append to one list (tag tells them apart).  Record order is arrival order: sort by
 * (tag, image, cell, anchor) to get the order in which the reference's loops visit them. */
TNN_API int32_t tnn_yolo_candidates(tnn_context *ctx, int64_t batch, int64_t n, int64_t classes, int64_t wh,
                                    const float *out, float thresh, int32_t tag, int64_t maxCand, float *cand);
/* out[b][cOff..cOff+c] = in  — per-image channel concat used by route layers */
TNN_API int32_t tnn_concat_channels(tnn_context *ctx, int64_t batch, int64_t c, int64_t blockSize, const float *in,
                                    int64_t outChannels, int64_t cOffset, float *out);

/* ------------------------------------------------------------------ channels-innermost (NHWC) inference graph
 * New with this backend.  The reference keeps every activation planar (NCHW) because its GEMM consumes a
 * materialised im2col matrix; on B200 the convolution reads its operands through the TMA unit, whose im2col mode
 * wants the channel index innermost.  A graph may therefore keep its INTERNAL activations as rows of channels,
 *   act[(b*H + h)*W + w][pitch],  channels 0..C-1 of a pixel contiguous, `pitch` >= C floats between pixels
 * (pitch > C lets a layer write straight into its slice of a route/concat buffer), and convert at the graph edges:
 * the first layer reads the planar image, the head layers write planar outputs for the yolo layer / host decode.
 * Same arithmetic as tnn_conv2d_forward (same products, fp32 accumulation, scale/bias/activation epilogue). */
typedef enum { TNN_LAYOUT_NCHW = 0, TNN_LAYOUT_NHWC = 1, TNN_LAYOUT_NHWC_SPLIT = 2 } tnn_layout;
/* forwardGPU steps 2-5 (nConvolutionLayer.pas:1040-1122) on channels-innermost activations:
 *   out = act( conv(in, weights) * scale[f] + bias[f] ),  weights [filters][c*kh*kw] as always.
 * inLayout NHWC: inPitch a multiple of 4 floats, 16-byte aligned base (channels past c read as zero) — tcgen05 implicit GEMM, activations and
 *   weights by TMA (im2col mode), output by TMA store (outLayout NHWC) or as planes (outLayout NCHW, outPitch ignored).
 * inLayout NCHW: only for small contractions (c*kh*kw <= 64, the image layer) with outLayout NHWC. */
TNN_API int32_t tnn_conv2d_forward_nhwc(tnn_context *ctx, const tnn_conv_desc *d, const float *in, int64_t inOffset,
                                        int64_t inPitch, int32_t inLayout, const float *weights, const float *scale,
                                        const float *bias, int32_t activation, float *out, int64_t outOffset,
                                        int64_t outPitch, int32_t outLayout);
/* The same with the following shortcut layer folded in (TAddLayer.forwardGPU naddlayer.pas:885-910, linear activation):
 *   out = act( conv(in, weights) * scale[f] + bias[f] ) + residual,   residual rows of residualPitch floats
 * (channels-innermost in and out; filters a multiple of 4).  The add kernel and one round trip of the tensor through
 * HBM disappear; the arithmetic is the separate layers' (one fp32 add after the activation). */
TNN_API int32_t tnn_conv2d_forward_nhwc_residual(tnn_context *ctx, const tnn_conv_desc *d, const float *in,
                                                 int64_t inOffset, int64_t inPitch, int32_t inLayout, const float *weights,
                                                 const float *scale, const float *bias, int32_t activation,
                                                 const float *residual, int64_t residualOffset, int64_t residualPitch,
                                                 float *out, int64_t outOffset, int64_t outPitch, int32_t outLayout);
/* Pre-split activations (TNN_MATH_BF16X3).  The BF16x3 arithmetic multiplies bf16 halves of every fp32 operand,
 *   hi = bf16_rn(x), lo = bf16_rn(x - hi)   (x = hi + lo to 2^-18 relative),
 * which the kernel above derives from the fp32 tile inside shared memory for every kernel tap.  A graph may instead keep
 * an INTERNAL activation as those two bf16 planes (TNN_LAYOUT_NHWC_SPLIT: rows of channels, `pitch` bf16 elements
 * between pixels, the lo plane `plane` elements after the hi plane; 4 bytes per element like fp32): the producing
 * layer's epilogue writes them (`outSplit`), the consuming convolution's TMA lands them directly in the tensor-core
 * layout and nothing converts — the same products and the same fp32 accumulation as the fp32-input form, bit for bit.
 * `out` (fp32 rows or planes) and `outSplit` may both be given (a tensor that is also a later shortcut's operand or a
 * graph output); either may be NULL.  in: const float* for NCHW / NHWC, the hi plane for NHWC_SPLIT (inPitch, inPlane in
 * bf16 elements, multiples of 8).  residual: fp32 rows as in tnn_conv2d_forward_nhwc_residual, or NULL. */
TNN_API int32_t tnn_conv2d_forward_split(tnn_context *ctx, const tnn_conv_desc *d, const void *in, int64_t inOffset,
                                         int64_t inPitch, int32_t inLayout, int64_t inPlane, const float *weights,
                                         const float *scale, const float *bias, int32_t activation,
                                         const float *residual, int64_t residualOffset, int64_t residualPitch,
                                         float *out, int64_t outOffset, int64_t outPitch, int32_t outLayout,
                                         void *outSplit, int64_t outSplitOffset, int64_t outSplitPitch,
                                         int64_t outSplitPlane);
/* The same, with the folded shortcut's other operand in either form: residualLayout TNN_LAYOUT_NHWC (fp32 rows,
 * residualPitch in floats, residualPlane ignored) or TNN_LAYOUT_NHWC_SPLIT (the hi plane; offset, pitch and plane in bf16
 * elements, multiples of 8; filters a multiple of 8).  From planes the operand is hi + lo — the value the planes hold,
 * i.e. the fp32 tensor to 2^-18 relative — so a graph can keep EVERY tensor between convolutions as planes only
 * (written once, 4 bytes per element, no fp32 twin for the shortcuts) and every convolution, 1x1 included, reads its
 * operands by TMA with no conversion in shared memory. */
TNN_API int32_t tnn_conv2d_forward_split_residual(tnn_context *ctx, const tnn_conv_desc *d, const void *in, int64_t inOffset,
                                                  int64_t inPitch, int32_t inLayout, int64_t inPlane, const float *weights,
                                                  const float *scale, const float *bias, int32_t activation,
                                                  const void *residual, int64_t residualOffset, int64_t residualPitch,
                                                  int32_t residualLayout, int64_t residualPlane, float *out,
                                                  int64_t outOffset, int64_t outPitch, int32_t outLayout, void *outSplit,
                                                  int64_t outSplitOffset, int64_t outSplitPitch, int64_t outSplitPlane);
/* A yolo head: the linear 1x1 convolution in front of a yolo layer WITH that layer's activation folded into its epilogue
 * (TYoloLayer.forwardGPU nyololayer.pas:1068-1112: copy + logistic on x, y, objectness and class entries, x / y mapped
 * v*scale_xy + 0.5*(1 - scale_xy)) — out is the yolo layer's output [batch][anchors*(classes+5)][OH*OW] (planes), the
 * copy and the activation pass over it disappear.  Same expressions as tnn_yolo_activate, same bits. */
TNN_API int32_t tnn_conv2d_forward_yolo_head(tnn_context *ctx, const tnn_conv_desc *d, const void *in, int64_t inOffset,
                                             int64_t inPitch, int32_t inLayout, int64_t inPlane, const float *weights,
                                             const float *scale, const float *bias, float *out, int64_t outOffset,
                                             int64_t classes, float scale_xy);
/* fp32 rows of channels <-> the pre-split planes (the way back is hi + lo, exact in fp32); channels a multiple of 8 */
TNN_API int32_t tnn_nhwc_to_split(tnn_context *ctx, int64_t rows, int64_t channels, const float *in, int64_t inPitch,
                                  void *outSplit, int64_t outPitch, int64_t outPlane);
TNN_API int32_t tnn_split_to_nhwc(tnn_context *ctx, int64_t rows, int64_t channels, const void *inSplit, int64_t inPitch,
                                  int64_t inPlane, float *out, int64_t outPitch);
/* [batch][channels][blockSize] -> [batch][blockSize][outPitch] and back */
TNN_API int32_t tnn_nchw_to_nhwc(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize, const float *in,
                                 float *out, int64_t outPitch);
TNN_API int32_t tnn_nhwc_to_nchw(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize, const float *in,
                                 int64_t inPitch, float *out);
/* rows of `channels` floats with independent pitches: out = in (route / concat slices, nconcatlayer.pas:214-220) and
 * out = a + b (shortcut, naddlayer.pas:891-910) */
TNN_API int32_t tnn_rows_copy(tnn_context *ctx, int64_t rows, int64_t channels, const float *in, int64_t inPitch,
                              float *out, int64_t outPitch);
TNN_API int32_t tnn_rows_add(tnn_context *ctx, int64_t rows, int64_t channels, const float *a, int64_t aPitch,
                             const float *b, int64_t bPitch, float *out, int64_t outPitch);
/* upSample forward (nupsamplelayer.pas:190-206) on channels-innermost rows: out[b][h*s+i][w*s+j][:] = scale*in[b][h][w][:] */
TNN_API int32_t tnn_upsample_nhwc(tnn_context *ctx, int64_t batch, int64_t channels, int64_t inHeight, int64_t inWidth,
                                  int64_t stride, float scale, const float *in, int64_t inPitch, float *out,
                                  int64_t outPitch);

/* ------------------------------------------------------------------ multi-GPU exchange (data-parallel training)
 * New with this backend: the reference is single-device ("todo ... multi GPU", nConvolutionLayer.pas:472,498).  The
 * path shards along the batch; its ONE exchange step is the sum of the update buffers (weight_updates / bias_updates /
 * scale_updates of every layer, laid out in one flat device buffer) over the replicas, between TNNet.backward and
 * TNNet.update (nnet.pas:371-403).  One process (or one TNNCuda instance + host thread) per GPU; NCCL over NVLink.
 * Rank 0 creates the 128-byte id and ships it to the other ranks by whatever means the host has. */
typedef struct tnn_comm tnn_comm;
TNN_API int32_t tnn_comm_unique_id(void *id, int64_t bytes);                 /* bytes >= 128 */
TNN_API int32_t tnn_comm_init_rank(tnn_context *ctx, int32_t nranks, int32_t rank, const void *id, tnn_comm **comm);
TNN_API int32_t tnn_comm_destroy(tnn_comm *comm);
TNN_API int32_t tnn_comm_info(tnn_comm *comm, int32_t *nranks, int32_t *rank, int32_t *nccl_version);
/* buf[offset .. offset+n) := sum over ranks, in place, enqueued on the context's stream (capturable in a tnn_graph) */
TNN_API int32_t tnn_allreduce_sum(tnn_context *ctx, tnn_comm *comm, float *buf, int64_t offset, int64_t n);
/* buf[offset .. offset+n) := rank `root`'s values (replicas start from the same weights) */
TNN_API int32_t tnn_broadcast(tnn_context *ctx, tnn_comm *comm, float *buf, int64_t offset, int64_t n, int32_t root);
/* batch-norm over a batch that is sharded across replicas (SyncBN): the statistics of the GLOBAL mini-batch, as the
 * single-device reference computes them (nbaselayer.pas:352; two-pass MeansAndVars ntensors.pas:9102-9177), by
 * exchanging the per-channel sums on the context's stream.  comm may be NULL (single device: the same two-pass
 * arithmetic).  After backward, mean_delta / variance_delta hold the exchanged raw sums (not the reference's scaled
 * values); scale_updates receives the LOCAL sum (the update bucket is summed over the replicas afterwards). */
TNN_API int32_t tnn_batchnorm_forward_training_sync(tnn_context *ctx, tnn_comm *comm, int64_t globalBatch, int64_t batch,
                                                    int64_t channels, int64_t blockSize, float *out, float *x,
                                                    float *x_norm, float *mean, float *variance, float *rolling_mean,
                                                    float *rolling_variance, float bn_momentum, const float *scales,
                                                    const float *biases, int32_t activation);
TNN_API int32_t tnn_batchnorm_backward_sync(tnn_context *ctx, tnn_comm *comm, int64_t globalBatch, int64_t batch,
                                            int64_t channels, int64_t blockSize, const float *out, int32_t activation,
                                            float *delta, const float *x, const float *x_norm, const float *mean,
                                            const float *variance, const float *scales, float *scale_updates,
                                            float *mean_delta, float *variance_delta);

#ifdef __cplusplus
}
#endif
#endif /* TENSORIUM_B200_H */
