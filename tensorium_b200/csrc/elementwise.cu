// elementwise.cu — HBM-bound elementwise ops of the TNNCuda surface.
// All are one read (+ one write) per element; contiguous cases use 128-bit loads/stores with a
// scalar head/tail for misaligned offsets, strided cases a scalar grid-stride loop.
// Replaces kernels 4-7, 9-15, 26, 34-41, 45-46 of the reference (source/cuda_sgemm.cu).
#include "common.cuh"
#include "plane_map.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 16;

inline int grid_for(int64_t work_items, int per_block) {
  int64_t b = cdiv64(work_items, per_block);
  if (b < 1) b = 1;
  if (b > kMaxBlocks) b = kMaxBlocks;
  return (int)b;
}

// ---- contiguous map over up to 3 arrays sharing the same 16B phase -------------------------------
// f(d, a, b) -> new d.  `head` scalar elements bring all pointers to 16B alignment.
template <bool READ_D, int NSRC, typename F>
__global__ void __launch_bounds__(kThreads) k_map_vec(float *d, const float *a, const float *b, int64_t n,
                                                      int head, F f) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (tid < head) {
    float dv = READ_D ? d[tid] : 0.f;
    d[tid] = f(dv, NSRC > 0 ? a[tid] : 0.f, NSRC > 1 ? b[tid] : 0.f);
  }
  const int64_t nvec = (n - head) >> 2;
  float4 *d4 = reinterpret_cast<float4 *>(d + head);
  const float4 *a4 = reinterpret_cast<const float4 *>(a + head);
  const float4 *b4 = reinterpret_cast<const float4 *>(b + head);
  for (int64_t i = tid; i < nvec; i += nthreads) {
    float4 dv = READ_D ? d4[i] : make_float4(0, 0, 0, 0);
    float4 av = NSRC > 0 ? a4[i] : make_float4(0, 0, 0, 0);
    float4 bv = NSRC > 1 ? b4[i] : make_float4(0, 0, 0, 0);
    dv.x = f(dv.x, av.x, bv.x);
    dv.y = f(dv.y, av.y, bv.y);
    dv.z = f(dv.z, av.z, bv.z);
    dv.w = f(dv.w, av.w, bv.w);
    d4[i] = dv;
  }
  const int64_t tail0 = head + (nvec << 2);
  if (tail0 + tid < n) {
    int64_t i = tail0 + tid;
    float dv = READ_D ? d[i] : 0.f;
    d[i] = f(dv, NSRC > 0 ? a[i] : 0.f, NSRC > 1 ? b[i] : 0.f);
  }
}

// ---- strided / misaligned scalar map ---------------------------------------------------------------
template <bool READ_D, int NSRC, typename F>
__global__ void __launch_bounds__(kThreads) k_map_strided(float *d, int64_t incd, const float *a, int64_t inca,
                                                          const float *b, int64_t incb, int64_t n, F f) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
    float dv = READ_D ? d[i * incd] : 0.f;
    d[i * incd] = f(dv, NSRC > 0 ? a[i * inca] : 0.f, NSRC > 1 ? b[i * incb] : 0.f);
  }
}

template <bool READ_D, int NSRC, typename F>
int launch_map(tnn_context *ctx, float *d, int64_t incd, const float *a, int64_t inca, const float *b, int64_t incb,
               int64_t n, F f) {
  if (n <= 0) return 0;
  if (NSRC < 1) { a = d; inca = incd; }
  if (NSRC < 2) { b = d; incb = incd; }
  const uintptr_t pd = (uintptr_t)d, pa = (uintptr_t)a, pb = (uintptr_t)b;
  const bool contiguous = incd == 1 && inca == 1 && incb == 1;
  const bool same_phase = ((pd ^ pa) & 15) == 0 && ((pd ^ pb) & 15) == 0 && (pd & 3) == 0;
  if (contiguous && same_phase) {
    int head = (int)(((16 - (pd & 15)) & 15) >> 2);
    if (head > n) head = (int)n;
    int blocks = grid_for(cdiv64(n, 4) + 8, kThreads);
    k_map_vec<READ_D, NSRC, F><<<blocks, kThreads, 0, ctx->stream>>>(d, a, b, n, head, f);
  } else {
    int blocks = grid_for(n, kThreads);
    k_map_strided<READ_D, NSRC, F><<<blocks, kThreads, 0, ctx->stream>>>(d, incd, a, inca, b, incb, n, f);
  }
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // namespace

__global__ void __launch_bounds__(kThreads) k_sgd(int64_t n, float *__restrict__ w, float *__restrict__ u, float rate,
                                                  float momentum, float decay_term, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  auto step = [&](float &wv, float &uv) {
    // weight_updates += decay_term*weights ; weights += rate*weight_updates ; weight_updates *= momentum
    if (decay_term != 0.f) uv = __fadd_rn(__fmul_rn(decay_term, wv), uv);
    wv = __fadd_rn(__fmul_rn(rate, uv), wv);
    uv = momentum * uv;
  };
  if (vec_ok) {
    const int64_t nvec = n >> 2;
    float4 *w4 = reinterpret_cast<float4 *>(w), *u4 = reinterpret_cast<float4 *>(u);
    for (int64_t i = tid; i < nvec; i += nthreads) {
      float4 wv = w4[i], uv = u4[i];
      step(wv.x, uv.x); step(wv.y, uv.y); step(wv.z, uv.z); step(wv.w, uv.w);
      w4[i] = wv; u4[i] = uv;
    }
    const int64_t i = (nvec << 2) + tid;
    if (i < n) { float wv = w[i], uv = u[i]; step(wv, uv); w[i] = wv; u[i] = uv; }
  } else {
    for (int64_t i = tid; i < n; i += nthreads) { float wv = w[i], uv = u[i]; step(wv, uv); w[i] = wv; u[i] = uv; }
  }
}

// the same update for up to kSgdMulti parameter tensors in ONE launch (blockIdx.y = tensor): a training step of a small
// net is bound by its launch count, and updateGPU contributes two or three launches per layer
constexpr int kSgdMulti = 32;
struct SgdItems {
  float *w[kSgdMulti];
  float *u[kSgdMulti];
  int64_t n[kSgdMulti];
  float rate[kSgdMulti];
  float decay[kSgdMulti];
};
__global__ void __launch_bounds__(kThreads) k_sgd_multi(const __grid_constant__ SgdItems it, float momentum) {
  const int t = blockIdx.y;
  const int64_t n = it.n[t];
  float *__restrict__ w = it.w[t];
  float *__restrict__ u = it.u[t];
  const float rate = it.rate[t], decay_term = it.decay[t];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  auto step = [&](float &wv, float &uv) {
    if (decay_term != 0.f) uv = __fadd_rn(__fmul_rn(decay_term, wv), uv);
    wv = __fadd_rn(__fmul_rn(rate, uv), wv);
    uv = momentum * uv;
  };
  if (((((uintptr_t)w) | ((uintptr_t)u)) & 15) == 0) {
    const int64_t nvec = n >> 2;
    float4 *w4 = reinterpret_cast<float4 *>(w), *u4 = reinterpret_cast<float4 *>(u);
    for (int64_t i = tid; i < nvec; i += nthreads) {
      float4 wv = w4[i], uv = u4[i];
      step(wv.x, uv.x); step(wv.y, uv.y); step(wv.z, uv.z); step(wv.w, uv.w);
      w4[i] = wv; u4[i] = uv;
    }
    const int64_t i = (nvec << 2) + tid;
    if (i < n) { float wv = w[i], uv = u[i]; step(wv, uv); w[i] = wv; u[i] = uv; }
  } else {
    for (int64_t i = tid; i < n; i += nthreads) { float wv = w[i], uv = u[i]; step(wv, uv); w[i] = wv; u[i] = uv; }
  }
}

int tnn_sgd_launch(tnn_context *ctx, int64_t N, float *params, float *updates, float rate, float momentum,
                   float decay_term) {
  const int vec_ok = ((((uintptr_t)params) | ((uintptr_t)updates)) & 15) == 0;
  k_sgd<<<grid_for(cdiv64(N, 4), kThreads), kThreads, 0, ctx->stream>>>(N, params, updates, rate, momentum,
                                                                        decay_term, vec_ok);
  TNN_LAUNCHED(ctx);
  return 0;
}

extern "C" {

// nncuda.pas:524-542 / activate_array nactivation.pas:508-621
int32_t tnn_activate_array(tnn_context *ctx, int64_t N, float *x, int64_t offset, int32_t activation) {
  TNN_CHECK_CTX(ctx);
  if (activation == TNN_ACT_LINEAR || N <= 0) return 0;
  TNN_REQUIRE(tnn_act_is_elementwise(activation), "activation is not elementwise / not implemented");
  const int a = activation;
  return launch_map<true, 0>(ctx, x + offset, 1, nullptr, 1, nullptr, 1, N,
                             [a] __device__(float d, float, float) { return tnn_act(d, a); });
}

// nncuda.pas:544-560 / cuda_sgemm.cu:512-520: output_sigmoid = logistic(x), output = x*logistic(x)
__global__ void k_swish(int64_t n, const float *__restrict__ x, float *__restrict__ sig, float *__restrict__ out) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
    float v = x[i];
    float s = 1.0f / (1.0f + expf(-v));
    sig[i] = s;
    out[i] = v * s;
  }
}
int32_t tnn_activate_array_swish(tnn_context *ctx, int64_t N, const float *x, int64_t offset, float *output_sigmoid,
                                 float *output) {
  TNN_CHECK_CTX(ctx);
  if (N <= 0) return 0;
  k_swish<<<grid_for(N, kThreads), kThreads, 0, ctx->stream>>>(N, x + offset, output_sigmoid + offset,
                                                              output + offset);
  TNN_LAUNCHED(ctx);
  return 0;
}

// nncuda.pas:562-580 / gradient_array nactivation.pas:624-717
int32_t tnn_derive_array(tnn_context *ctx, int64_t N, const float *x, int64_t offset, int32_t activation,
                         float *delta) {
  TNN_CHECK_CTX(ctx);
  if (activation == TNN_ACT_LINEAR || N <= 0) return 0;
  TNN_REQUIRE(tnn_act_is_elementwise(activation) && activation != TNN_ACT_SWISH, "gradient not implemented");
  const int a = activation;
  return launch_map<true, 1>(ctx, delta + offset, 1, x + offset, 1, nullptr, 1, N,
                             [a] __device__(float d, float o, float) { return d * tnn_grad(o, a); });
}

// nncuda.pas:582-600; CPU twin ntensors.pas:7709-7727 (incb is accepted and, like the reference kernel
// cuda_sgemm.cu:277-284, not used)
// plane size 1 (connected layers: dst[row][c] += bias[c]): the plane-map kernel would give every element its own
// plane; one pass over rows of C instead (float4 when C and the base allow)
__global__ void k_bias_rows(int64_t n, int C, float *__restrict__ dst, const float *__restrict__ bias) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads)
    dst[i] = dst[i] + __ldg(bias + (int)(i % C));
}
__global__ void k_bias_rows4(int64_t n4, int C4, float4 *__restrict__ dst, const float4 *__restrict__ bias) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += nthreads) {
    float4 v = dst[i];
    const float4 b = __ldg(bias + (int)(i % C4));
    v.x = v.x + b.x; v.y = v.y + b.y; v.z = v.z + b.z; v.w = v.w + b.w;
    dst[i] = v;
  }
}

int32_t tnn_forward_bias(tnn_context *ctx, int64_t dstSize, float *dst, int64_t offset, int64_t srcSize,
                         const float *src, int64_t incb, int64_t batch) {
  TNN_CHECK_CTX(ctx);
  (void)incb;
  if (srcSize > 0 && batch > 0 && dstSize == srcSize * batch && srcSize < (1ll << 31) && dst && src) {
    float *d = dst + offset;
    if ((srcSize & 3) == 0 && ((((uintptr_t)d) | ((uintptr_t)src)) & 15) == 0) {
      k_bias_rows4<<<grid_for(dstSize / 4, kThreads), kThreads, 0, ctx->stream>>>(dstSize / 4, (int)(srcSize / 4), (float4 *)d,
                                                                                  (const float4 *)src);
    } else {
      k_bias_rows<<<grid_for(dstSize, kThreads), kThreads, 0, ctx->stream>>>(dstSize, (int)srcSize, d, src);
    }
    TNN_LAUNCHED(ctx);
    return 0;
  }
  return planemap::launch<0, 0>(ctx, dst + offset, nullptr, nullptr, nullptr, nullptr, dstSize, srcSize, batch,
                               [src] __device__(float &v, float, float, float &, float &, int c) { v = v + __ldg(src + c); });
}

// nncuda.pas:1468-1488; CPU twin ntensors.pas:7687-7707
int32_t tnn_forward_scale(tnn_context *ctx, int64_t dstSize, float *dst, int64_t offset, int64_t scaleSize,
                          const float *scale, int64_t incb, int64_t batch) {
  TNN_CHECK_CTX(ctx);
  (void)incb;
  return planemap::launch<0, 0>(ctx, dst + offset, nullptr, nullptr, nullptr, nullptr, dstSize, scaleSize, batch,
                               [scale] __device__(float &v, float, float, float &, float &, int c) { v = v * __ldg(scale + c); });
}

// nncuda.pas:1490-1510 (separately rounded multiply and add, like forwardScale followed by forwardBias)
int32_t tnn_forward_scale_add(tnn_context *ctx, int64_t dstSize, float *dst, int64_t offset, int64_t scaleSize,
                              const float *scales, const float *biases, int64_t incb, int64_t batch) {
  TNN_CHECK_CTX(ctx);
  (void)incb;
  return planemap::launch<0, 0>(ctx, dst + offset, nullptr, nullptr, nullptr, nullptr, dstSize, scaleSize, batch,
                               [scales, biases] __device__(float &v, float, float, float &, float &, int c) {
                                 v = __fadd_rn(__fmul_rn(v, __ldg(scales + c)), __ldg(biases + c));
                               });
}

// nncuda.pas:1371-1396; CPU twin blockNormalize ntensors.pas:8693-8718 (_snormblkvv :4357-4382, _snormvv :4330-4340)
int32_t tnn_normalize(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups, const float *means,
                      int64_t meansStride, const float *vars, int64_t varsStride, float *dst, int64_t dstOffset) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(srcSize > 0 && groups > 0, "bad sizes");
  const bool block1 = dstSize == srcSize * groups;
  return planemap::launch<0, 0>(ctx, dst + dstOffset, nullptr, nullptr, nullptr, nullptr, dstSize, srcSize, groups,
                               [means, meansStride, vars, varsStride, block1] __device__(float &v, float, float, float &,
                                                                                         float &, int c) {
                                 const float var = __ldg(vars + c * varsStride);
                                 const float d = block1 ? sqrtf(fmaxf(var, TNN_EPS)) : fmaxf(sqrtf(var), TNN_EPS);
                                 v = (v - __ldg(means + c * meansStride)) / d;
                               });
}

// ---- BLAS-1 style ----------------------------------------------------------------------------------
int32_t tnn_addvv(tnn_context *ctx, int64_t N, const float *s1, int64_t o1, int64_t inca, const float *s2, int64_t o2,
                  int64_t incb, float *dst, int64_t od, int64_t incc) {
  TNN_CHECK_CTX(ctx);
  return launch_map<false, 2>(ctx, dst + od, incc, s1 + o1, inca, s2 + o2, incb, N,
                              [] __device__(float, float a, float b) { return a + b; });
}
int32_t tnn_subvv(tnn_context *ctx, int64_t N, const float *s1, int64_t o1, int64_t inca, const float *s2, int64_t o2,
                  int64_t incb, float *dst, int64_t od, int64_t incc) {
  TNN_CHECK_CTX(ctx);
  return launch_map<false, 2>(ctx, dst + od, incc, s1 + o1, inca, s2 + o2, incb, N,
                              [] __device__(float, float a, float b) { return a - b; });
}
int32_t tnn_mulvv(tnn_context *ctx, int64_t N, const float *s1, int64_t o1, int64_t inca, const float *s2, int64_t o2,
                  int64_t incb, float *dst, int64_t od, int64_t incc) {
  TNN_CHECK_CTX(ctx);
  return launch_map<false, 2>(ctx, dst + od, incc, s1 + o1, inca, s2 + o2, incb, N,
                              [] __device__(float, float a, float b) { return a * b; });
}

__global__ void k_fmavv(int64_t n, const float *a, int64_t ia, const float *b, int64_t ib, const float *c, int64_t ic,
                        float *d, int64_t id) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads)
    d[i * id] = __fadd_rn(__fmul_rn(a[i * ia], b[i * ib]), c[i * ic]);
}
int32_t tnn_fmavv(tnn_context *ctx, int64_t N, const float *s1, int64_t o1, int64_t inca, const float *s2, int64_t o2,
                  int64_t incb, const float *s3, int64_t o3, int64_t incc, float *dst, int64_t od, int64_t incd) {
  TNN_CHECK_CTX(ctx);
  if (N <= 0) return 0;
  k_fmavv<<<grid_for(N, kThreads), kThreads, 0, ctx->stream>>>(N, s1 + o1, inca, s2 + o2, incb, s3 + o3, incc,
                                                              dst + od, incd);
  TNN_LAUNCHED(ctx);
  return 0;
}

// y += a*x with separately rounded multiply and add (saxpy_pas ntensors.pas:1130-1136, the default build)
int32_t tnn_axpy(tnn_context *ctx, int64_t N, float a, const float *x, int64_t xo, int64_t incx, float *y, int64_t yo,
                 int64_t incy) {
  TNN_CHECK_CTX(ctx);
  return launch_map<true, 1>(ctx, y + yo, incy, x + xo, incx, nullptr, 1, N,
                             [a] __device__(float d, float xv, float) { return __fadd_rn(__fmul_rn(a, xv), d); });
}

int32_t tnn_power(tnn_context *ctx, int64_t N, const float *x, int64_t xo, int64_t incx, float a, float *y, int64_t yo,
                  int64_t incy) {
  TNN_CHECK_CTX(ctx);
  return launch_map<false, 1>(ctx, y + yo, incy, x + xo, incx, nullptr, 1, N,
                              [a] __device__(float, float xv, float) { return powf(xv, a); });
}

// x *= a; a == 0 stores zeros without reading (the delta-zeroing use, nnet.pas:287-295)
int32_t tnn_scale(tnn_context *ctx, int64_t N, float a, float *x, int64_t stride) {
  TNN_CHECK_CTX(ctx);
  if (a == 0.0f && stride == 1) {
    if (N <= 0) return 0;
    TNN_CUDA(cudaMemsetAsync(x, 0, (size_t)N * sizeof(float), ctx->stream));
    return 0;
  }
  return launch_map<true, 0>(ctx, x, stride, nullptr, 1, nullptr, 1, N,
                             [a] __device__(float d, float, float) { return a * d; });
}

int32_t tnn_fill(tnn_context *ctx, int64_t N, float *x, int64_t offset, float val, int64_t stride) {
  TNN_CHECK_CTX(ctx);
  return launch_map<false, 0>(ctx, x + offset, stride, nullptr, 1, nullptr, 1, N,
                              [val] __device__(float, float, float) { return val; });
}

int32_t tnn_copy(tnn_context *ctx, int64_t N, const float *src, int64_t so, int64_t inca, float *dst, int64_t dof,
                 int64_t incb) {
  TNN_CHECK_CTX(ctx);
  if (N <= 0) return 0;
  if (inca == 1 && incb == 1) {
    TNN_CUDA(cudaMemcpyAsync(dst + dof, src + so, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  return launch_map<false, 1>(ctx, dst + dof, incb, src + so, inca, nullptr, 1, N,
                              [] __device__(float, float a, float) { return a; });
}

int32_t tnn_fmavss(tnn_context *ctx, int64_t N, const float *src, int64_t offset, float scalar, float bias,
                   float *dst) {
  TNN_CHECK_CTX(ctx);
  return launch_map<false, 1>(ctx, dst + offset, 1, src + offset, 1, nullptr, 1, N, [scalar, bias] __device__(float, float a, float) {
    return __fadd_rn(__fmul_rn(a, scalar), bias);
  });
}

int32_t tnn_clamp(tnn_context *ctx, int64_t N, float alpha, const float *src, float *dst, int64_t stride,
                  int64_t offset) {
  TNN_CHECK_CTX(ctx);
  return launch_map<false, 1>(ctx, dst + offset, stride, src + offset, stride, nullptr, 1, N,
                              [alpha] __device__(float, float a, float) { return fmaxf(fminf(a, alpha), -alpha); });
}

int32_t tnn_inverse_sqrt(tnn_context *ctx, int64_t N, float alpha, const float *src, float *dst, int64_t stride,
                         int64_t offset) {
  TNN_CHECK_CTX(ctx);
  (void)alpha;
  return launch_map<false, 1>(ctx, dst + offset, stride, src + offset, stride, nullptr, 1, N,
                              [] __device__(float, float a, float) { return 1.0f / sqrtf(fmaxf(a, TNN_EPS)); });
}

// ---- losses (cuda_sgemm.cu:814-830, 954-972, 1655-1662) --------------------------------------------
__global__ void k_loss(int64_t n, const float *pred, const float *truth, float *delta, float *error, int kind) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
    const float t = truth[i], p = pred[i];
    if (kind == 0) {  // crossEntropyLogistics
      error[i] = -t * logf(fmaxf(p, TNN_EPS)) - (1 - t) * logf(fmaxf(1 - p, TNN_EPS));
      delta[i] = t - p;
    } else if (kind == 1) {  // crossEntropySoftmax
      error[i] = (t != 0) ? -logf(fmaxf(p, TNN_EPS)) : 0.f;
      delta[i] = t - p;
    } else {  // cost_l2
      const float r = t - p;
      delta[i] = r;
      error[i] = r * r;
    }
  }
}
static int launch_loss(tnn_context *ctx, int64_t N, const float *pred, const float *truth, float *delta, float *error,
                       int kind) {
  if (N <= 0) return 0;
  k_loss<<<grid_for(N, kThreads), kThreads, 0, ctx->stream>>>(N, pred, truth, delta, error, kind);
  TNN_LAUNCHED(ctx);
  return 0;
}
int32_t tnn_cross_entropy_logistic(tnn_context *ctx, int64_t N, const float *pred, const float *truth, float *delta,
                                   float *error) {
  TNN_CHECK_CTX(ctx);
  return launch_loss(ctx, N, pred, truth, delta, error, 0);
}
int32_t tnn_cross_entropy_softmax(tnn_context *ctx, int64_t N, const float *pred, const float *truth, float *delta,
                                  float *error) {
  TNN_CHECK_CTX(ctx);
  return launch_loss(ctx, N, pred, truth, delta, error, 1);
}
int32_t tnn_cost_l2(tnn_context *ctx, int64_t N, const float *pred, const float *truth, float *delta, float *error) {
  TNN_CHECK_CTX(ctx);
  return launch_loss(ctx, N, pred, truth, delta, error, 2);
}

// ---- dropout (cuda_sgemm.cu:1620-1653) ---------------------------------------------------------------
__global__ void k_dropout_fwd(int64_t n, uint64_t seed, float probability, float scale, const float *src, float *rnd,
                              float *dst) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
    unsigned long long x = seed + (unsigned long long)i;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    x *= 0x2545F4914F6CDD1DULL;
    float r = (float)(x % 0x10000000ull);
    r /= (float)0x10000000ull;
    rnd[i] = r;
    dst[i] = r < probability ? 0.f : src[i] * scale;
  }
}
__global__ void k_dropout_fwd_dev(int64_t n, const unsigned long long *seed_dev, unsigned long long off, float probability,
                                  float scale, const float *src, float *rnd, float *dst) {
  const unsigned long long seed = *seed_dev + off;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads) {
    unsigned long long x = seed + (unsigned long long)i;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    x *= 0x2545F4914F6CDD1DULL;
    float r = (float)(x % 0x10000000ull);
    r /= (float)0x10000000ull;
    rnd[i] = r;
    dst[i] = r < probability ? 0.f : src[i] * scale;
  }
}
__global__ void k_counter_add(unsigned long long *ctr, unsigned long long inc) { *ctr += inc; }
__global__ void k_dropout_bwd(int64_t n, float probability, float scale, const float *src, const float *rnd,
                              float *dst) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nthreads)
    dst[i] = rnd[i] < probability ? 0.f : src[i] * scale;
}
int32_t tnn_forward_dropout(tnn_context *ctx, int64_t N, const float *src, float probability, float scale, float *rnd,
                            float *dst, uint64_t seed) {
  TNN_CHECK_CTX(ctx);
  if (N <= 0) return 0;
  k_dropout_fwd<<<grid_for(N, kThreads), kThreads, 0, ctx->stream>>>(N, seed, probability, scale, src, rnd, dst);
  TNN_LAUNCHED(ctx);
  return 0;
}
int32_t tnn_forward_dropout_dev(tnn_context *ctx, int64_t N, const float *src, float probability, float scale,
                                float *rnd, float *dst, const uint64_t *seed_dev, uint64_t seed_offset) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(seed_dev != nullptr, "null seed pointer");
  if (N <= 0) return 0;
  k_dropout_fwd_dev<<<grid_for(N, kThreads), kThreads, 0, ctx->stream>>>(
      N, (const unsigned long long *)seed_dev, (unsigned long long)seed_offset, probability, scale, src, rnd, dst);
  TNN_LAUNCHED(ctx);
  return 0;
}
int32_t tnn_counter_add(tnn_context *ctx, uint64_t *ctr, uint64_t inc) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(ctr != nullptr, "null counter");
  k_counter_add<<<1, 1, 0, ctx->stream>>>((unsigned long long *)ctr, (unsigned long long)inc);
  TNN_LAUNCHED(ctx);
  return 0;
}
int32_t tnn_backward_dropout(tnn_context *ctx, int64_t N, const float *src, float probability, float scale,
                             const float *rnd, float *dst) {
  TNN_CHECK_CTX(ctx);
  if (N <= 0) return 0;
  k_dropout_bwd<<<grid_for(N, kThreads), kThreads, 0, ctx->stream>>>(N, probability, scale, src, rnd, dst);
  TNN_LAUNCHED(ctx);
  return 0;
}

// ---- LSTM gate update (nlstmlayer.pas:934-973: 4 addvv + 4 activate + mulvv + fmavv + copy + activate + mulvv +
// 2 copies -> 1 launch).  wz = h(t-1) W^T for the four gates [batch][4*hidden] (f | i | g | o, no bias yet),
// uz = x(t) U^T + bu, same layout.  Same order of operations per element as the operator sequence:
// z = (wz + bw) + uz; c = c*f + i*g; h = tanh(c)*o.
__global__ void k_lstm_gates(int64_t n, int hidden, const float *__restrict__ wz, const float *__restrict__ bw,
                             const float *__restrict__ uz, float *__restrict__ c, float *__restrict__ h,
                             float *__restrict__ cell_out, float *__restrict__ h_out) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += nthreads) {
    const int64_t b = e / hidden;
    const int j = (int)(e - b * hidden);
    const int64_t row = b * 4 * hidden;
    float z[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int64_t k = row + (int64_t)g * hidden + j;
      z[g] = __fadd_rn(__fadd_rn(wz[k], __ldg(bw + g * hidden + j)), uz[k]);
    }
    const float f = tnn_act(z[0], TNN_ACT_LOGISTIC), i = tnn_act(z[1], TNN_ACT_LOGISTIC);
    const float g = tnn_act(z[2], TNN_ACT_TANH), o = tnn_act(z[3], TNN_ACT_LOGISTIC);
    const float cn = __fadd_rn(__fmul_rn(c[e], f), __fmul_rn(i, g));  // as tnn_fmavv: separately rounded product and sum
    const float hn = __fmul_rn(o, tnn_act(cn, TNN_ACT_TANH));
    c[e] = cn;
    h[e] = hn;
    cell_out[e] = cn;
    h_out[e] = hn;
  }
}
int32_t tnn_lstm_gates(tnn_context *ctx, int64_t batch, int64_t hidden, const float *wz, const float *bw,
                       const float *uz, int64_t uzOffset, float *c, float *h, float *cell_out, float *h_out,
                       int64_t outOffset) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(wz && bw && uz && c && h && cell_out && h_out, "null pointer");
  const int64_t n = batch * hidden;
  if (n <= 0) return 0;
  TNN_REQUIRE(hidden < (1ll << 29), "hidden size too large");
  k_lstm_gates<<<grid_for(n, kThreads), kThreads, 0, ctx->stream>>>(n, (int)hidden, wz, bw, uz + uzOffset, c, h,
                                                                    cell_out + outOffset, h_out + outOffset);
  TNN_LAUNCHED(ctx);
  return 0;
}

// ---- LSTM step, backward (TLSTMLayer.backward, nlstmlayer.pas:661-812: per step 4 addvv + 6 activate + 7 mulvv +
// 5 gradient_array + add + 10 copies, then the eight sub-layers' delta.Clamp(-1,1) nconnectedlayer.pas:262) in ONE
// launch.  Gate pre-activations are re-formed as in the forward kernel; per element, in the reference's order:
//   t2 = dh*o ; t2 *= tanh'(tanh c) ; t2 += dc                 (dL/dc of this step)
//   d_o = tanh(c)*dh * o(1-o) ; d_g = t2*i * (1-g^2) ; d_i = t2*g * i(1-i) ; d_f = t2*c_prev * f(1-f) ; dc = t2*f
// dz[b][gate*hidden + j] (gate order f|i|g|o) receives the four pre-activation deltas clamped to [-1, 1].
__global__ void k_lstm_gates_backward(int64_t n, int hidden, const float *__restrict__ wz, const float *__restrict__ bw,
                                      const float *__restrict__ uz, const float *__restrict__ cell,
                                      const float *__restrict__ prev_cell, const float *__restrict__ dh,
                                      float *__restrict__ dc, float *__restrict__ dz) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += nthreads) {
    const int64_t b = e / hidden;
    const int j = (int)(e - b * hidden);
    const int64_t row = b * 4 * hidden;
    float z[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int64_t k = row + (int64_t)g * hidden + j;
      z[g] = __fadd_rn(__fadd_rn(wz[k], __ldg(bw + g * hidden + j)), uz[k]);
    }
    const float f = tnn_act(z[0], TNN_ACT_LOGISTIC), i = tnn_act(z[1], TNN_ACT_LOGISTIC);
    const float g = tnn_act(z[2], TNN_ACT_TANH), o = tnn_act(z[3], TNN_ACT_LOGISTIC);
    const float t3 = dh[e];
    const float th = tnn_act(cell[e], TNN_ACT_TANH);
    float t2 = __fmul_rn(t3, o);
    t2 = __fmul_rn(t2, tnn_grad(th, TNN_ACT_TANH));
    t2 = __fadd_rn(t2, dc[e]);
    const float d_o = __fmul_rn(__fmul_rn(th, t3), tnn_grad(o, TNN_ACT_LOGISTIC));
    const float d_g = __fmul_rn(__fmul_rn(t2, i), tnn_grad(g, TNN_ACT_TANH));
    const float d_i = __fmul_rn(__fmul_rn(t2, g), tnn_grad(i, TNN_ACT_LOGISTIC));
    const float cp = prev_cell ? prev_cell[e] : 0.0f;
    const float d_f = __fmul_rn(__fmul_rn(t2, cp), tnn_grad(f, TNN_ACT_LOGISTIC));
    dc[e] = __fmul_rn(t2, f);
    auto clamp1 = [](float v) { return fminf(fmaxf(v, -1.0f), 1.0f); };
    dz[row + j] = clamp1(d_f);
    dz[row + hidden + j] = clamp1(d_i);
    dz[row + 2 * (int64_t)hidden + j] = clamp1(d_g);
    dz[row + 3 * (int64_t)hidden + j] = clamp1(d_o);
  }
}

int32_t tnn_lstm_gates_backward(tnn_context *ctx, int64_t batch, int64_t hidden, const float *wz, const float *bw,
                                const float *uz, int64_t zOffset, const float *cell, int64_t cellOffset,
                                int64_t prevCellOffset, const float *dh, int64_t dhOffset, float *dc, float *dz,
                                int64_t dzOffset) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(wz && bw && uz && cell && dh && dc && dz, "null pointer");
  const int64_t n = batch * hidden;
  if (n <= 0) return 0;
  TNN_REQUIRE(hidden < (1ll << 29), "hidden size too large");
  k_lstm_gates_backward<<<grid_for(n, kThreads), kThreads, 0, ctx->stream>>>(
      n, (int)hidden, wz + zOffset, bw, uz + zOffset, cell + cellOffset, prevCellOffset >= 0 ? cell + prevCellOffset : nullptr,
      dh + dhOffset, dc, dz + dzOffset);
  TNN_LAUNCHED(ctx);
  return 0;
}

// ---- fused SGD update (nConvolutionLayer.pas:1309-1322: 3 launches -> 1; 16 B/param) -----------------
int32_t tnn_sgd_update(tnn_context *ctx, int64_t N, float *params, float *updates, float rate, float momentum,
                       float decay_term) {
  TNN_CHECK_CTX(ctx);
  if (N <= 0) return 0;
  return tnn_sgd_launch(ctx, N, params, updates, rate, momentum, decay_term);
}

int32_t tnn_sgd_update_multi(tnn_context *ctx, int32_t count, float *const *params, float *const *updates,
                             const uint64_t *sizes, const float *rates, const float *decay_terms, float momentum) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(count >= 0 && (count == 0 || (params && updates && sizes && rates && decay_terms)), "null table");
  for (int32_t base = 0; base < count; base += kSgdMulti) {
    SgdItems it;
    memset(&it, 0, sizeof(it));
    int m = 0;
    int64_t largest = 0;
    for (int32_t i = base; i < count && m < kSgdMulti; ++i) {
      if (sizes[i] == 0) continue;
      it.w[m] = params[i];
      it.u[m] = updates[i];
      it.n[m] = (int64_t)sizes[i];
      it.rate[m] = rates[i];
      it.decay[m] = decay_terms[i];
      if (it.n[m] > largest) largest = it.n[m];
      ++m;
    }
    if (!m) continue;
    dim3 grid((unsigned)grid_for(cdiv64(largest, 4), kThreads), (unsigned)m);
    k_sgd_multi<<<grid, kThreads, 0, ctx->stream>>>(it, momentum);
    TNN_LAUNCHED(ctx);
  }
  return 0;
}

}  // extern "C"

