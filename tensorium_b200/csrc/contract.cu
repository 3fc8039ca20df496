// contract.cu — SGEMM and convolution entry points: shape checks, kernel selection, epilogues.
// Replaces TNNCuda.gemm / gemmBatched / gemmStridedBatched (source/nncuda.pas:624-895) and the
// im2col -> gemm -> bias -> activate sequences of TConvolutionalLayer.forwardGPU / backwardGPU
// (source/nConvolutionLayer.pas:1022-1292).
#include "common.cuh"
#include "gemm_simt.cuh"
#include "umma_conv.cuh"
#include "conv_direct.cuh"
#include "umma_halo.cuh"
#include "umma_wgrad.cuh"
#include "umma_gemm.cuh"

namespace {

struct GemmEpi {  // C = alpha*acc + beta*C   (cblas_sgemm ntensors.pas:2231-2286)
  float *C;
  int64_t ldc, zstride;
  float alpha, beta;
  __device__ void operator()(int z, int m, int n, float acc) const {
    float *p = C + (int64_t)z * zstride + (int64_t)m * ldc + n;
    *p = (beta == 0.f) ? alpha * acc : fmaf(alpha, acc, beta * *p);
  }
};
struct GemmAtomicEpi {  // split-k partial: C += alpha*acc  (beta == 1 only)
  float *C;
  int64_t ldc;
  float alpha;
  __device__ void operator()(int, int m, int n, float acc) const { atomicAdd(C + (int64_t)m * ldc + n, alpha * acc); }
};

struct ConvFwdEpi {  // out[b][f][p] = act(acc*scale[f] + bias[f])
  float *out;
  const float *scale, *bias;
  int P;
  int64_t outImg;
  int act;
  __device__ void operator()(int, int f, int n, float acc) const {
    const int b = n / P, p = n - b * P;
    float v = acc;
    if (scale) v = __fmul_rn(v, __ldg(scale + f));
    if (bias) v = __fadd_rn(v, __ldg(bias + f));
    out[(int64_t)b * outImg + (int64_t)f * P + p] = tnn_act(v, act);
  }
};
struct WgradEpi {  // weight_updates[f][n] += acc
  float *wu;
  int64_t ld;
  bool atomic;
  __device__ void operator()(int, int f, int n, float acc) const {
    float *p = wu + (int64_t)f * ld + n;
    if (atomic) atomicAdd(p, acc); else *p = *p + acc;
  }
};
struct DgradEpi {  // in_delta[b][c][ih][iw] += acc
  float *dx;
  int HW;
  int64_t inImg;
  __device__ void operator()(int, int c, int n, float acc) const {
    const int b = n / HW, p = n - b * HW;
    float *q = dx + (int64_t)b * inImg + (int64_t)c * HW + p;
    *q = *q + acc;
  }
};

int fill_geom(const tnn_conv_desc *d, simt::ConvGeom &g) {
  TNN_REQUIRE(d != nullptr, "null conv descriptor");
  TNN_REQUIRE(d->batch > 0 && d->c > 0 && d->h > 0 && d->w > 0 && d->filters > 0 && d->kh > 0 && d->kw > 0,
              "non-positive conv dimension");
  TNN_REQUIRE(d->stride_y > 0 && d->stride_x > 0 && d->dilation_y > 0 && d->dilation_x > 0 && d->pad_h >= 0 &&
                  d->pad_w >= 0, "bad stride/dilation/padding");
  const int64_t OH = (d->h + 2 * d->pad_h - (d->dilation_y * (d->kh - 1) + 1)) / d->stride_y + 1;
  const int64_t OW = (d->w + 2 * d->pad_w - (d->dilation_x * (d->kw - 1) + 1)) / d->stride_x + 1;
  TNN_REQUIRE(OH > 0 && OW > 0, "empty convolution output");
  TNN_REQUIRE(d->batch * OH * OW < (1ll << 31) && d->batch * d->h * d->w < (1ll << 31) &&
                  d->c * d->kh * d->kw < (1ll << 31), "convolution too large for 32-bit pixel indices");
  g.C = (int)d->c; g.H = (int)d->h; g.W = (int)d->w; g.F = (int)d->filters;
  g.kh = (int)d->kh; g.kw = (int)d->kw; g.ph = (int)d->pad_h; g.pw = (int)d->pad_w;
  g.sy = (int)d->stride_y; g.sx = (int)d->stride_x; g.dy = (int)d->dilation_y; g.dx = (int)d->dilation_x;
  g.OH = (int)OH; g.OW = (int)OW;
  g.inImg = d->c * d->h * d->w;
  g.outImg = d->filters * OH * OW;
  return 0;
}

int pick_splits(tnn_context *ctx, int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = cdiv64(M, simt::BM) * cdiv64(N, simt::BN);
  const int64_t sms = ctx->prop.multiProcessorCount;
  if (tiles >= 2 * sms || K < 2048) return 1;
  int64_t s = (2 * sms) / tiles;
  const int64_t maxs = K / 1024;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return (int)s;
}

int gemm_simt_dispatch(tnn_context *ctx, bool tA, bool tB, int64_t M, int64_t N, int64_t K, float alpha,
                       const float *A, int64_t lda, int64_t sA, const float *B, int64_t ldb, int64_t sB, float beta,
                       float *C, int64_t ldc, int64_t sC, int batch) {
  simt::MatLoader la{A, lda, sA, !tA};
  simt::MatLoader lb{B, ldb, sB, tB};
  GemmEpi epi{C, ldc, sC, alpha, beta};
  if (!tA && !tB) return simt::launch<true, false>(ctx, M, N, K, batch, 1, la, lb, epi);
  if (!tA && tB) return simt::launch<true, true>(ctx, M, N, K, batch, 1, la, lb, epi);
  if (tA && !tB) return simt::launch<false, false>(ctx, M, N, K, batch, 1, la, lb, epi);
  return simt::launch<false, true>(ctx, M, N, K, batch, 1, la, lb, epi);
}

// strideC == 0 && beta == 1: C += alpha * sum_b op(A_b) op(B_b), the batch folded into k (+ split-k)
int gemm_simt_fold(tnn_context *ctx, bool tA, bool tB, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                   int64_t lda, int64_t sA, const float *B, int64_t ldb, int64_t sB, float *C, int64_t ldc,
                   int64_t batch) {
  simt::MatLoaderFoldK la{A, lda, sA, K, !tA};
  simt::MatLoaderFoldK lb{B, ldb, sB, K, tB};
  const int64_t Kt = K * batch;
  const int splits = pick_splits(ctx, M, N, Kt);
  if (splits > 1) {
    GemmAtomicEpi epi{C, ldc, alpha};
    if (!tA && !tB) return simt::launch<true, false>(ctx, M, N, Kt, 1, splits, la, lb, epi);
    if (!tA && tB) return simt::launch<true, true>(ctx, M, N, Kt, 1, splits, la, lb, epi);
    if (tA && !tB) return simt::launch<false, false>(ctx, M, N, Kt, 1, splits, la, lb, epi);
    return simt::launch<false, true>(ctx, M, N, Kt, 1, splits, la, lb, epi);
  }
  GemmEpi epi{C, ldc, 0, alpha, 1.0f};
  if (!tA && !tB) return simt::launch<true, false>(ctx, M, N, Kt, 1, 1, la, lb, epi);
  if (!tA && tB) return simt::launch<true, true>(ctx, M, N, Kt, 1, 1, la, lb, epi);
  if (tA && !tB) return simt::launch<false, false>(ctx, M, N, Kt, 1, 1, la, lb, epi);
  return simt::launch<false, true>(ctx, M, N, Kt, 1, 1, la, lb, epi);
}

}  // namespace

int tnn_gemm_on_conv(tnn_context *ctx, bool tA, bool tB, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                     int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc);   // conv_nhwc.cu

extern "C" {

int32_t tnn_gemm_strided_batched(tnn_context *ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                                 float alpha, const float *A, int64_t aOffset, int64_t lda, int64_t strideA,
                                 const float *B, int64_t bOffset, int64_t ldb, int64_t strideB, float beta, float *C,
                                 int64_t cOffset, int64_t ldc, int64_t strideC, int64_t batchCount) {
  TNN_CHECK_CTX(ctx);
  if (M <= 0 || N <= 0 || batchCount <= 0) return 0;
  TNN_REQUIRE(K >= 0 && lda > 0 && ldb > 0 && ldc >= N, "bad leading dimension");
  TNN_REQUIRE(batchCount <= 65535, "batch too large");
  A += aOffset; B += bOffset; C += cOffset;
  const bool tA = transA != 0, tB = transB != 0;
  if (batchCount > 1 && strideC == 0) {
    if (beta == 1.0f) {
      if (ctx->math_mode != TNN_MATH_FP32 && umma::gemm_supported(ctx, M, N, K * batchCount)) {
        int rc = umma::gemm(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, 1.0f, C, ldc, 0,
                            (int)batchCount, /*foldBatchIntoK=*/true);
        if (rc != umma::kNotTaken) return rc;
      }
      return gemm_simt_fold(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, batchCount);
    }
    // overlapping outputs with beta != 1: replay the reference's sequential host loop (nncuda.pas:882-887)
    for (int64_t b = 0; b < batchCount; ++b) {
      int rc = tnn_gemm(ctx, transA, transB, M, N, K, alpha, A + b * strideA, 0, lda, B + b * strideB, 0, ldb, beta, C,
                        0, ldc);
      if (rc) return rc;
    }
    return 0;
  }
  if (ctx->math_mode != TNN_MATH_FP32 && batchCount == 1) {
    // large products: a 1x1 convolution on the CTA-pair kernel of conv_nhwc.cu
    int rc = tnn_gemm_on_conv(ctx, tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    if (rc != umma::kNotTaken) return rc;
  }
  if (ctx->math_mode != TNN_MATH_FP32 && umma::gemm_supported(ctx, M, N, K)) {
    int rc = ctx->tma_gemm ? umma::gemm_tma(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,
                                            strideC, (int)batchCount)
                           : umma::kNotTaken;
    if (rc != umma::kNotTaken) return rc;
    rc = umma::gemm(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC,
                        (int)batchCount, false);
    if (rc != umma::kNotTaken) return rc;
  }
  return gemm_simt_dispatch(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC,
                            (int)batchCount);
}

int32_t tnn_gemm(tnn_context *ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K, float alpha,
                 const float *A, int64_t aOffset, int64_t lda, const float *B, int64_t bOffset, int64_t ldb, float beta,
                 float *C, int64_t cOffset, int64_t ldc) {
  return tnn_gemm_strided_batched(ctx, transA, transB, M, N, K, alpha, A, aOffset, lda, 0, B, bOffset, ldb, 0, beta, C,
                                  cOffset, ldc, 0, 1);
}

int32_t tnn_gemm_batched(tnn_context *ctx, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                         float alpha, const float *const *A, int64_t aOffset, int64_t lda, const float *const *B,
                         int64_t bOffset, int64_t ldb, float beta, float *const *C, int64_t cOffset, int64_t ldc,
                         int64_t batchCount) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(A && B && C, "null pointer array");
  for (int64_t b = 0; b < batchCount; ++b) {
    int rc = tnn_gemm(ctx, transA, transB, M, N, K, alpha, A[b], aOffset, lda, B[b], bOffset, ldb, beta, C[b],
                      cOffset, ldc);
    if (rc) return rc;
  }
  return 0;
}

int32_t tnn_conv2d_forward_residual(tnn_context *ctx, const tnn_conv_desc *d, const float *in, int64_t inOffset,
                                    const float *weights, const float *scale, const float *bias, int32_t activation,
                                    const float *residual, int64_t residualOffset, float *out, int64_t outOffset) {
  TNN_CHECK_CTX(ctx);
  simt::ConvGeom g;
  int rc = fill_geom(d, g);
  if (rc) return rc;
  TNN_REQUIRE(activation == TNN_ACT_LINEAR || tnn_act_is_elementwise(activation), "activation not elementwise");
  in += inOffset;
  out += outOffset;
  if (residual) residual += residualOffset;
  const int64_t K = (int64_t)g.C * g.kh * g.kw;
  const int64_t Npix = d->batch * g.OH * g.OW;
  // tiny contractions (C*kh*kw <= 64) with few channels (the first layer), or any tiny contraction in FP32 mode:
  // exact-fp32 direct convolution in registers
  if (direct::supported(g) && (g.C < 16 || ctx->math_mode == TNN_MATH_FP32)) {
    ctx->last_conv_kernel = 1;
    rc = direct::conv_forward(ctx, g, (int)d->batch, in, weights, scale, bias, activation, out);
  } else {
    rc = umma::kNotTaken;
    if (ctx->math_mode != TNN_MATH_FP32) {
      rc = umma::conv_forward(ctx, g, (int)d->batch, in, weights, scale, bias, activation, out, residual);
      if (rc != umma::kNotTaken) return rc;  // the tensor-core epilogues add the residual themselves
    }
    ctx->last_conv_kernel = 5;
    simt::MatLoader la{weights, K, 0, true};
    simt::Im2colCols lb{in, g};
    ConvFwdEpi epi{out, scale, bias, g.OH * g.OW, g.outImg, activation};
    rc = simt::launch<true, false>(ctx, g.F, Npix, K, 1, 1, la, lb, epi);
  }
  if (rc || !residual) return rc;
  return tnn_addvv(ctx, d->batch * g.outImg, out, 0, 1, residual, 0, 1, out, 0, 1);
}

int32_t tnn_last_conv_kernel(tnn_context *ctx) { return ctx ? ctx->last_conv_kernel : 0; }

int32_t tnn_conv2d_forward(tnn_context *ctx, const tnn_conv_desc *d, const float *in, int64_t inOffset,
                           const float *weights, const float *scale, const float *bias, int32_t activation,
                           float *out, int64_t outOffset) {
  return tnn_conv2d_forward_residual(ctx, d, in, inOffset, weights, scale, bias, activation, nullptr, 0, out, outOffset);
}

int32_t tnn_conv2d_backward_filter(tnn_context *ctx, const tnn_conv_desc *d, const float *in, int64_t inOffset,
                                   const float *delta, int64_t deltaOffset, float *weight_updates) {
  TNN_CHECK_CTX(ctx);
  simt::ConvGeom g;
  int rc = fill_geom(d, g);
  if (rc) return rc;
  in += inOffset;
  delta += deltaOffset;
  const int64_t N = (int64_t)g.C * g.kh * g.kw;
  const int64_t Kt = d->batch * (int64_t)g.OH * g.OW;
  if (ctx->math_mode != TNN_MATH_FP32) {
    if (ctx->tma_activations) {  // (the same A/B switch as the forward kernels: 0 keeps the round-1 gathering GEMM)
      rc = umma::conv_wgrad_tiles(ctx, g, (int)d->batch, in, delta, weight_updates);
      if (rc != umma::kNotTaken) return rc;
    }
    rc = umma::conv_backward_filter(ctx, g, (int)d->batch, in, delta, weight_updates);
    if (rc != umma::kNotTaken) return rc;
  }
  simt::DeltaRowsFoldK la{delta, g};
  simt::Im2colRowsFoldK lb{in, g};
  const int splits = pick_splits(ctx, g.F, N, Kt);
  WgradEpi epi{weight_updates, N, splits > 1};
  return simt::launch<true, true>(ctx, g.F, N, Kt, 1, splits, la, lb, epi);
}

int32_t tnn_conv2d_backward_data(tnn_context *ctx, const tnn_conv_desc *d, const float *weights, const float *delta,
                                 int64_t deltaOffset, float *in_delta, int64_t inDeltaOffset) {
  TNN_CHECK_CTX(ctx);
  simt::ConvGeom g;
  int rc = fill_geom(d, g);
  if (rc) return rc;
  delta += deltaOffset;
  in_delta += inDeltaOffset;
  const int64_t K = (int64_t)g.F * g.kh * g.kw;
  const int64_t Npix = d->batch * (int64_t)g.H * g.W;
  if (ctx->math_mode != TNN_MATH_FP32) {
    rc = umma::conv_backward_data(ctx, g, (int)d->batch, weights, delta, in_delta);
    if (rc != umma::kNotTaken) return rc;
  }
  simt::WeightsByChannel la{weights, g};
  simt::DeltaGatherCols lb{delta, g};
  DgradEpi epi{in_delta, g.H * g.W, g.inImg};
  return simt::launch<true, false>(ctx, g.C, Npix, K, 1, 1, la, lb, epi);
}

}  // extern "C"
