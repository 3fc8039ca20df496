// common.cuh — context, error handling and small device helpers shared by all kernels of
// libtensorium_b200.so (the sm_100a replacement for the reference's nncuda.pas backend).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tensorium_b200.h"

struct tnn_context {
  int device;
  cudaStream_t stream;
  bool own_stream;
  cudaDeviceProp prop;
  int math_mode;
  int tma_activations;  // 1: stride-1 convolutions fetch activations by TMA when the geometry allows (default)
  int tma_gemm;         // 1: aligned SGEMM operands are fetched by TMA (default; TNN_TMA_GEMM=0 turns it off)
  int gather_kb16;      // 1: the gathering 3xTF32 convolution runs 16-k pipeline stages (TNN_GATHER_KB16=1; slower, kept for A/B runs)
  int patch_kb16;       // 1: TMA-activation convolutions run 16-channel pipeline stages in 3xTF32 mode (TNN_PATCH_KB16=1; no measurable gain, opt-in)
  int tail_split;       // 1: a last wave that fills <= half of the SMs runs as half-width tiles (default; TNN_TAIL_SPLIT=0)
  int nhwc_pair;        // 1: the NHWC convolution runs wide tiles on CTA pairs (tcgen05.mma.cta_group::2); TNN_NHWC_PAIR=0 turns it off
  int nhwc_kb;          // channels per pipeline stage of the NHWC convolution in 3xTF32 mode: 0 = default (16), 16 or 32 (TNN_NHWC_KB)
  int nhwc_bn_auto;     // 1: with several filter tiles per pixel tile, the tile width (256 / 224 / 192 / 160) that minimises the wave count x tile cost (default; TNN_NHWC_BN_AUTO=0: always 256)
  int pdl;              // 1: the channels-innermost convolutions launch with programmatic stream serialization (their prologue overlaps the previous kernel's tail); TNN_PDL=0 off
  int nhwc_epi_sets;    // 0: automatic; 1 / 2: sets of epilogue warps in the CTA-pair channels-innermost kernel (TNN_NHWC_EPI_SETS)
  int nhwc_image;       // 1: the image layer of a channels-innermost graph runs on the tensor cores (BF16x3 / TF32 modes); TNN_NHWC_IMAGE=0: direct kernel
  int last_conv_kernel; // which kernel the last tnn_conv2d_forward took (TNN_CONV_KERNEL_*), for measurement tools
  int64_t launches;
  int capturing;              // between tnn_graph_begin_capture and tnn_graph_end_capture
  int64_t capture_launches0;
  // scratch for two-stage reductions (partials + arrival counters), grown on demand
  float *red_partials;
  size_t red_partials_bytes;
  unsigned int *red_counters;
  size_t red_counters_n;
  float *gemm_scratch;          // operand copies of the SGEMM that runs on the channels-innermost convolution kernel (conv_nhwc.cu)
  size_t gemm_scratch_bytes;
  int gemm_nhwc;                // 1: large SGEMMs run as 1x1 convolutions on the CTA-pair kernel (default; TNN_GEMM_NHWC=0)
  // scratch for the contraction kernels (pre-split / transposed weight copies, split-K partials)
  float *work;
  size_t work_bytes;
  // prepared copies of weight tensors the caller declared immutable (tnn_set_weights_static)
  struct tnn_wcache *wcache;
  // instantiated CUDA graphs bake the scratch pointers into their kernel nodes: while any graph is alive a scratch
  // buffer that must grow is RETIRED (kept allocated until tnn_destroy) instead of freed
  int live_graphs;
  void *retired[64];
  int n_retired;
};

// Buffer for the prepared (re-ordered, TF32-split) copy of `weights`: hi plane then lo plane, `count` floats each.
// *fresh tells the caller to run its prepare kernel; it is false when `weights` is static and a copy in this layout
// (`kind`, math mode, count) is already cached.
int tnn_weight_buffer(tnn_context *ctx, const float *weights, int kind, size_t count, float **hi, float **lo, bool *fresh);

void tnn_set_error(const char *fmt, ...);

// fp32-class accuracy modes split every operand in two (3xTF32: TF32 hi/lo, 3 MMAs; BF16x3: bf16 hi/lo, 3 MMAs)
static inline bool tnn_math_is_split(const tnn_context *ctx) {
  return ctx->math_mode == TNN_MATH_3XTF32 || ctx->math_mode == TNN_MATH_BF16X3;
}

#define TNN_CHECK_CTX(ctx)                         \
  do {                                             \
    if (!(ctx)) {                                  \
      tnn_set_error("%s: null context", __func__); \
      return 1;                                    \
    }                                              \
  } while (0)

#define TNN_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      tnn_set_error("%s:%d %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define TNN_LAUNCHED(ctx)                                                                  \
  do {                                                                                     \
    (ctx)->launches++;                                                                     \
    cudaError_t e__ = cudaPeekAtLastError();                                               \
    if (e__ != cudaSuccess) {                                                              \
      tnn_set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,                  \
                    cudaGetErrorString(cudaGetLastError()));                               \
      return 3;                                                                            \
    }                                                                                      \
  } while (0)

#define TNN_REQUIRE(cond, msg)                               \
  do {                                                       \
    if (!(cond)) {                                           \
      tnn_set_error("%s: %s (%s)", __func__, msg, #cond);    \
      return 4;                                              \
    }                                                        \
  } while (0)

int tnn_reserve_reduction(tnn_context *ctx, size_t partial_floats, size_t counters);
int tnn_reserve_work(tnn_context *ctx, size_t bytes);
// free a scratch buffer, or park it until tnn_destroy when a live CUDA graph may still reference it
int tnn_release_scratch(tnn_context *ctx, void *p);

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- device helpers
#define TNN_EPS 0.000001f  // sEPSILON, ntensors.pas:95

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Scalar activation, following the CPU functions of nactivation.pas:272-401 (not the NVRTC kernels
// where they differ).  Decimal-constant multiplies are done in fp64 and rounded once, like the
// Pascal expressions (real-typed literal * single), so the piecewise-linear ones are bit-exact.
__device__ __forceinline__ float tnn_act(float x, int a) {
  switch (a) {
    case TNN_ACT_LOGISTIC: return 1.0f / (1.0f + expf(-x));
    case TNN_ACT_RELU: return x * (float)(x > 0);
    case TNN_ACT_RELU6: return x * (float)(x > 0) * (float)(x <= 6);  // 0 above 6: reference quirk (nactivation.pas:316)
    case TNN_ACT_RELIE: return x > 0 ? x : (float)(0.01 * (double)x);
    case TNN_ACT_RAMP: return (float)((double)x * (double)(x > 0) + 0.1 * (double)x);
    case TNN_ACT_TANH: { float px = expf(x), nx = expf(-x); return (px - nx) / (px + nx); }
    case TNN_ACT_PLSE:
      if (x < -4) return (float)(0.01 * ((double)x + 4));
      if (x > 4) return (float)(0.01 * ((double)x - 4) + 1);
      return (float)(0.125 * (double)x + 0.5);
    case TNN_ACT_REVLEAKY:
    case TNN_ACT_LEAKY: return x < 0 ? (float)(0.1 * (double)x) : x;
    case TNN_ACT_ELU: return (x >= 0) ? x : (expf(x) - 1.0f);
    case TNN_ACT_LOGGY: return 2.0f / (1.0f + expf(-x)) - 1.0f;
    case TNN_ACT_STAIR: {
      long long n = (long long)floorf(x);
      if (n % 2 == 0) return floorf(x / 2);
      return (float)(((double)x - (double)n) + floor((double)x / 2));
    }
    case TNN_ACT_HARDTAN: return x < -1 ? -1.0f : (x > 1 ? 1.0f : x);
    case TNN_ACT_LHTAN:
      if (x < 0) return (float)(0.001 * (double)x);
      if (x > 1) return (float)(0.001 * ((double)x - 1) + 1);
      return x;
    case TNN_ACT_SELU: return (x >= 0) ? (float)(1.0507 * (double)x) : (float)(1.0507 * 1.6732 * ((double)expf(x) - 1.0));
    case TNN_ACT_GELU: return (float)(0.5 * (double)x * (1.0 + tanh(0.797885 * (double)x + 0.035677 * (double)x * (double)x * (double)x)));
    case TNN_ACT_SWISH: return x * (1.0f / (1.0f + expf(-x)));
    default: return x;  // LINEAR and the non-elementwise types
  }
}

// Epilogue variant for the fused conv kernels: no fp64 (F2F/DMUL run at a small fraction of the fp32 rate and
// stalled the tcgen05 epilogue warps); differs from tnn_act by at most 1 ulp on the decimal-constant branches.
__device__ __forceinline__ float tnn_act_fast(float x, int a) {
  switch (a) {
    case TNN_ACT_LINEAR: return x;
    case TNN_ACT_REVLEAKY:
    case TNN_ACT_LEAKY: return x < 0 ? 0.1f * x : x;
    case TNN_ACT_RELU: return x > 0 ? x : 0.f * x;
    case TNN_ACT_LOGISTIC: return 1.0f / (1.0f + __expf(-x));
    default: return tnn_act(x, a);
  }
}

// Gradient as a function of the layer OUTPUT (nactivation.pas:403-501)
__device__ __forceinline__ float tnn_grad(float x, int a) {
  switch (a) {
    case TNN_ACT_LOGISTIC: return (float)((1.0 - (double)x) * (double)x);
    case TNN_ACT_RELU: return (float)(x > 0);
    case TNN_ACT_RELU6: return (float)((x > 0) && (x < 6));
    case TNN_ACT_RELIE: return x > 0 ? 1.0f : 0.01f;
    case TNN_ACT_RAMP: return (float)((double)(x > 0) + 0.1);
    case TNN_ACT_TANH: return (float)(1.0 - (double)x * (double)x);
    case TNN_ACT_PLSE: return (x < 0 || x > 1) ? 0.01f : 0.125f;
    case TNN_ACT_REVLEAKY:
    case TNN_ACT_LEAKY: return x > 0 ? 1.0f : 0.1f;
    case TNN_ACT_ELU: return (x >= 0) ? 1.0f : (x + 1);
    case TNN_ACT_LOGGY: { double y = ((double)x + 1.0) / 2.0; return (float)(2 * (1 - y) * y); }
    case TNN_ACT_STAIR: return floorf(x) == x ? 0.0f : 1.0f;
    case TNN_ACT_HARDTAN: return (x > -1 && x < 1) ? 1.0f : 0.0f;
    case TNN_ACT_LHTAN: return (x > 0 && x < 1) ? 1.0f : 0.001f;
    case TNN_ACT_SELU: return (x >= 0) ? 1.0507f : (float)((double)x + 1.0507 * 1.6732);
    case TNN_ACT_GELU: {
      double xd = x, x3 = xd * xd * xd, t = 0.0356774 * x3 + 0.797885 * xd;
      double sech = 2.0 / (exp(t) + exp(-t));
      return (float)(0.5 * tanh(t) + (0.0535161 * x3 + 0.398942 * xd) * sech * sech + 0.5);
    }
    default: return 1.0f;
  }
}

static inline bool tnn_act_is_elementwise(int a) {
  return a >= 0 && a <= TNN_ACT_SWISH && a != TNN_ACT_SOFTMAX;
}
