// conv_direct.cuh — direct convolution on CUDA cores for layers whose contraction is tiny (C*kh*kw <= 64,
// e.g. the 3-channel first layer of YOLOv3 / CIFAR nets: K = 27).  Such a layer is bound by its output write
// (128 B/pixel for 32 filters) and by the fp32 FMA rate, not by the tensor cores: padding K = 27 up to MMA
// granularity and staging it through the tcgen05 pipeline cost 10x more than computing it in registers.
// Replaces the same im2col + gemm + bias + activate sequence (nConvolutionLayer.pas:1040-1122); arithmetic is
// exact fp32 FMA with k ascending, i.e. bit-identical to the CUDA-core implicit GEMM and to the reference's
// saxpy sweep in its FMA build.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"

namespace direct {

constexpr int PIX = 4;    // consecutive output pixels (along ow) per thread
constexpr int FT = 32;    // filters per CTA pass
constexpr int KMAX = 64;  // largest contraction handled here
constexpr int THREADS = 128;

static __device__ __noinline__ float act_slow(float v, int a) { return tnn_act(v, a); }

template <bool K3S1>
__global__ void __launch_bounds__(THREADS, 3)
k_conv_direct(simt::ConvGeom g, int batch, const float *__restrict__ in, const float *__restrict__ w,
              const float *__restrict__ scale, const float *__restrict__ bias, int act, float *__restrict__ out,
              int64_t outPitch) {  // outPitch != 0: channels-innermost output rows of `outPitch` floats (NHWC)
  __shared__ __align__(16) float ws[KMAX][FT];  // ws[k][f] = w[f0+f][k]
  const int K = g.C * g.kh * g.kw;
  const int f0 = blockIdx.y * FT;
  for (int i = threadIdx.x; i < K * FT; i += THREADS) {
    const int k = i / FT, f = i - k * FT;
    ws[k][f] = (f0 + f < g.F) ? __ldg(w + (int64_t)(f0 + f) * K + k) : 0.f;
  }
  __syncthreads();

  const int owq = (g.OW + PIX - 1) / PIX;
  const int64_t total = (int64_t)batch * g.OH * owq;
  const int64_t t = (int64_t)blockIdx.x * THREADS + threadIdx.x;
  if (t >= total) return;
  const int q = (int)(t % owq);
  const int64_t r = t / owq;
  const int oh = (int)(r % g.OH);
  const int b = (int)(r / g.OH);
  const int ow0 = q * PIX;

  float acc[PIX][FT];
#pragma unroll
  for (int p = 0; p < PIX; ++p)
#pragma unroll
    for (int f = 0; f < FT; ++f) acc[p][f] = 0.f;

  const float *img = in + (int64_t)b * g.inImg;
  int k = 0;
  if (K3S1) {
    // 3x3, stride 1, dilation 1: the PIX outputs of a row share a window of PIX+2 input columns per (c, kr) —
    // 6 loads instead of 12, taps unrolled.  Same products in the same k order as the general loop below.
    const int iw0 = ow0 - g.pw;
    for (int c = 0; c < g.C; ++c) {
      const float *plane = img + (int64_t)c * g.H * g.W;
#pragma unroll 1  // the fully unrolled body (27 taps x 128 FMAs) overflows the instruction cache: ncu showed
                  // stalled_no_instruction as the top stall reason
      for (int kr = 0; kr < 3; ++kr) {
        const int ih = oh - g.ph + kr;
        const bool rowok = (unsigned)ih < (unsigned)g.H;
        const float *rowp = plane + (int64_t)ih * g.W + iw0;
        float xw[PIX + 2];
#pragma unroll
        for (int j = 0; j < PIX + 2; ++j) xw[j] = (rowok && (unsigned)(iw0 + j) < (unsigned)g.W) ? __ldg(rowp + j) : 0.f;
#pragma unroll
        for (int kc = 0; kc < 3; ++kc, ++k) {
#pragma unroll
          for (int f = 0; f < FT; f += 4) {
            const float4 wv = *reinterpret_cast<const float4 *>(&ws[k][f]);
#pragma unroll
            for (int p = 0; p < PIX; ++p) {
              acc[p][f + 0] = fmaf(xw[p + kc], wv.x, acc[p][f + 0]);
              acc[p][f + 1] = fmaf(xw[p + kc], wv.y, acc[p][f + 1]);
              acc[p][f + 2] = fmaf(xw[p + kc], wv.z, acc[p][f + 2]);
              acc[p][f + 3] = fmaf(xw[p + kc], wv.w, acc[p][f + 3]);
            }
          }
        }
      }
    }
  } else
  for (int c = 0; c < g.C; ++c) {
    const float *plane = img + (int64_t)c * g.H * g.W;
    for (int kr = 0; kr < g.kh; ++kr) {
      const int ih = oh * g.sy - g.ph + kr * g.dy;
      const bool rowok = (unsigned)ih < (unsigned)g.H;
      const float *rowp = plane + (int64_t)ih * g.W;
      for (int kc = 0; kc < g.kw; ++kc, ++k) {
        float x[PIX];
#pragma unroll
        for (int p = 0; p < PIX; ++p) {
          const int iw = (ow0 + p) * g.sx - g.pw + kc * g.dx;
          x[p] = (rowok && (unsigned)iw < (unsigned)g.W) ? __ldg(rowp + iw) : 0.f;
        }
#pragma unroll
        for (int f = 0; f < FT; f += 4) {
          const float4 wv = *reinterpret_cast<const float4 *>(&ws[k][f]);
#pragma unroll
          for (int p = 0; p < PIX; ++p) {
            acc[p][f + 0] = fmaf(x[p], wv.x, acc[p][f + 0]);
            acc[p][f + 1] = fmaf(x[p], wv.y, acc[p][f + 1]);
            acc[p][f + 2] = fmaf(x[p], wv.z, acc[p][f + 2]);
            acc[p][f + 3] = fmaf(x[p], wv.w, acc[p][f + 3]);
          }
        }
      }
    }
  }

  if (outPitch) {
    // channels-innermost rows: the thread's PIX pixels are consecutive rows, each 32 filters = 128 contiguous bytes
    const bool leaky = act == TNN_ACT_LEAKY || act == TNN_ACT_REVLEAKY;
    const bool simple = leaky || act == TNN_ACT_LINEAR;
    float *opix = out + (((int64_t)b * g.OH + oh) * g.OW + ow0) * outPitch + f0;
    const bool vec4 = (f0 + FT <= g.F) && ((outPitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
    for (int f = 0; f < FT; f += 4) {
      float s[4], bi[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[u] = (scale && f0 + f + u < g.F) ? __ldg(scale + f0 + f + u) : 1.0f;
        bi[u] = (bias && f0 + f + u < g.F) ? __ldg(bias + f0 + f + u) : 0.0f;
      }
#pragma unroll
      for (int p = 0; p < PIX; ++p) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float y = acc[p][f + u];
          if (scale) y = __fmul_rn(y, s[u]);
          if (bias) y = __fadd_rn(y, bi[u]);
          v[u] = simple ? (leaky ? (y < 0 ? 0.1f * y : y) : y) : act_slow(y, act);
        }
        if (ow0 + p < g.OW) {
          float *o = opix + (int64_t)p * outPitch + f;
          if (vec4) {
            *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (f0 + f + u < g.F) o[u] = v[u];
          }
        }
      }
    }
    return;
  }
  const int P = g.OH * g.OW;
  float *orow = out + (int64_t)b * g.outImg + (int64_t)oh * g.OW + ow0;
  const bool vec = (ow0 + PIX <= g.OW) && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) && ((P & 3) == 0);
  const bool leaky = act == TNN_ACT_LEAKY || act == TNN_ACT_REVLEAKY;
  const bool simple = leaky || act == TNN_ACT_LINEAR;
#pragma unroll
  for (int f = 0; f < FT; ++f) {
    if (f0 + f < g.F) {
      const float s = scale ? __ldg(scale + f0 + f) : 1.0f;
      const float bi = bias ? __ldg(bias + f0 + f) : 0.0f;
      float v[PIX];
#pragma unroll
      for (int p = 0; p < PIX; ++p) {
        float y = acc[p][f];
        if (scale) y = __fmul_rn(y, s);
        if (bias) y = __fadd_rn(y, bi);
        v[p] = simple ? (leaky ? (y < 0 ? 0.1f * y : y) : y) : act_slow(y, act);
      }
      float *o = orow + (int64_t)(f0 + f) * P;
      if (vec) {
        *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int p = 0; p < PIX; ++p)
          if (ow0 + p < g.OW) o[p] = v[p];
      }
    }
  }
}

inline bool supported(const simt::ConvGeom &g) { return (int64_t)g.C * g.kh * g.kw <= KMAX; }

inline int conv_forward(tnn_context *ctx, const simt::ConvGeom &g, int batch, const float *in, const float *weights,
                        const float *scale, const float *bias, int act, float *out, int64_t outPitch = 0) {
  const int owq = (g.OW + PIX - 1) / PIX;
  const int64_t total = (int64_t)batch * g.OH * owq;
  dim3 grid((unsigned)cdiv64(total, THREADS), (unsigned)cdiv64(g.F, FT));
  TNN_REQUIRE(grid.y <= 65535, "too many filter tiles");
  if (g.kh == 3 && g.kw == 3 && g.sy == 1 && g.sx == 1 && g.dy == 1 && g.dx == 1)
    k_conv_direct<true><<<grid, THREADS, 0, ctx->stream>>>(g, batch, in, weights, scale, bias, act, out, outPitch);
  else
    k_conv_direct<false><<<grid, THREADS, 0, ctx->stream>>>(g, batch, in, weights, scale, bias, act, out, outPitch);
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // namespace direct
