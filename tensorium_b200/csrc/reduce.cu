// reduce.cu — per-channel reductions and the batch-norm passes built on them.
// Replaces kernels 8, 33, 42-44 and the batchNormGPU / batchNormBackGPU sequences of the reference
// (source/cuda_sgemm.cu:725-780,1446-1497,1690-1800; source/nbaselayer.pas:550-714).
// Reference launches one 512-thread CTA per channel with a serial thread-0 fold; here each channel is
// split over enough CTAs to fill 148 SMs, partials are combined in a fixed order by the last CTA
// to arrive (deterministic, no float atomics).
#include "common.cuh"
#include "plane_map.cuh"

namespace {

constexpr int kRT = 256;

struct Split {
  int items;          // CTAs per channel
  int partsPerPlane;  // >1: one plane is cut in parts
  int planesPerItem;  // >1: several planes (images) per CTA
  int64_t partLen;    // elements of a plane handled by one part (multiple of 4)
};

Split make_split(int64_t channels, int64_t groups, int64_t blockSize) {
  Split s;
  int64_t want = cdiv64(148 * 4, channels);
  int64_t by_work = cdiv64(groups * blockSize, 2048);
  if (want > by_work) want = by_work;
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  if (groups >= want) {
    s.partsPerPlane = 1;
    s.planesPerItem = (int)cdiv64(groups, want);
    s.items = (int)cdiv64(groups, s.planesPerItem);
  } else {
    s.planesPerItem = 1;
    s.partsPerPlane = (int)cdiv64(want, groups);
    s.items = (int)(groups * s.partsPerPlane);
  }
  s.partLen = ((cdiv64(blockSize, s.partsPerPlane) + 3) / 4) * 4;
  return s;
}

// Acc: void acc(int64_t idx, int c, float (&a)[NACC]);   Fin: void fin(int c, const float (&s)[NACC])
template <int NACC, typename Acc, typename Fin>
__global__ void __launch_bounds__(kRT) k_channel_reduce(int64_t channels, int64_t groups, int64_t blockSize, Split sp,
                                                        float *partials, unsigned int *counters, Acc acc, Fin fin) {
  const int item = blockIdx.x;
  const int c = blockIdx.y;
  const int pg = item / sp.partsPerPlane;
  const int part = item - pg * sp.partsPerPlane;
  const int64_t b0 = (int64_t)pg * sp.planesPerItem;
  const int64_t b1 = min(groups, b0 + sp.planesPerItem);
  const int64_t r0 = (int64_t)part * sp.partLen;
  const int64_t r1 = min(blockSize, r0 + sp.partLen);
  float a[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) a[k] = 0.f;
  for (int64_t b = b0; b < b1; ++b) {
    const int64_t base = (b * channels + c) * blockSize;
    int64_t r = r0 + threadIdx.x;
    // four independent elements per trip: the loads of one trip are in flight together
    for (; r + 3 * kRT < r1; r += 4 * kRT) {
      acc(base + r, c, a);
      acc(base + r + kRT, c, a);
      acc(base + r + 2 * kRT, c, a);
      acc(base + r + 3 * kRT, c, a);
    }
    for (; r < r1; r += kRT) acc(base + r, c, a);
  }
  __shared__ float sm[NACC][kRT / 32];
  __shared__ bool is_last;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    float v = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      float v = 0.f;
      for (int w = 0; w < kRT / 32; ++w) v += sm[k][w];
      tot[k] = v;
    }
    if (sp.items == 1) {
      fin(c, tot);
      is_last = false;
    } else {
      float *p = partials + ((int64_t)c * sp.items + item) * NACC;
#pragma unroll
      for (int k = 0; k < NACC; ++k) p[k] = tot[k];
      __threadfence();
      const unsigned int prev = atomicAdd(&counters[c], 1u);
      is_last = (prev == (unsigned int)sp.items - 1);
      if (is_last) {
        __threadfence();
        float s[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) s[k] = 0.f;
        const volatile float *q = partials + (int64_t)c * sp.items * NACC;
        for (int i = 0; i < sp.items; ++i)
#pragma unroll
          for (int k = 0; k < NACC; ++k) s[k] += q[i * NACC + k];
        fin(c, s);
        counters[c] = 0;
      }
    }
  }
}

template <int NACC, typename Acc, typename Fin>
int launch_reduce(tnn_context *ctx, int64_t channels, int64_t groups, int64_t blockSize, Acc acc, Fin fin) {
  if (channels <= 0) return 0;
  TNN_REQUIRE(groups > 0 && blockSize > 0, "tensor sizes do not align");
  TNN_REQUIRE(channels <= 65535, "too many channels for one launch");
  Split sp = make_split(channels, groups, blockSize);
  int rc = tnn_reserve_reduction(ctx, (size_t)channels * sp.items * NACC, (size_t)channels);
  if (rc) return rc;
  dim3 grid(sp.items, (unsigned)channels);
  k_channel_reduce<NACC, Acc, Fin><<<grid, kRT, 0, ctx->stream>>>(channels, groups, blockSize, sp, ctx->red_partials,
                                                                 ctx->red_counters, acc, fin);
  TNN_LAUNCHED(ctx);
  return 0;
}

// The same reduction with the inputs handed over as pointers: when every plane is 16-byte aligned the kernel reads them
// with 128-bit loads, two independent quads per input and trip in flight (BASELINE north star: "128-bit vectorised
// coalesced loads with warp-shuffle reductions").  The index-functor kernel above reads 4 bytes per load and thread:
// ncu on a 268 MB tensor showed the batch-norm backward sums at 14 % of the DRAM peak.
// AccV: void acc(const float (&v)[NIN], int c, float (&a)[NACC])     v[j] = in[j][element]
template <int NIN>
struct Inputs {
  const float *p[NIN];
};
template <int NACC, int NIN, bool VEC, typename AccV, typename Fin>
__global__ void __launch_bounds__(kRT) k_channel_reduce_v(int64_t channels, int64_t groups, int64_t blockSize, Split sp,
                                                          float *partials, unsigned int *counters, Inputs<NIN> in, AccV acc,
                                                          Fin fin) {
  const int item = blockIdx.x;
  const int c = blockIdx.y;
  const int pg = item / sp.partsPerPlane;
  const int part = item - pg * sp.partsPerPlane;
  const int64_t b0 = (int64_t)pg * sp.planesPerItem;
  const int64_t b1 = min(groups, b0 + sp.planesPerItem);
  const int64_t r0 = (int64_t)part * sp.partLen;
  const int64_t r1 = min(blockSize, r0 + sp.partLen);
  float a[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) a[k] = 0.f;
  auto one = [&](const float4 (&q)[NIN]) {
    float v[NIN];
#pragma unroll
    for (int j = 0; j < NIN; ++j) v[j] = q[j].x;
    acc(v, c, a);
#pragma unroll
    for (int j = 0; j < NIN; ++j) v[j] = q[j].y;
    acc(v, c, a);
#pragma unroll
    for (int j = 0; j < NIN; ++j) v[j] = q[j].z;
    acc(v, c, a);
#pragma unroll
    for (int j = 0; j < NIN; ++j) v[j] = q[j].w;
    acc(v, c, a);
  };
  if (VEC) {   // r0, r1 - r0 and every plane base are multiples of 4 (checked by the launcher)
    // The item's planes form ONE index space of quads (plane, quad in plane): with small planes (8 x 8 = 16 quads) a
    // plane-by-plane walk kept 16 of the 256 threads busy and paid one memory latency per plane
    // (128 channels x 128 images x 64: 27 us for a 4 MB tensor).
    const int qpp = (int)((r1 - r0) >> 2);
    const int64_t nq = (b1 - b0) * qpp;
    auto at = [&](int64_t q) {
      const int64_t pl = q / qpp;
      return ((b0 + pl) * channels + c) * blockSize + r0 + ((q - pl * qpp) << 2);
    };
    int64_t q = threadIdx.x;
    for (; q + kRT < nq; q += 2 * kRT) {
      const int64_t e0 = at(q), e1 = at(q + kRT);
      float4 q0[NIN], q1[NIN];
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        q0[j] = __ldg(reinterpret_cast<const float4 *>(in.p[j] + e0));
        q1[j] = __ldg(reinterpret_cast<const float4 *>(in.p[j] + e1));
      }
      one(q0);
      one(q1);
    }
    for (; q < nq; q += kRT) {
      const int64_t e0 = at(q);
      float4 q0[NIN];
#pragma unroll
      for (int j = 0; j < NIN; ++j) q0[j] = __ldg(reinterpret_cast<const float4 *>(in.p[j] + e0));
      one(q0);
    }
  } else {
    for (int64_t b = b0; b < b1; ++b) {
      const int64_t base = (b * channels + c) * blockSize;
      for (int64_t r = r0 + threadIdx.x; r < r1; r += kRT) {
        float v[NIN];
#pragma unroll
        for (int j = 0; j < NIN; ++j) v[j] = __ldg(in.p[j] + base + r);
        acc(v, c, a);
      }
    }
  }
  __shared__ float sm[NACC][kRT / 32];
  __shared__ bool is_last;
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    float v = warp_sum(a[k]);
    if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      float v = 0.f;
      for (int w = 0; w < kRT / 32; ++w) v += sm[k][w];
      tot[k] = v;
    }
    if (sp.items == 1) {
      fin(c, tot);
      is_last = false;
    } else {
      float *p = partials + ((int64_t)c * sp.items + item) * NACC;
#pragma unroll
      for (int k = 0; k < NACC; ++k) p[k] = tot[k];
      __threadfence();
      const unsigned int prev = atomicAdd(&counters[c], 1u);
      is_last = (prev == (unsigned int)sp.items - 1);
      if (is_last) {
        __threadfence();
        float s[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) s[k] = 0.f;
        const volatile float *q = partials + (int64_t)c * sp.items * NACC;
        for (int i = 0; i < sp.items; ++i)
#pragma unroll
          for (int k = 0; k < NACC; ++k) s[k] += q[i * NACC + k];
        fin(c, s);
        counters[c] = 0;
      }
    }
  }
}

template <int NACC, int NIN, typename AccV, typename Fin>
int launch_reduce_v(tnn_context *ctx, int64_t channels, int64_t groups, int64_t blockSize, Inputs<NIN> in, AccV acc, Fin fin) {
  if (channels <= 0) return 0;
  TNN_REQUIRE(groups > 0 && blockSize > 0, "tensor sizes do not align");
  TNN_REQUIRE(channels <= 65535, "too many channels for one launch");
  Split sp = make_split(channels, groups, blockSize);
  int rc = tnn_reserve_reduction(ctx, (size_t)channels * sp.items * NACC, (size_t)channels);
  if (rc) return rc;
  bool vec = (blockSize % 4) == 0;
  for (int j = 0; j < NIN; ++j) vec = vec && (((uintptr_t)in.p[j]) & 15) == 0;
  dim3 grid(sp.items, (unsigned)channels);
  if (vec)
    k_channel_reduce_v<NACC, NIN, true, AccV, Fin><<<grid, kRT, 0, ctx->stream>>>(channels, groups, blockSize, sp, ctx->red_partials,
                                                                                 ctx->red_counters, in, acc, fin);
  else
    k_channel_reduce_v<NACC, NIN, false, AccV, Fin><<<grid, kRT, 0, ctx->stream>>>(channels, groups, blockSize, sp, ctx->red_partials,
                                                                                  ctx->red_counters, in, acc, fin);
  TNN_LAUNCHED(ctx);
  return 0;
}

inline int64_t block_size(int64_t total, int64_t channels, int64_t groups) {
  if (channels <= 0 || groups <= 0) return 0;
  int64_t bs = total / (channels * groups);
  return (bs * channels * groups == total) ? bs : 0;
}

}  // namespace

// kA: the activation as a compile-time constant (the gradient switch folds away), -1 = decided per element at run time
template <int kA>
static int bn_backward_impl(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize, const float *out,
                            int activation, float *delta, const float *x, const float *x_norm, const float *mean,
                            const float *variance, const float *scales, float *scale_updates, float *mean_delta,
                            float *variance_delta) {
  const float *dl = delta;
  const float batchSize = (float)(blockSize * batch);
  const int a = activation;
  int rc = launch_reduce_v<3, 4>(
      ctx, channels, batch, blockSize, Inputs<4>{{dl, out, x, x_norm}},
      [mean, a] __device__(const float(&v)[4], int c, float(&s)[3]) {
        const int act = kA >= 0 ? kA : a;
        const float g = (act == TNN_ACT_LINEAR) ? v[0] : v[0] * tnn_grad(v[1], act);
        s[0] = fmaf(v[3], g, s[0]);
        s[1] += g;
        s[2] = fmaf(g, v[2] - __ldg(mean + c), s[2]);
      },
      [scales, scale_updates, variance, mean_delta, variance_delta] __device__(int c, const float(&t)[3]) {
        scale_updates[c] = scale_updates[c] + t[0];
        const double v = (double)fmaxf(variance[c], TNN_EPS);
        const float sc = scales[c];
        mean_delta[c] = (float)((double)(t[1] * sc) * (-1.0 / sqrt(v)));
        variance_delta[c] = (float)((double)(t[2] * sc) * -0.5 * pow(v, -1.5));
      });
  if (rc) return rc;
  return planemap::launch<2, 0>(
      ctx, delta, out, x, nullptr, nullptr, batch * channels * blockSize, channels, batch,
      [mean, variance, mean_delta, variance_delta, scales, batchSize, a] __device__(float &d, float ov, float xv, float &,
                                                                                    float &, int c) {
        const int act = kA >= 0 ? kA : a;
        float g = (act == TNN_ACT_LINEAR) ? d : d * tnn_grad(ov, act);
        g = g * __ldg(scales + c);
        const float md = __ldg(mean_delta + c) / batchSize;
        const float vd = 2 * __ldg(variance_delta + c) / batchSize;
        const float sd = sqrtf(fmaxf(__ldg(variance + c), TNN_EPS));
        d = __fadd_rn(g / sd, __fadd_rn(__fmul_rn(vd, xv - __ldg(mean + c)), md));
      });
}

template <int kA>
static int bn_forward_training_impl(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize, float *out, float *x,
                                    float *x_norm, float *mean, float *variance, float *rolling_mean,
                                    float *rolling_variance, float bn_momentum, const float *scales, const float *biases,
                                    int32_t activation) {
  const int64_t total = batch * channels * blockSize;
  // Statistics in ONE read: sums of d = x - s and d*d, with s = the channel's first element as the shift (so the
  // subtraction in  var = (sum d^2 - (sum d)^2/N)/(N-1)  cancels little), finished in double; the rolling averages
  // are updated by the same finaliser.  (The reference reads the tensor twice: MeansAndVars ntensors.pas:9102-9177.)
  const float *src = out;
  const double S = (double)(batch * blockSize);
  const float m = bn_momentum;
  int rc = launch_reduce_v<2, 1>(
      ctx, channels, batch, blockSize, Inputs<1>{{src}},
      [src, blockSize] __device__(const float(&v)[1], int c, float(&acc)[2]) {
        const float d = v[0] - __ldg(src + (int64_t)c * blockSize);
        acc[0] += d;
        acc[1] = fmaf(d, d, acc[1]);
      },
      [src, blockSize, mean, variance, rolling_mean, rolling_variance, S, m] __device__(int c, const float(&t)[2]) {
        const double sd = (double)t[0], sq = (double)t[1];
        const float mu = (float)((double)src[(int64_t)c * blockSize] + sd / S);
        double var = (sq - sd * sd / S) / (S - 1.0);
        const float v = (float)(var > 0.0 ? var : 0.0);
        mean[c] = mu;
        variance[c] = v;
        rolling_mean[c] = __fadd_rn(__fmul_rn(m, mu), __fmul_rn(1 - m, rolling_mean[c]));
        rolling_variance[c] = __fadd_rn(__fmul_rn(m, v), __fmul_rn(1 - m, rolling_variance[c]));
      });
  if (rc) return rc;
  const int a = activation;
  const bool block1 = blockSize == 1;
  return planemap::launch<0, 2>(
      ctx, out, nullptr, nullptr, x, x_norm, total, channels, batch,
      [mean, variance, scales, biases, a, block1] __device__(float &v, float, float, float &xo, float &xn, int c) {
        const float var = __ldg(variance + c);
        const float d = block1 ? sqrtf(fmaxf(var, TNN_EPS)) : fmaxf(sqrtf(var), TNN_EPS);
        xo = v;
        float t = (v - __ldg(mean + c)) / d;
        xn = t;
        t = __fmul_rn(t, __ldg(scales + c));
        t = __fadd_rn(t, __ldg(biases + c));
        v = tnn_act(t, kA >= 0 ? kA : a);
      });
}

extern "C" {

// block size 1 (connected layers: src is [batch][channels]): a column sum — 32 channels x 8 batch slices per CTA, reads
// coalesced along the channels, slices folded in a fixed order (deterministic)
__global__ void __launch_bounds__(256) k_column_sum(const float *__restrict__ src, int64_t channels, int64_t batch,
                                                    float *__restrict__ dst) {
  __shared__ float part[8][33];
  const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + cx;
  float a = 0.f;
  if (c < channels)
    for (int64_t b = sl; b < batch; b += 8) a += src[b * channels + c];
  part[sl][cx] = a;
  __syncthreads();
  if (sl == 0 && c < channels) {
    float t = part[0][cx];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += part[i][cx];
    dst[c] = dst[c] + t;
  }
}

// backwardBias nncuda.pas:602-622; CPU twin addSums ntensors.pas:7729-7781: dst[f] += sum_{b,block} src
int32_t tnn_backward_bias(tnn_context *ctx, int64_t dstSize, float *dst, int64_t srcSize, const float *src,
                          int64_t srcOffset, int64_t incb, int64_t batch) {
  TNN_CHECK_CTX(ctx);
  (void)incb;
  const int64_t bs = block_size(srcSize, dstSize, batch);
  TNN_REQUIRE(bs > 0, "tensor sizes do not align");
  const float *s = src + srcOffset;
  if (bs == 1) {
    k_column_sum<<<(unsigned)((dstSize + 31) / 32), 256, 0, ctx->stream>>>(s, dstSize, batch, dst);
    TNN_LAUNCHED(ctx);
    return 0;
  }
  return launch_reduce<1>(
      ctx, dstSize, batch, bs, [s] __device__(int64_t i, int, float(&a)[1]) { a[0] += s[i]; },
      [dst] __device__(int c, const float(&t)[1]) { dst[c] = dst[c] + t[0]; });
}

// addDots nncuda.pas:1448-1466; CPU twin ntensors.pas:7783-7837
int32_t tnn_add_dots(tnn_context *ctx, int64_t N, int64_t dstSize, int64_t groups, const float *src1,
                     const float *src2, int64_t srcOffset, float *dst) {
  TNN_CHECK_CTX(ctx);
  const int64_t bs = block_size(N, dstSize, groups);
  TNN_REQUIRE(bs > 0, "tensor sizes do not align");
  const float *p = src1 + srcOffset, *q = src2 + srcOffset;
  return launch_reduce<1>(
      ctx, dstSize, groups, bs, [p, q] __device__(int64_t i, int, float(&a)[1]) { a[0] = fmaf(p[i], q[i], a[0]); },
      [dst] __device__(int c, const float(&t)[1]) { dst[c] = dst[c] + t[0]; });
}

// means nncuda.pas:1330-1348; CPU twin MeansAndVars ntensors.pas:9102-9177 (first pass)
int32_t tnn_means(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups, const float *src,
                  int64_t offset, float *means) {
  TNN_CHECK_CTX(ctx);
  const int64_t bs = block_size(srcSize, dstSize, groups);
  TNN_REQUIRE(bs > 0, "tensor sizes do not align");
  const float *s = src + offset;
  const float S = (float)(groups * bs);
  return launch_reduce<1>(
      ctx, dstSize, groups, bs, [s] __device__(int64_t i, int, float(&a)[1]) { a[0] += s[i]; },
      [means, S] __device__(int c, const float(&t)[1]) { means[c] = t[0] / S; });
}

// variances nncuda.pas:1350-1369 — UNBIASED (/(S-1)) like the CPU (ntensors.pas:9136,9165-9166), second pass
int32_t tnn_variances(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups, const float *src,
                      int64_t offset, const float *means, float *vars) {
  TNN_CHECK_CTX(ctx);
  const int64_t bs = block_size(srcSize, dstSize, groups);
  TNN_REQUIRE(bs > 0, "tensor sizes do not align");
  const float *s = src + offset;
  const float S2 = (float)(groups * bs - 1);
  return launch_reduce<1>(
      ctx, dstSize, groups, bs,
      [s, means] __device__(int64_t i, int c, float(&a)[1]) {
        const float d = s[i] - __ldg(means + c);
        a[0] = fmaf(d, d, a[0]);
      },
      [vars, S2] __device__(int c, const float(&t)[1]) { vars[c] = t[0] / S2; });
}

// meansAndVars nncuda.pas:1302-1328
int32_t tnn_means_and_vars(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups, const float *src,
                           int64_t offset, float *means, float *vars) {
  int rc = tnn_means(ctx, srcSize, dstSize, groups, src, offset, means);
  if (rc) return rc;
  return tnn_variances(ctx, srcSize, dstSize, groups, src, offset, means, vars);
}

// meansAndVarsDelta nncuda.pas:1398-1419; CPU twin sMeanAndVarianceDelta ntensors.pas:8831-8899
int32_t tnn_means_and_vars_delta(tnn_context *ctx, int64_t srcSize, int64_t dstSize, int64_t groups,
                                 const float *delta, const float *x, int64_t offset, const float *mean,
                                 const float *variance, float *mean_delta, float *variance_delta) {
  TNN_CHECK_CTX(ctx);
  const int64_t bs = block_size(srcSize, dstSize, groups);
  TNN_REQUIRE(bs > 0, "tensor sizes do not align");
  const float *d = delta + offset, *xx = x + offset;
  return launch_reduce<2>(
      ctx, dstSize, groups, bs,
      [d, xx, mean] __device__(int64_t i, int c, float(&a)[2]) {
        const float dv = d[i];
        a[0] += dv;
        a[1] = fmaf(dv, xx[i] - __ldg(mean + c), a[1]);
      },
      [variance, mean_delta, variance_delta] __device__(int c, const float(&t)[2]) {
        const double v = (double)fmaxf(variance[c], TNN_EPS);
        mean_delta[c] = (float)((double)t[0] * (-1.0 / sqrt(v)));
        variance_delta[c] = (float)((double)t[1] * -0.5 * pow(v, -1.5));
      });
}

// normalizeDelta nncuda.pas:1421-1446; CPU twin sNormalizeDelta ntensors.pas:8902-8951
int32_t tnn_normalize_delta(tnn_context *ctx, int64_t deltaSize, int64_t meanSize, int64_t groups, float *delta,
                            const float *x, int64_t offset, const float *mean, const float *variance,
                            const float *mean_delta, const float *variance_delta) {
  TNN_CHECK_CTX(ctx);
  const int64_t bs = block_size(deltaSize, meanSize, groups);
  TNN_REQUIRE(bs > 0, "tensor sizes do not align");
  const float batchSize = (float)(bs * groups);
  return planemap::launch<1, 0>(
      ctx, delta + offset, x + offset, nullptr, nullptr, nullptr, deltaSize, meanSize, groups,
      [mean, variance, mean_delta, variance_delta, batchSize] __device__(float &d, float xv, float, float &, float &,
                                                                         int c) {
        const float md = __ldg(mean_delta + c) / batchSize;
        const float vd = 2 * __ldg(variance_delta + c) / batchSize;
        const float sd = sqrtf(fmaxf(__ldg(variance + c), TNN_EPS));
        d = __fadd_rn(d / sd, __fadd_rn(__fmul_rn(vd, xv - __ldg(mean + c)), md));
      });
}

// ---- fused batch-norm entry points ---------------------------------------------------------------------------

// inference batchNormGPU (nbaselayer.pas:628-643: normalize, forwardScale, forwardBias) + ActivateArray: 1 pass
int32_t tnn_batchnorm_forward_inference(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize,
                                        float *out, const float *rolling_mean, const float *rolling_variance,
                                        const float *scales, const float *biases, int32_t activation) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(activation == TNN_ACT_LINEAR || tnn_act_is_elementwise(activation), "activation not elementwise");
  const int a = activation;
  const bool block1 = blockSize == 1;
  return planemap::launch<0, 0>(
      ctx, out, nullptr, nullptr, nullptr, nullptr, batch * channels * blockSize, channels, batch,
      [rolling_mean, rolling_variance, scales, biases, a, block1] __device__(float &v, float, float, float &, float &,
                                                                             int c) {
        const float var = __ldg(rolling_variance + c);
        const float d = block1 ? sqrtf(fmaxf(var, TNN_EPS)) : fmaxf(sqrtf(var), TNN_EPS);
        float t = (v - __ldg(rolling_mean + c)) / d;
        t = __fmul_rn(t, __ldg(scales + c));
        t = __fadd_rn(t, __ldg(biases + c));
        v = tnn_act(t, a);
      });
}

// TYoloLayer.forwardGPU activation loop (nyololayer.pas:1081-1099) as one plane map
int32_t tnn_yolo_activate(tnn_context *ctx, int64_t batch, int64_t n, int64_t classes, int64_t wh, float *out,
                          float scale_xy) {
  TNN_CHECK_CTX(ctx);
  const int entries = (int)(classes + 5);
  const float a = scale_xy, b = 0.5f * (1.0f - scale_xy);
  return planemap::launch<0, 0>(ctx, out, nullptr, nullptr, nullptr, nullptr, batch * n * entries * wh, n * entries,
                                batch, [entries, a, b] __device__(float &v, float, float, float &, float &, int c) {
                                  const int e = c % entries;
                                  if (e == 2 || e == 3) return;
                                  float t = 1.0f / (1.0f + expf(-v));
                                  if (e < 2 && a != 1.0f) t = __fadd_rn(__fmul_rn(t, a), b);
                                  v = t;
                                });
}

// training batchNormGPU (nbaselayer.pas:583-643) + ActivateArray: 1 reduction read + 1 read / 3 writes; the rolling
// averages follow nbaselayer.pas:590-608 (scale then axpy, separately rounded)
int32_t tnn_batchnorm_forward_training(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize,
                                       float *out, float *x, float *x_norm, float *mean, float *variance,
                                       float *rolling_mean, float *rolling_variance, float bn_momentum,
                                       const float *scales, const float *biases, int32_t activation) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(activation == TNN_ACT_LINEAR || tnn_act_is_elementwise(activation), "activation not elementwise");
  switch (activation) {   // the common activations as compile-time constants: the per-element switch folds away
#define TNN_BNF(A) bn_forward_training_impl<A>(ctx, batch, channels, blockSize, out, x, x_norm, mean, variance, rolling_mean, \
                                            rolling_variance, bn_momentum, scales, biases, activation)
    case TNN_ACT_LINEAR: return TNN_BNF(TNN_ACT_LINEAR);
    case TNN_ACT_RELU: return TNN_BNF(TNN_ACT_RELU);
    case TNN_ACT_LEAKY: return TNN_BNF(TNN_ACT_LEAKY);
    default: return TNN_BNF(-1);
#undef TNN_BNF
  }
}

// DeriveArray + batchNormBackGPU (nConvolutionLayer.pas:1186; nbaselayer.pas:662-714) in two passes.
// With g = delta*f'(out):  scale_updates += sum x_norm*g ;  d = g*scale ;
// mean_delta = sum(d) * -1/sqrt(max(var,eps)) ; variance_delta = sum d*(x-mean) * -0.5*max(var,eps)^-1.5 ;
// delta = d/sqrt(max(var,eps)) + (2*variance_delta*(x-mean) + mean_delta)/(batch*blockSize)
int32_t tnn_batchnorm_backward(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize,
                               const float *out, int32_t activation, float *delta, const float *x,
                               const float *x_norm, const float *mean, const float *variance, const float *scales,
                               float *scale_updates, float *mean_delta, float *variance_delta) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(activation == TNN_ACT_LINEAR || (tnn_act_is_elementwise(activation) && activation != TNN_ACT_SWISH),
              "activation gradient not implemented");
  switch (activation) {
#define TNN_BNB(A) bn_backward_impl<A>(ctx, batch, channels, blockSize, out, activation, delta, x, x_norm, mean, variance, \
                                    scales, scale_updates, mean_delta, variance_delta)
    case TNN_ACT_LINEAR: return TNN_BNB(TNN_ACT_LINEAR);
    case TNN_ACT_RELU: return TNN_BNB(TNN_ACT_RELU);
    case TNN_ACT_LEAKY: return TNN_BNB(TNN_ACT_LEAKY);
    default: return TNN_BNB(-1);
#undef TNN_BNB
  }
}

// ---- batch-norm under batch sharding (SyncBN).  The reference computes the statistics over the whole mini-batch
// (nbaselayer.pas:352, MeansAndVars ntensors.pas:9102-9177); with the batch split over replicas the per-channel sums
// are exchanged (tnn_allreduce_sum on the context's stream) so that every replica normalises with the statistics of
// the GLOBAL batch: the reference's two-pass formula — mean = sum(x)/S, variance = sum((x-mean)^2)/(S-1) with
// S = globalBatch*blockSize — and, backward, the two sums mean_delta / variance_delta are built from.
// comm == NULL (or one rank) degenerates to the single-device two-pass computation.
int32_t tnn_batchnorm_forward_training_sync(tnn_context *ctx, tnn_comm *comm, int64_t globalBatch, int64_t batch,
                                            int64_t channels, int64_t blockSize, float *out, float *x, float *x_norm,
                                            float *mean, float *variance, float *rolling_mean, float *rolling_variance,
                                            float bn_momentum, const float *scales, const float *biases,
                                            int32_t activation) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(activation == TNN_ACT_LINEAR || tnn_act_is_elementwise(activation), "activation not elementwise");
  TNN_REQUIRE(globalBatch >= batch, "the global batch is the sum of the replicas' batches");
  const int64_t total = batch * channels * blockSize;
  const float *src = out;
  const double S = (double)(globalBatch * blockSize);
  // pass 1: local sum(x) -> mean[] ; exchanged ; mean = sum / S
  int rc = launch_reduce<1>(
      ctx, channels, batch, blockSize, [src] __device__(int64_t i, int, float(&acc)[1]) { acc[0] += src[i]; },
      [mean] __device__(int c, const float(&t)[1]) { mean[c] = t[0]; });
  if (rc) return rc;
  if (comm) {
    rc = tnn_allreduce_sum(ctx, comm, mean, 0, channels);
    if (rc) return rc;
  }
  rc = tnn_scale(ctx, channels, (float)(1.0 / S), mean, 1);
  if (rc) return rc;
  // pass 2: local sum((x - mean)^2) -> variance[] ; exchanged ; variance = sum / (S-1) ; rolling averages
  rc = launch_reduce<1>(
      ctx, channels, batch, blockSize,
      [src, mean] __device__(int64_t i, int c, float(&acc)[1]) {
        const float d = src[i] - __ldg(mean + c);
        acc[0] = fmaf(d, d, acc[0]);
      },
      [variance] __device__(int c, const float(&t)[1]) { variance[c] = t[0]; });
  if (rc) return rc;
  if (comm) {
    rc = tnn_allreduce_sum(ctx, comm, variance, 0, channels);
    if (rc) return rc;
  }
  const float m = bn_momentum;
  const float inv = (float)(1.0 / (S - 1.0));
  rc = planemap::launch<0, 0>(
      ctx, variance, nullptr, nullptr, nullptr, nullptr, channels, channels, 1,
      [mean, rolling_mean, rolling_variance, m, inv] __device__(float &v, float, float, float &, float &, int c) {
        v = __fmul_rn(v, inv);
        rolling_mean[c] = __fadd_rn(__fmul_rn(m, mean[c]), __fmul_rn(1 - m, rolling_mean[c]));
        rolling_variance[c] = __fadd_rn(__fmul_rn(m, v), __fmul_rn(1 - m, rolling_variance[c]));
      });
  if (rc) return rc;
  const int a = activation;
  const bool block1 = blockSize == 1;
  return planemap::launch<0, 2>(
      ctx, out, nullptr, nullptr, x, x_norm, total, channels, batch,
      [mean, variance, scales, biases, a, block1] __device__(float &v, float, float, float &xo, float &xn, int c) {
        const float var = __ldg(variance + c);
        const float d = block1 ? sqrtf(fmaxf(var, TNN_EPS)) : fmaxf(sqrtf(var), TNN_EPS);
        xo = v;
        float t = (v - __ldg(mean + c)) / d;
        xn = t;
        t = __fmul_rn(t, __ldg(scales + c));
        t = __fadd_rn(t, __ldg(biases + c));
        v = tnn_act(t, a);
      });
}

int32_t tnn_batchnorm_backward_sync(tnn_context *ctx, tnn_comm *comm, int64_t globalBatch, int64_t batch,
                                    int64_t channels, int64_t blockSize, const float *out, int32_t activation,
                                    float *delta, const float *x, const float *x_norm, const float *mean,
                                    const float *variance, const float *scales, float *scale_updates, float *mean_delta,
                                    float *variance_delta) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(activation == TNN_ACT_LINEAR || (tnn_act_is_elementwise(activation) && activation != TNN_ACT_SWISH),
              "activation gradient not implemented");
  TNN_REQUIRE(globalBatch >= batch, "the global batch is the sum of the replicas' batches");
  const int a = activation;
  const float *dl = delta;
  // local sums; scale_updates keeps the LOCAL sum (the gradient bucket sums it over the replicas later), the two
  // sums behind mean_delta / variance_delta are exchanged now because every replica's delta needs the global ones
  int rc = launch_reduce<3>(
      ctx, channels, batch, blockSize,
      [dl, out, x, x_norm, mean, a] __device__(int64_t i, int c, float(&s)[3]) {
        const float g = (a == TNN_ACT_LINEAR) ? dl[i] : dl[i] * tnn_grad(out[i], a);
        s[0] = fmaf(x_norm[i], g, s[0]);
        s[1] += g;
        s[2] = fmaf(g, x[i] - __ldg(mean + c), s[2]);
      },
      [scale_updates, mean_delta, variance_delta] __device__(int c, const float(&t)[3]) {
        scale_updates[c] = scale_updates[c] + t[0];
        mean_delta[c] = t[1];
        variance_delta[c] = t[2];
      });
  if (rc) return rc;
  if (comm) {
    rc = tnn_allreduce_sum(ctx, comm, mean_delta, 0, channels);
    if (rc) return rc;
    rc = tnn_allreduce_sum(ctx, comm, variance_delta, 0, channels);
    if (rc) return rc;
  }
  const float batchSize = (float)(blockSize * globalBatch);
  // delta = d/sqrt(max(var,eps)) + (2*variance_delta*(x-mean) + mean_delta)/(globalBatch*blockSize), the per-channel
  // factors of sMeanAndVarianceDelta (ntensors.pas:8831-8899) applied to the exchanged sums on the fly
  return planemap::launch<2, 0>(
      ctx, delta, out, x, nullptr, nullptr, batch * channels * blockSize, channels, batch,
      [mean, variance, mean_delta, variance_delta, scales, batchSize, a] __device__(float &d, float ov, float xv,
                                                                                    float &, float &, int c) {
        float g = (a == TNN_ACT_LINEAR) ? d : d * tnn_grad(ov, a);
        const float sc = __ldg(scales + c);
        g = g * sc;
        const double v = (double)fmaxf(__ldg(variance + c), TNN_EPS);
        const float mdelta = (float)((double)(__ldg(mean_delta + c) * sc) * (-1.0 / sqrt(v)));
        const float vdelta = (float)((double)(__ldg(variance_delta + c) * sc) * -0.5 * pow(v, -1.5));
        const float md = mdelta / batchSize;
        const float vd = 2 * vdelta / batchSize;
        const float sd = sqrtf((float)v);
        d = __fadd_rn(g / sd, __fadd_rn(__fmul_rn(vd, xv - __ldg(mean + c)), md));
      });
}

}  // extern "C"
