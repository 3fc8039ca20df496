// gemm_simt.cuh — CUDA-core (FFMA) tiled contraction used for TNN_MATH_FP32 and for shapes the tcgen05
// kernels do not take.  One code path serves plain SGEMM and the three implicit-GEMM convolutions:
// operands are described by loader functors (row context + element-at-k), results leave through an
// epilogue functor, so no im2col / col2im buffer exists.
// k is accumulated in ascending order with FMA, i.e. exactly the order of the reference's saxpy sweep
// (s_nn, ntensors.pas:2093-2099) in its AVX2/FMA build.
#pragma once
#include "common.cuh"

namespace simt {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

// Loader concept:
//   struct Row;                                   per-row (or per-column) context
//   __device__ Row  row(int z, int mn) const;     mn already bounds-checked (< M or < N)
//   __device__ float at(const Row&, int64_t k) const;   k < K
// Epilogue concept:  __device__ void operator()(int z, int m, int n, float acc) const;
template <bool KFAST_A, bool KFAST_B, typename LA, typename LB, typename EPI>
__global__ void __launch_bounds__(THREADS)
k_gemm(int M, int N, int64_t K, int64_t kChunk, int zPerSplit, LA la, LB lb, EPI epi) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // blockIdx.z enumerates (batch z, k-split s): z = blockIdx.z / zPerSplit... (zPerSplit = number of k-splits)
  const int z = blockIdx.z / zPerSplit;
  const int split = blockIdx.z - z * zPerSplit;
  const int64_t kBegin = (int64_t)split * kChunk;
  const int64_t kEnd = min(K, kBegin + kChunk);

  typename LA::Row ra[KFAST_A ? 4 : 1];
  bool va[KFAST_A ? 4 : 1];
  typename LB::Row rb[KFAST_B ? 4 : 1];
  bool vb[KFAST_B ? 4 : 1];
  if (KFAST_A) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + (t >> 4) + 16 * i;
      va[i] = m < M;
      ra[i] = la.row(z, va[i] ? m : 0);
    }
  } else {
    const int m = m0 + (t & 63);
    va[0] = m < M;
    ra[0] = la.row(z, va[0] ? m : 0);
  }
  if (KFAST_B) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + (t >> 4) + 16 * i;
      vb[i] = n < N;
      rb[i] = lb.row(z, vb[i] ? n : 0);
    }
  } else {
    const int n = n0 + (t & 63);
    vb[0] = n < N;
    rb[0] = lb.row(z, vb[0] ? n : 0);
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kBegin; k0 < kEnd; k0 += BK) {
    float av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (KFAST_A) {
        const int64_t k = k0 + (t & 15);
        av[i] = (va[i] && k < kEnd) ? la.at(ra[i], k) : 0.f;
      } else {
        const int64_t k = k0 + (t >> 6) + 4 * i;
        av[i] = (va[0] && k < kEnd) ? la.at(ra[0], k) : 0.f;
      }
      if (KFAST_B) {
        const int64_t k = k0 + (t & 15);
        bv[i] = (vb[i] && k < kEnd) ? lb.at(rb[i], k) : 0.f;
      } else {
        const int64_t k = k0 + (t >> 6) + 4 * i;
        bv[i] = (vb[0] && k < kEnd) ? lb.at(rb[0], k) : 0.f;
      }
    }
    __syncthreads();  // previous slab fully consumed
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (KFAST_A) As[t & 15][(t >> 4) + 16 * i] = av[i]; else As[(t >> 6) + 4 * i][t & 63] = av[i];
      if (KFAST_B) Bs[t & 15][(t >> 4) + 16 * i] = bv[i]; else Bs[(t >> 6) + 4 * i][t & 63] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) epi(z, m, n, acc[i][j]);
    }
  }
}

template <bool KFAST_A, bool KFAST_B, typename LA, typename LB, typename EPI>
int launch(tnn_context *ctx, int64_t M, int64_t N, int64_t K, int batch, int kSplits, LA la, LB lb, EPI epi) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  TNN_REQUIRE(M < (1ll << 31) && N < (1ll << 31), "matrix dimension too large");
  if (kSplits < 1) kSplits = 1;
  int64_t kChunk = cdiv64(cdiv64(K, kSplits), BK) * BK;
  if (kChunk < BK) kChunk = BK;
  kSplits = (int)cdiv64(K > 0 ? K : 1, kChunk);
  dim3 grid((unsigned)cdiv64(N, BN), (unsigned)cdiv64(M, BM), (unsigned)(batch * kSplits));
  TNN_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "grid too large");
  k_gemm<KFAST_A, KFAST_B, LA, LB, EPI><<<grid, THREADS, 0, ctx->stream>>>((int)M, (int)N, K, kChunk, kSplits, la, lb,
                                                                          epi);
  TNN_LAUNCHED(ctx);
  return 0;
}

// ---------------------------------------------------------------- loaders

// plain strided-batched matrix, element (mn, k); `kfast` says whether k is the contiguous index
struct MatLoader {
  const float *p;
  int64_t ld, zstride;
  bool k_contig;  // true: elem = p[mn*ld + k]; false: elem = p[k*ld + mn]
  struct Row { const float *q; };
  __device__ Row row(int z, int mn) const {
    return Row{p + (int64_t)z * zstride + (k_contig ? (int64_t)mn * ld : (int64_t)mn)};
  }
  __device__ float at(const Row &r, int64_t k) const { return k_contig ? r.q[k] : r.q[k * ld]; }
};

// strideC == 0 batched accumulate (wgrad-style gemmStridedBatched): the batch is folded into k
struct MatLoaderFoldK {
  const float *p;
  int64_t ld, zstride, K1;  // K1 = per-batch K
  bool k_contig;
  struct Row { const float *q; };
  __device__ Row row(int, int mn) const { return Row{p + (k_contig ? (int64_t)mn * ld : (int64_t)mn)}; }
  __device__ float at(const Row &r, int64_t k) const {
    const int64_t b = k / K1, kk = k - b * K1;
    return k_contig ? r.q[b * zstride + kk] : r.q[b * zstride + kk * ld];
  }
};

struct ConvGeom {
  int C, H, W, F, kh, kw, ph, pw, sy, sx, dy, dx, OH, OW;
  int64_t inImg, outImg;  // C*H*W, F*OH*OW
};

// im2col gather, column n = (b, oh, ow) folded over the batch, k = (c, kr, kc)  [conv forward B operand]
struct Im2colCols {
  const float *in;
  ConvGeom g;
  struct Row { const float *base; int ih0, iw0; };
  __device__ Row row(int, int n) const {
    const int P = g.OH * g.OW;
    const int b = n / P, p = n - b * P;
    const int oh = p / g.OW, ow = p - oh * g.OW;
    return Row{in + (int64_t)b * g.inImg, oh * g.sy - g.ph, ow * g.sx - g.pw};
  }
  __device__ float at(const Row &r, int64_t k) const {
    const int kk = g.kh * g.kw;
    const int c = (int)(k / kk), rem = (int)(k - (int64_t)c * kk);
    const int kr = rem / g.kw, kc = rem - kr * g.kw;
    const int ih = r.ih0 + kr * g.dy, iw = r.iw0 + kc * g.dx;
    if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W) return __ldg(r.base + ((int64_t)c * g.H + ih) * g.W + iw);
    return 0.f;
  }
};

// wgrad A operand: delta[b][f][p], row = f, k = (b, p)
struct DeltaRowsFoldK {
  const float *delta;
  ConvGeom g;
  struct Row { const float *q; };
  __device__ Row row(int, int f) const { return Row{delta + (int64_t)f * g.OH * g.OW}; }
  __device__ float at(const Row &r, int64_t k) const {
    const int P = g.OH * g.OW;
    const int64_t b = k / P, p = k - b * P;
    return __ldg(r.q + b * g.outImg + p);
  }
};

// wgrad B operand: im2col rows, column n = (c, kr, kc), k = (b, oh, ow)
struct Im2colRowsFoldK {
  const float *in;
  ConvGeom g;
  struct Row { const float *base; int dh, dw; };
  __device__ Row row(int, int n) const {
    const int kk = g.kh * g.kw;
    const int c = n / kk, rem = n - c * kk;
    const int kr = rem / g.kw, kc = rem - kr * g.kw;
    return Row{in + (int64_t)c * g.H * g.W, kr * g.dy - g.ph, kc * g.dx - g.pw};
  }
  __device__ float at(const Row &r, int64_t k) const {
    const int P = g.OH * g.OW;
    const int64_t b = k / P;
    const int p = (int)(k - b * P);
    const int oh = p / g.OW, ow = p - oh * g.OW;
    const int ih = oh * g.sy + r.dh, iw = ow * g.sx + r.dw;
    if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W) return __ldg(r.base + b * g.inImg + (int64_t)ih * g.W + iw);
    return 0.f;
  }
};

// dgrad A operand: weights[f][c][kr][kc] read as row = c, k = (f, kr, kc)
struct WeightsByChannel {
  const float *w;
  ConvGeom g;
  struct Row { const float *q; };
  __device__ Row row(int, int c) const { return Row{w + (int64_t)c * g.kh * g.kw}; }
  __device__ float at(const Row &r, int64_t k) const {
    const int kk = g.kh * g.kw;
    const int64_t f = k / kk, rem = k - f * kk;
    return __ldg(r.q + f * (int64_t)g.C * kk + rem);
  }
};

// dgrad B operand: delta gathered per input pixel, column n = (b, ih, iw), k = (f, kr, kc).
// Uses the reference col2im origin (kernel_row - pad)*dilation (ntensors.pas:11693).
struct DeltaGatherCols {
  const float *delta;
  ConvGeom g;
  struct Row { const float *base; int ih, iw; };
  __device__ Row row(int, int n) const {
    const int HW = g.H * g.W;
    const int b = n / HW, p = n - b * HW;
    const int ih = p / g.W, iw = p - ih * g.W;
    return Row{delta + (int64_t)b * g.outImg, ih, iw};
  }
  __device__ float at(const Row &r, int64_t k) const {
    const int kk = g.kh * g.kw;
    const int f = (int)(k / kk), rem = (int)(k - (int64_t)f * kk);
    const int kr = rem / g.kw, kc = rem - kr * g.kw;
    const int hh = r.ih - (kr - g.ph) * g.dy, ww = r.iw - (kc - g.pw) * g.dx;
    if (hh < 0 || ww < 0 || hh % g.sy || ww % g.sx) return 0.f;
    const int oh = hh / g.sy, ow = ww / g.sx;
    if (oh >= g.OH || ow >= g.OW) return 0.f;
    return __ldg(r.base + ((int64_t)f * g.OH + oh) * g.OW + ow);
  }
};

}  // namespace simt
