/*
 * tensorium_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's (achinastone/Tensorium) live CPU
 * arithmetic for the convolution hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (libtensorium_b200.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 * for this path and cannot be compiled here (FreePascal is absent), so this
 * restatement is pinned only against hand-derived vectors (tests/golden/) and
 * an fp64 cross-check, not against reference outputs.
 *
 * Every function cites the reference file:line (relative to
 * /root/reference/source/) whose loop nest, evaluation order and rounding it
 * follows.  Two accumulation modes mirror the reference's two build flavours:
 *   ORC_MODE_PAS  (0)  default Makefiles (USE_AVX2=0): saxpy_pas / sdot_pas,
 *                      separate single-precision multiply and add
 *                      (ntensors.pas:1130-1145).
 *   ORC_MODE_FMA  (1)  -dUSE_AVX2 builds (MSCOCO_Yolo sample): saxpy_avx2 uses
 *                      vfmadd231ps (ntensors.pas:1308-1435), sdot_avx2 keeps 8
 *                      lane partial sums folded (l+l4)…, hadd, hadd
 *                      (ntensors.pas:1233-1306), vssum_avx2 likewise
 *                      (ntensors.pas:3590-3620).  Batch-norm keeps the
 *                      pure-Pascal branches in both modes (the AVX variants are
 *                      flagged unstable by the reference itself,
 *                      ntensors.pas:8853-8855,8934; srss drops lanes 4-7 when
 *                      N%8==0, ntensors.pas:1493-1527).
 *
 * Build: see oracle/Makefile (-O2 -mavx2 -mfma -ffp-contract=off -fopenmp).
 * -ffp-contract=off is REQUIRED so PAS mode keeps mul and add separately rounded.
 *
 * Threading mirrors steroids' static range split (steroids.pas:210-220):
 * rows of C for GEMM (ntensors.pas:2153), images for im2col/col2im
 * (ntensors.pas:11521,11867), channels for BN statistics (ntensors.pas:9169).
 * Results do not depend on the thread count (each unit is computed by one thread
 * in the reference order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t sz; /* Pascal SizeInt on x86-64 */

#define ORC_MODE_PAS 0
#define ORC_MODE_FMA 1

static int g_mode = ORC_MODE_PAS;
static const float sEPSILON = 0.000001f; /* ntensors.pas:95 */

void orc_set_mode(int mode) { g_mode = mode ? ORC_MODE_FMA : ORC_MODE_PAS; }
int orc_get_mode(void) { return g_mode; }
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- BLAS-1 */

/* saxpy_pas ntensors.pas:1130-1136 / saxpy_avx2 ntensors.pas:1308-1435 */
static inline void saxpy(sz N, float a, const float *x, float *y) {
  if (g_mode == ORC_MODE_FMA) {
    for (sz i = 0; i < N; ++i) y[i] = fmaf(a, x[i], y[i]);
  } else {
    for (sz i = 0; i < N; ++i) y[i] = a * x[i] + y[i];
  }
}

/* sdot_pas ntensors.pas:1138-1145 / sdot_avx2 ntensors.pas:1233-1306 (SIMD_REGS=8) */
static inline float sdot(sz N, const float *A, const float *B) {
  if (g_mode == ORC_MODE_FMA) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sz nb = N >> 3, i = 0;
    for (sz b = 0; b < nb; ++b, i += 8)
      for (int l = 0; l < 8; ++l) acc[l] = fmaf(A[i + l], B[i + l], acc[l]);
    int rem = (int)(N & 7);
    for (int l = 0; l < rem; ++l) acc[l] = fmaf(A[i + l], B[i + l], acc[l]);
    float s0 = acc[0] + acc[4], s1 = acc[1] + acc[5];
    float s2 = acc[2] + acc[6], s3 = acc[3] + acc[7];
    float h0 = s0 + s1, h1 = s2 + s3;
    return h0 + h1;
  }
  float r = 0;
  for (sz i = 0; i < N; ++i) r = r + A[i] * B[i];
  return r;
}

/* cblas_sdot ntensors.pas:2208-2220 */
static inline float cblas_sdot(sz N, const float *A, sz inca, const float *B, sz incb) {
  if (inca == 1 && incb == 1) return sdot(N, A, B);
  float r = 0;
  for (sz i = 0; i < N; ++i) r = r + A[i * inca] * B[i * incb];
  return r;
}

/* vsSumI ntensors.pas:3624-3635 (+ vssum_avx2 :3590-3620 when stride=1 in AVX2 builds) */
static inline float vsSumI(sz N, const float *src, sz stride) {
  if (g_mode == ORC_MODE_FMA && stride == 1) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sz nb = N >> 3, i = 0;
    for (sz b = 0; b < nb; ++b, i += 8)
      for (int l = 0; l < 8; ++l) acc[l] = acc[l] + src[i + l];
    float s0 = acc[0] + acc[4], s1 = acc[1] + acc[5];
    float s2 = acc[2] + acc[6], s3 = acc[3] + acc[7];
    float r = (s0 + s1) + (s2 + s3);
    for (sz k = 0; k < (N & 7); ++k) r = r + src[i + k];
    return r;
  }
  float r = 0;
  for (sz i = 0; i < N; ++i) r = r + src[i * stride];
  return r;
}

/* vsRSS ntensors.pas:3646-3658 (pure-Pascal branch) */
static inline float vsRSS(sz N, float mean, const float *src, sz stride) {
  float r = 0;
  for (sz i = 0; i < N; ++i) {
    float d = src[i * stride] - mean;
    r = r + d * d;
  }
  return r;
}

/* cblas_saxpy ntensors.pas:2381-2393 == TTensor.axpy for equal sizes (ntensors.pas:7603-7627) */
void orc_axpy(sz N, float a, const float *x, sz incx, float *y, sz incy) {
  if (incx == 1 && incy == 1) {
    saxpy(N, a, x, y);
    return;
  }
  for (sz i = 0; i < N; ++i) y[i * incy] = a * x[i * incx] + y[i * incy];
}

/* cblas_sscal ntensors.pas:1566-1578 == TTensor.Multiply(scalar) */
void orc_scal(sz N, float alpha, float *a, sz inca) {
  for (sz i = 0; i < N; ++i) a[i * inca] = alpha * a[i * inca];
}

float orc_dot(sz N, const float *A, sz inca, const float *B, sz incb) { return cblas_sdot(N, A, inca, B, incb); }
float orc_sum(sz N, const float *A, sz inca) { return vsSumI(N, A, inca); }

/* ---------------------------------------------------------------- SGEMM */

/* cblas_sgemm ntensors.pas:2231-2286 with s_nn :2061-2133, s_nt :1957-1983,
 * s_tn :2008-2034, s_tt :2159-2184.  Row-major.  beta is applied first as a
 * separate pass (C := beta*C when beta<>1), then alpha*op(A)*op(B) is ADDED.
 * The multithreaded build covers all M rows of the TT case (the single-thread
 * build's s_tt(0, M-2) skips the last row, ntensors.pas:2204 — not inherited). */
void orc_sgemm(int transA, int transB, sz M, sz N, sz K, float alpha, const float *A, sz lda,
               const float *B, sz ldb, float beta, float *C, sz ldc) {
  if (beta != 1.0f)
    for (sz i = 0; i < M; ++i) orc_scal(N, beta, C + i * ldc, 1);

  if (!transA && !transB) {
#pragma omp parallel for schedule(static)
    for (sz i = 0; i < M; ++i) {
      const float *Ai = A + i * lda;
      float *Ci = C + i * ldc;
      for (sz kk = 0; kk < K; ++kk) {
        float a_part = alpha * Ai[kk];
        saxpy(N, a_part, B + kk * ldb, Ci);
      }
    }
  } else if (!transA && transB) {
#pragma omp parallel for schedule(static)
    for (sz i = 0; i < M; ++i)
      for (sz j = 0; j < N; ++j) {
        float sum = alpha * sdot(K, A + i * lda, B + j * ldb);
        C[i * ldc + j] = C[i * ldc + j] + sum;
      }
  } else if (transA && !transB) {
#pragma omp parallel for schedule(static)
    for (sz i = 0; i < M; ++i)
      for (sz kk = 0; kk < K; ++kk) {
        float a_part = alpha * A[kk * lda + i];
        saxpy(N, a_part, B + kk * ldb, C + i * ldc);
      }
  } else {
#pragma omp parallel for schedule(static)
    for (sz i = 0; i < M; ++i)
      for (sz j = 0; j < N; ++j) {
        float sum = 0;
        for (sz kk = 0; kk < K; ++kk) sum = sum + alpha * A[i + kk * lda] * B[kk + j * ldb];
        C[i * ldc + j] = C[i * ldc + j] + sum;
      }
  }
}

/* cblas_sgemm_batch_strided ntensors.pas:2288-2304 (sequential over the batch,
 * so strideC=0 accumulates — the wgrad case, nConvolutionLayer.pas:636-640) */
void orc_sgemm_strided_batched(int transA, int transB, sz M, sz N, sz K, float alpha, const float *A,
                               sz lda, sz strideA, const float *B, sz ldb, sz strideB, float beta,
                               float *C, sz ldc, sz strideC, sz batchCount) {
  for (sz i = 0; i < batchCount; ++i)
    orc_sgemm(transA, transB, M, N, K, alpha, A + i * strideA, lda, B + i * strideB, ldb, beta,
              C + i * strideC, ldc);
}

/* fp64 "truth" GEMM for error attribution (not a reference routine) */
void orc_dgemm_truth(int transA, int transB, sz M, sz N, sz K, float alpha, const float *A, sz lda,
                     const float *B, sz ldb, float beta, const float *Cin, double *Cout, sz ldc) {
#pragma omp parallel for schedule(static)
  for (sz i = 0; i < M; ++i)
    for (sz j = 0; j < N; ++j) {
      double s = 0;
      for (sz kk = 0; kk < K; ++kk) {
        double a = transA ? A[kk * lda + i] : A[i * lda + kk];
        double b = transB ? B[j * ldb + kk] : B[kk * ldb + j];
        s += a * b;
      }
      Cout[i * ldc + j] = (double)beta * (Cin ? Cin[i * ldc + j] : 0.0) + (double)alpha * s;
    }
}

/* ---------------------------------------------------------------- im2col / col2im */

static inline sz out_dim(sz in, sz pad, sz k, sz stride, sz dil) {
  return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}
sz orc_out_dim(sz in, sz pad, sz k, sz stride, sz dil) { return out_dim(in, pad, k, stride, dil); }

/* sim2Col ntensors.pas:11415-11491: col[c][kh][kw][oh][ow] = im[c][oh*sy-ph+kh*dy][ow*sx-pw+kw*dx]
 * or 0; bounds via unsigned compare (:11441,:11447).  Witness: im2col_cpu_ext nCol2Im.pas:321-422. */
void orc_im2col(sz channels, sz height, sz width, sz kh, sz kw, sz padH, sz padW, sz strideY,
                sz strideX, sz dilY, sz dilX, const float *im, sz imOffset, float *col, sz colOffset) {
  sz output_w = out_dim(width, padW, kw, strideX, dilX);
  sz output_h = out_dim(height, padH, kh, strideY, dilY);
  sz channel_size = height * width;
  sz out_channel_size = output_w * output_h;
  sz kernel_size = kw * kh;
  for (sz idx = 0; idx < channels; ++idx) {
    const float *d_im = im + imOffset + channel_size * idx;
    float *d_col = col + colOffset + kernel_size * out_channel_size * idx;
    for (sz kernel_row = 0; kernel_row < kh; ++kernel_row)
      for (sz kernel_col = 0; kernel_col < kw; ++kernel_col) {
        sz input_row = -padH + kernel_row * dilY;
        for (sz output_rows = 0; output_rows < output_h; ++output_rows) {
          if ((uint64_t)input_row < (uint64_t)height) {
            sz input_col = -padW + kernel_col * dilX;
            for (sz output_col = 0; output_col < output_w; ++output_col) {
              if ((uint64_t)input_col < (uint64_t)width)
                d_col[output_col] = d_im[input_row * width + input_col];
              else
                d_col[output_col] = 0;
              input_col += strideX;
            }
          } else {
            for (sz output_col = 0; output_col < output_w; ++output_col) d_col[output_col] = 0;
          }
          d_col += output_w;
          input_row += strideY;
        }
      }
  }
}

/* sim2colStridedBatched ntensors.pas:11493-11532 */
void orc_im2col_strided_batched(sz channels, sz height, sz width, sz kh, sz kw, sz padH, sz padW,
                                sz strideY, sz strideX, sz dilY, sz dilX, const float *im, sz imStride,
                                sz imOffset, float *col, sz colStride, sz colOffset, sz batchCount) {
#pragma omp parallel for schedule(static)
  for (sz b = 0; b < batchCount; ++b)
    orc_im2col(channels, height, width, kh, kw, padH, padW, strideY, strideX, dilY, dilX,
               im + b * imStride, imOffset, col + b * colStride, colOffset);
}

/* scol2im + c2i ntensors.pas:11650-11763: scatter-ADD (caller zeroes im).  Note the reference
 * computes input_row := (kernel_row - pad_h) * dilation_h (:11693), which equals im2col's
 * -pad + kernel_row*dil only for dilation=1; restated as written.
 * Witness: col2im_cpu_ext nCol2Im.pas:253-297. */
void orc_col2im(sz channels, sz height, sz width, sz kh, sz kw, sz padH, sz padW, sz strideH,
                sz strideW, sz dilH, sz dilW, const float *col, sz colOffset, float *im, sz imOffset) {
  sz output_h = out_dim(height, padH, kh, strideH, dilH);
  sz output_w = out_dim(width, padW, kw, strideW, dilW);
  sz out_channel_size = output_h * output_w;
  sz channel_size = height * width;
  sz kernel_size = kh * kw;
  for (sz i = 0; i < channels * kernel_size; ++i) {
    sz chan = i / kernel_size;
    const float *data_col = col + colOffset + i * out_channel_size;
    float *data_im = im + imOffset + channel_size * chan;
    sz index = i % kernel_size;
    sz kernel_row = index / kw;
    sz kernel_col = index % kw;
    sz input_row = (kernel_row - padH) * dilH;
    for (sz output_rows = 0; output_rows < output_h; ++output_rows) {
      if ((uint64_t)input_row >= (uint64_t)height) {
        data_col += output_w;
      } else {
        sz input_col = (kernel_col - padW) * dilW;
        for (sz output_col = 0; output_col < output_w; ++output_col) {
          if ((uint64_t)input_col < (uint64_t)width) {
            sz ix = input_row * width + input_col;
            data_im[ix] = data_im[ix] + data_col[0];
          }
          ++data_col;
          input_col += strideW;
        }
      }
      input_row += strideH;
    }
  }
}

/* scol2imStridedBatched ntensors.pas:11833-11879 */
void orc_col2im_strided_batched(sz channels, sz height, sz width, sz kh, sz kw, sz padH, sz padW,
                                sz strideY, sz strideX, sz dilY, sz dilX, const float *col,
                                sz colStride, sz colOffset, float *im, sz imStride, sz imOffset,
                                sz batchCount) {
#pragma omp parallel for schedule(static)
  for (sz b = 0; b < batchCount; ++b)
    orc_col2im(channels, height, width, kh, kw, padH, padW, strideY, strideX, dilY, dilX,
               col + b * colStride, colOffset, im + b * imStride, imOffset);
}

/* TTensor.Conv2D ntensors.pas:8252-8349: 1x1/stride1/dil1 skips im2col (:8286-8312), then one
 * NN GEMM per image with beta=0 (:8326-8333).  workspace must hold batch*C*kh*kw*OH*OW floats
 * unless the 1x1 shortcut applies (may be NULL then). */
void orc_conv2d(sz batch, sz c, sz h, sz w, sz filters, sz kh, sz kw, sz padH, sz padW, sz strideY,
                sz strideX, sz dilY, sz dilX, const float *in, const float *weights, float *out,
                float *workspace) {
  sz oh = out_dim(h, padH, kh, strideY, dilY);
  sz ow = out_dim(w, padW, kw, strideX, dilX);
  sz kSize = kh * kw;
  sz outImgSize = oh * ow;
  sz k = c * kSize;
  sz imColSize = c * kSize * outImgSize;
  sz strideC = outImgSize * filters;
  const float *colPtr;
  sz strideB;
  if (kSize != 1 || dilX * dilY != 1 || strideX * strideY != 1) {
    strideB = imColSize;
    orc_im2col_strided_batched(c, h, w, kh, kw, padH, padW, strideY, strideX, dilY, dilX, in,
                               c * h * w, 0, workspace, strideB, 0, batch);
    colPtr = workspace;
  } else {
    strideB = c * h * w;
    colPtr = in;
  }
  for (sz b = 0; b < batch; ++b)
    orc_sgemm(0, 0, filters, outImgSize, k, 1.0f, weights, k, colPtr + b * strideB, outImgSize, 0.0f,
              out + b * strideC, outImgSize);
}

/* ---------------------------------------------------------------- bias / scale / reductions */

/* TTensor.forwardBias ntensors.pas:7709-7727 -> vsAddB ntensors.pas:4066-4091 */
void orc_forward_bias(float *dst, sz dstSize, sz groups, const float *bias, sz N) {
  sz blockSize = dstSize / (groups * N);
  for (sz k = 0; k < groups; ++k)
    for (sz i = 0; i < N; ++i) {
      float *c = dst + (k * N + i) * blockSize;
      float bb = bias[i];
      for (sz j = 0; j < blockSize; ++j) c[j] = c[j] + bb;
    }
}

/* TTensor.forwardScale ntensors.pas:7687-7707 -> vsMulB ntensors.pas:4112-4139 */
void orc_forward_scale(float *dst, sz dstSize, sz groups, const float *scale, sz N) {
  sz blockSize = dstSize / (groups * N);
  for (sz k = 0; k < groups; ++k)
    for (sz i = 0; i < N; ++i) {
      float *c = dst + (k * N + i) * blockSize;
      float bb = scale[i];
      for (sz j = 0; j < blockSize; ++j) c[j] = c[j] * bb;
    }
}

/* TTensor.addSums ntensors.pas:7729-7781 (bias gradient): dst[i] += sum_b sum(block) */
void orc_add_sums(float *dst, sz nDst, const float *src, sz N, sz srcGroups) {
  sz blockSize = N / (srcGroups * nDst);
  if (blockSize == 1) {
    for (sz i = 0; i < nDst; ++i) dst[i] = dst[i] + vsSumI(srcGroups, src + i, nDst);
    return;
  }
  for (sz i = 0; i < nDst; ++i) {
    float s = 0;
    for (sz j = 0; j < srcGroups; ++j) s = s + vsSumI(blockSize, src + (j * nDst + i) * blockSize, 1);
    dst[i] = dst[i] + s;
  }
}

/* TTensor.addDots ntensors.pas:7783-7837: dst[i] += sum_b dot(src1 block, src2 block) */
void orc_add_dots(float *dst, sz nDst, const float *src1, const float *src2, sz N, sz groups) {
  sz blockSize = N / (nDst * groups);
  if (blockSize == 1) {
    for (sz i = 0; i < nDst; ++i) dst[i] = dst[i] + cblas_sdot(groups, src1 + i, nDst, src2 + i, nDst);
    return;
  }
  for (sz i = 0; i < nDst; ++i) {
    float s = 0;
    for (sz b = 0; b < groups; ++b) {
      sz idx = (i + b * nDst) * blockSize;
      s = s + cblas_sdot(blockSize, src1 + idx, 1, src2 + idx, 1);
    }
    dst[i] = dst[i] + s;
  }
}

/* TTensor.MeansAndVars ntensors.pas:9102-9177: two-pass, mean = sum/S, var = rss/(S-1) */
void orc_means_and_vars(const float *x, sz size, sz groups, float *means, float *vars, sz N) {
  sz blockSize = size / (N * groups);
  float S = (float)(groups * blockSize);
  float S2 = (float)(groups * blockSize - 1);
#pragma omp parallel for schedule(static)
  for (sz idx = 0; idx < N; ++idx) {
    float m = 0, v = 0;
    for (sz b = 0; b < groups; ++b) m = m + vsSumI(blockSize, x + (idx + b * N) * blockSize, 1);
    m = m / S;
    means[idx] = m;
    for (sz b = 0; b < groups; ++b) v = v + vsRSS(blockSize, m, x + (idx + b * N) * blockSize, 1);
    vars[idx] = v / S2;
  }
}

/* TTensor.blockNormalize ntensors.pas:8693-8718 -> _snormblkvv :4357-4382 (blockSize>1:
 * d = max(sqrt(var), eps)) / _snormvv :4330-4340 (blockSize=1: sqrt(max(var, eps))) */
void orc_block_normalize(float *x, sz size, sz groups, const float *mean, const float *variance, sz N) {
  sz blockSize = size / (groups * N);
  for (sz g = 0; g < groups; ++g) {
    float *D = x + g * blockSize * N;
    if (blockSize == 1) {
      for (sz i = 0; i < N; ++i) D[i] = (D[i] - mean[i]) / sqrtf(fmaxf(variance[i], sEPSILON));
    } else {
      for (sz i = 0; i < N; ++i) {
        float *o = D + i * blockSize;
        float d = fmaxf(sqrtf(variance[i]), sEPSILON);
        float m = mean[i];
        for (sz j = 0; j < blockSize; ++j) o[j] = (o[j] - m) / d; /* snormvss :4315-4321 */
      }
    }
  }
}

/* sMeanAndVarianceDelta ntensors.pas:8831-8899 (pure-Pascal branch).  Pascal evaluates
 * -1.0/sqrt(..) and Power(.., -3/2) in extended precision and rounds the product to single. */
void orc_mean_and_variance_delta(const float *delta, const float *x, sz N, sz groups, const float *mean,
                                 const float *variance, float *mean_delta, float *variance_delta,
                                 sz nDst) {
  sz blockSize = N / (nDst * groups);
#pragma omp parallel for schedule(static)
  for (sz i = 0; i < nDst; ++i) {
    float m = 0, v = 0;
    for (sz j = 0; j < groups; ++j) {
      m = m + vsSumI(blockSize, delta + (i + j * nDst) * blockSize, 1);
      for (sz k = 0; k < blockSize; ++k) {
        sz index = (i + j * nDst) * blockSize + k;
        v = v + delta[index] * (x[index] - mean[i]);
      }
    }
    long double vv = fmaxf(variance[i], sEPSILON);
    mean_delta[i] = (float)((long double)m * (-1.0L / sqrtl(vv)));
    variance_delta[i] = (float)((long double)v * -0.5L * powl(vv, -3.0L / 2.0L));
  }
}

/* sNormalizeDelta ntensors.pas:8902-8951 -> sNormalizeDelta_pas :8818-8828 */
void orc_normalize_delta(float *delta, const float *x, sz size, sz groups, const float *mean,
                         const float *variance, const float *mean_delta, const float *variance_delta,
                         sz N) {
  sz blockSize = size / (N * groups);
  sz batchSize = blockSize * groups;
  for (sz j = 0; j < groups; ++j)
    for (sz i = 0; i < N; ++i) {
      sz index = (i + j * N) * blockSize;
      float *dd = delta + index;
      const float *xx = x + index;
      float meanDelta = mean_delta[i] / (float)batchSize;
      float varDelta = 2 * variance_delta[i] / (float)batchSize;
      float aMean = mean[i];
      float aStdDev = sqrtf(fmaxf(variance[i], sEPSILON));
      for (sz k = 0; k < blockSize; ++k) dd[k] = dd[k] / aStdDev + (varDelta * (xx[k] - aMean) + meanDelta);
    }
}

/* ---------------------------------------------------------------- activations */

/* TActivationType ordinals ntypes.pas:66-71 */
enum {
  acLOGISTIC = 0, acRELU, acRELU6, acRELIE, acLINEAR, acRAMP, acTANH, acPLSE, acREVLEAKY, acLEAKY,
  acELU, acLOGGY, acSTAIR, acHARDTAN, acLHTAN, acSELU, acSOFTMAX, acGELU, acSWISH, acMISH,
  acHARD_MISH, acNORM_CHAN, acNORM_CHAN_SOFTMAX, acNORM_CHAN_SOFTMAX_MAXVAL
};

/* Scalar activation functions nactivation.pas:272-401.  FPC evaluates mixed single/real
 * expressions (exp(), real-typed literals such as 0.1) in extended precision on x86-64 Linux and
 * rounds once on assignment to single; long double mirrors that. */
typedef long double ext;
static inline float act_one(float xf, int a) {
  ext x = xf;
  switch (a) {
    case acLOGISTIC: return (float)(1.0L / (1.0L + expl(-x)));                 /* :294-298 */
    case acRELU: return xf * (float)(xf > 0);                                  /* :306-311 */
    case acRELU6: return xf * (float)(xf > 0) * (float)(xf <= 6);              /* :313-318 (0 above 6) */
    case acRELIE: return xf > 0 ? xf : (float)(0.01L * x);                     /* :335-339 */
    case acLINEAR: return xf;
    case acRAMP: return (float)(x * (ext)(xf > 0) + 0.1L * x);                 /* :341-344 */
    case acTANH: { ext px = expl(x), nx = expl(-x); return (float)((px - nx) / (px + nx)); } /* :352-361 */
    case acPLSE:                                                               /* :373-378 */
      if (xf < -4) return (float)(0.01L * (x + 4));
      if (xf > 4) return (float)(0.01L * (x - 4) + 1);
      return (float)(0.125L * x + 0.5L);
    case acREVLEAKY:
    case acLEAKY: return xf < 0 ? (float)(0.1L * x) : xf;                      /* activate_array :555-566 */
    case acELU: return (float)((ext)(xf >= 0) * x + (ext)(xf < 0) * (expl(x) - 1)); /* :320-323 */
    case acLOGGY: return (float)(2.0L / (1.0L + expl(-x)) - 1);                /* :300-304 */
    case acSTAIR: {                                                            /* :272-279 */
      sz n = (sz)floorl(x);
      if (n % 2 == 0) return (float)floorl(x / 2);
      return (float)((x - n) + floorl(x / 2));
    }
    case acHARDTAN: return xf < -1 ? -1.0f : (xf > 1 ? 1.0f : xf);             /* :281-286 */
    case acLHTAN:                                                              /* :380-385 */
      if (xf < 0) return (float)(0.001L * x);
      if (xf > 1) return (float)(0.001L * (x - 1) + 1);
      return xf;
    case acSELU: return (float)((ext)(xf >= 0) * 1.0507L * x + (ext)(xf < 0) * 1.0507L * 1.6732L * (expl(x) - 1)); /* :325-328 */
    case acGELU: return (float)(0.5L * x * (1 + tanhl(0.797885L * x + 0.035677L * powl(x, 3)))); /* :330-333 */
    case acSWISH: return (float)(x * (1.0L / (1.0L + expl(-x))));              /* activate_array :596-598 */
    default: return xf;
  }
}

/* Gradients take the layer OUTPUT nactivation.pas:403-501 */
static inline float grad_one(float xf, int a) {
  ext x = xf;
  switch (a) {
    case acLOGISTIC: return (float)((1 - x) * x);                              /* :421-424 */
    case acRELU: return (float)(xf > 0);
    case acRELU6: return (float)((xf > 0) && (xf < 6));                        /* :445-448 */
    case acRELIE: return xf > 0 ? 1.0f : 0.01f;
    case acLINEAR: return 1.0f;
    case acRAMP: return (float)((ext)(xf > 0) + 0.1L);
    case acTANH: return (float)(1 - x * x);
    case acPLSE: return (xf < 0 || xf > 1) ? 0.01f : 0.125f;
    case acREVLEAKY:
    case acLEAKY: return xf > 0 ? 1.0f : 0.1f;
    case acELU: return (float)((ext)(xf >= 0) + (ext)(xf < 0) * (x + 1));
    case acLOGGY: { ext y = (x + 1.0L) / 2.0L; return (float)(2 * (1 - y) * y); }
    case acSTAIR: return floorl(x) == x ? 0.0f : 1.0f;
    case acHARDTAN: return (xf > -1 && xf < 1) ? 1.0f : 0.0f;
    case acLHTAN: return (xf > 0 && xf < 1) ? 1.0f : 0.001f;
    case acSELU: return (float)((ext)(xf >= 0) * 1.0507L + (ext)(xf < 0) * (x + 1.0507L * 1.6732L));
    case acGELU: {                                                             /* :483-488 */
      ext x3 = powl(x, 3);
      ext t = 0.0356774L * x3 + 0.797885L * x;
      ext sech = 2.0L / (expl(t) + expl(-t));
      return (float)(0.5L * tanhl(t) + (0.0535161L * x3 + 0.398942L * x) * sech * sech + 0.5L);
    }
    default: return 1.0f;
  }
}

/* activate_array nactivation.pas:508-621 (scalar branches; in place) */
void orc_activate_array(float *x, sz N, int a) {
  if (a == acLINEAR) return;
  if (a == acSOFTMAX) { /* softmax_activate nactivation.pas:387-401 */
    float mx = x[0];
    for (sz i = 1; i < N; ++i) if (x[i] > mx) mx = x[i];
    for (sz i = 0; i < N; ++i) x[i] = (float)expl((ext)x[i] - (ext)mx);
    float s = vsSumI(N, x, 1);
    for (sz i = 0; i < N; ++i) x[i] = x[i] / s;
    return;
  }
  for (sz i = 0; i < N; ++i) x[i] = act_one(x[i], a);
}

/* gradient_array nactivation.pas:624-717: delta[i] *= f'(output[i]) */
void orc_gradient_array(const float *x, sz N, int a, float *delta) {
  if (a == acLINEAR) return;
  for (sz i = 0; i < N; ++i) delta[i] = delta[i] * grad_one(x[i], a);
}

/* ---------------------------------------------------------------- conv layer helpers */

/* TBaseConvolutionalLayer.fuseBatchNorm nConvolutionLayer.pas:102-126 */
void orc_fuse_batchnorm(float *weights, float *biases, const float *scales, const float *rolling_mean,
                        const float *rolling_variance, sz filters, sz filter_size) {
  for (sz f = 0; f < filters; ++f) {
    float precomputed = scales[f] / sqrtf(fmaxf(rolling_variance[f], sEPSILON));
    biases[f] = biases[f] - rolling_mean[f] * precomputed;
    for (sz i = 0; i < filter_size; ++i) {
      sz w_index = f * filter_size + i;
      weights[w_index] = weights[w_index] * precomputed;
    }
  }
}

/* misc elementwise used by the callers either side of the path */
void orc_addvv(sz N, const float *a, const float *b, float *c) { for (sz i = 0; i < N; ++i) c[i] = a[i] + b[i]; }
void orc_fill(sz N, float *x, float v) { for (sz i = 0; i < N; ++i) x[i] = v; }

/* ---------------------------------------------------------------------------------------------
 * Callers either side of the convolution on the CIFAR-10 training path.
 * ------------------------------------------------------------------------------------------- */

/* TTensor.Clamp ntensors.pas:10077-10100 (used by TConnectedLayer.backward nconnectedlayer.pas:261) */
void orc_clamp(sz N, float lo, float hi, float *x) {
  for (sz i = 0; i < N; ++i) {
    if (x[i] < lo) x[i] = lo;
    else if (x[i] > hi) x[i] = hi;
  }
}

/* TMaxPoolLayer.forwardMaxPool nMaxPoolLayer.pas:404-451 (training branch; the inference helper
 * maxpoolMP :265-320 walks the same window in the same order).  indexes may be NULL. */
void orc_maxpool_forward(sz batch, sz c, sz h, sz w, sz outH, sz outW, sz ksize, sz stride_x, sz stride_y,
                         sz padding, const float *in, float *out, sz *indexes) {
  const sz w_offset = -(padding / 2), h_offset = -(padding / 2);
  for (sz b = 0; b < batch; ++b)
    for (sz k = 0; k < c; ++k)
      for (sz i = 0; i < outH; ++i)
        for (sz j = 0; j < outW; ++j) {
          sz out_index = j + outW * (i + outH * (k + c * b));
          float max = -3.402823466e+38F;
          sz max_i = -1;
          for (sz n = 0; n < ksize; ++n)
            for (sz m = 0; m < ksize; ++m) {
              sz cur_h = h_offset + i * stride_y + n;
              sz cur_w = w_offset + j * stride_x + m;
              sz index = cur_w + w * (cur_h + h * (k + b * c));
              float val = (cur_h >= 0 && cur_h < h && cur_w >= 0 && cur_w < w) ? in[index] : -3.402823466e+38F;
              if (val > max) {
                max_i = index;
                max = val;
              }
            }
          out[out_index] = max;
          if (indexes) indexes[out_index] = max_i;
        }
}

/* TMaxPoolLayer.backwardMaxPool nMaxPoolLayer.pas:501-513 */
void orc_maxpool_backward(sz total, const sz *indexes, const float *delta, float *upstream) {
  for (sz i = 0; i < total; ++i) upstream[indexes[i]] = upstream[indexes[i]] + delta[i];
}

/* softmax nsoftmaxlayer.pas:83-105 and TSoftmaxLayer.softmaxBatch :108-121.  exp is the single-precision
 * libm expf (FPC's exp(single) evaluates in extended and rounds; the two agree to 1 ulp). */
static void softmax_one(sz n, const float *input, float temp, sz stride, float *output) {
  if (n == 0) return;
  float sum = 0, largest = input[0];
  for (sz i = 1; i < n; ++i)
    if (input[i * stride] > largest) largest = input[i * stride];
  for (sz i = 0; i < n; ++i) {
    float e = expf((input[i * stride] - largest) / temp);
    sum = sum + e;
    output[i * stride] = e;
  }
  for (sz i = 0; i < n; ++i) output[i * stride] = output[i * stride] / sum;
}
void orc_softmax_batch(const float *input, sz n, sz batch, sz batch_size, sz groups, sz group_size, sz stride,
                       float temp, float *output) {
  for (sz b = 0; b < batch; ++b)
    for (sz g = 0; g < groups; ++g)
      softmax_one(n, input + b * batch_size + g * group_size, temp, stride, output + b * batch_size + g * group_size);
}

/* TSoftmaxLayer.softmaxCrossEntropy nsoftmaxlayer.pas:123-137 */
void orc_softmax_cross_entropy(sz N, const float *pred, const float *truth, float *delta, float *error) {
  for (sz i = 0; i < N; ++i) {
    float t = truth[i], p = pred[i];
    error[i] = t != 0 ? -logf(p > sEPSILON ? p : sEPSILON) : 0.f;
    delta[i] = t - p;
  }
}

/* TDropoutLayer.forward / backward ndropoutlayer.pas:82-133.  The CPU path draws r from FPC's random();
 * the CUDA path (cuda_sgemm.cu:1620-1653 via nncuda.pas:1512-1538) hashes seed+i with xorshift64* and keeps
 * 28 bits.  The oracle takes the hashed form so both sides see the same mask for a given seed. */
void orc_dropout_rand(sz N, uint64_t seed, float *rnd) {
  for (sz i = 0; i < N; ++i) {
    uint64_t x = seed + (uint64_t)i;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    x *= 0x2545F4914F6CDD1DULL;
    float r = (float)(x % 0x10000000ull);
    rnd[i] = r / (float)0x10000000ull;
  }
}
void orc_dropout_apply(sz N, const float *rnd, float probability, float scale, float *x) {
  for (sz i = 0; i < N; ++i) x[i] = rnd[i] < probability ? 0.f : x[i] * scale;
}
