// misc.cu — im2col / col2im (explicit buffers, kept for drop-in callers; the conv entry points never
// materialise them), pooling, upsample, softmax and route concat.
// Replaces kernels 16-18, 20-25 of the reference (source/cuda_sgemm.cu; wrappers nncuda.pas:1056-1286).
#include "common.cuh"

namespace {
constexpr int kT = 256;
inline int grid1d(int64_t n, int t = kT) {
  int64_t b = cdiv64(n, t);
  if (b < 1) b = 1;
  if (b > 148 * 32) b = 148 * 32;
  return (int)b;
}
__host__ __device__ inline int64_t conv_out(int64_t in, int64_t pad, int64_t k, int64_t stride, int64_t dil) {
  return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}
}  // namespace

// col[(c*kh+kr)*kw+kc][oh][ow] = im[c][oh*sy-ph+kr*dy][ow*sx-pw+kc*dx] or 0   (sim2Col ntensors.pas:11415-11491)
__global__ void __launch_bounds__(kT) k_im2col(int64_t total, int C, int H, int W, int kh, int kw, int ph, int pw,
                                               int sy, int sx, int dy, int dx, int OH, int OW,
                                               const float *__restrict__ im, float *__restrict__ col) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int ow = (int)(i % OW);
    int64_t t = i / OW;
    const int oh = (int)(t % OH);
    t /= OH;
    const int kc = (int)(t % kw);
    t /= kw;
    const int kr = (int)(t % kh);
    const int c = (int)(t / kh);
    const int ih = oh * sy - ph + kr * dy;
    const int iw = ow * sx - pw + kc * dx;
    float v = 0.f;
    if ((unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W) v = __ldg(im + ((int64_t)c * H + ih) * W + iw);
    col[i] = v;
  }
}

// Gather form of scol2im (ntensors.pas:11650-11763): every im element sums the col entries that map to it,
// in ascending (kernel_row, kernel_col) order — the order the sequential CPU scatter adds them in —
// and the sum is ADDED to the existing im value (the caller zeroes, nnet.pas:287-295).
// Follows the reference's (kernel_row - pad)*dilation origin (ntensors.pas:11693).
__global__ void __launch_bounds__(kT) k_col2im(int64_t total, int C, int H, int W, int kh, int kw, int ph, int pw,
                                               int sy, int sx, int dy, int dx, int OH, int OW,
                                               const float *__restrict__ col, float *__restrict__ im) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int iw = (int)(i % W);
    const int64_t t = i / W;
    const int ih = (int)(t % H);
    const int c = (int)(t / H);
    float acc = im[i];
    for (int kr = 0; kr < kh; ++kr) {
      const int hh = ih - (kr - ph) * dy;
      if (hh < 0 || hh % sy) continue;
      const int oh = hh / sy;
      if (oh >= OH) continue;
      for (int kc = 0; kc < kw; ++kc) {
        const int ww = iw - (kc - pw) * dx;
        if (ww < 0 || ww % sx) continue;
        const int ow = ww / sx;
        if (ow >= OW) continue;
        acc += __ldg(col + ((((int64_t)c * kh + kr) * kw + kc) * OH + oh) * OW + ow);
      }
    }
    im[i] = acc;
  }
}

// forward_maxpool cuda_sgemm.cu:857-903 (offset = -padding/2, first maximum wins, 64-bit indexes)
__global__ void __launch_bounds__(kT) k_maxpool_fwd(int64_t total, int outH, int outW, const float *__restrict__ in,
                                                    int h, int w, int sx, int sy, int padding, int ksize,
                                                    unsigned long long *idx, float *__restrict__ out) {
  const int off = -padding / 2;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += nthreads) {
    const int x = (int)(o % outW);
    const int64_t t = o / outW;
    const int y = (int)(t % outH);
    const int64_t k = t / outH;
    float mx = -3.402823466e+38F;
    long long mi = -1;
    for (int n = 0; n < ksize; ++n)
      for (int m = 0; m < ksize; ++m) {
        const int ch = off + y * sy + n, cw = off + x * sx + m;
        const long long index = cw + (long long)w * (ch + (long long)h * k);
        const float val = (ch >= 0 && ch < h && cw >= 0 && cw < w) ? __ldg(in + index) : -3.402823466e+38F;
        if (val > mx) { mi = index; mx = val; }
      }
    out[o] = mx;
    if (idx) idx[o] = (unsigned long long)mi;
  }
}

// backward_maxpool cuda_sgemm.cu:905-916: output[indexes[id]] += delta[id]; windows can overlap -> atomic
__global__ void __launch_bounds__(kT) k_maxpool_bwd(int64_t total, float *out, const unsigned long long *idx,
                                                    const float *__restrict__ delta) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    // a window that held no value greater than -FLT_MAX (all NaN, all padding) has no arg-max: the forward pass stored
    // index -1; such an output routes no gradient instead of writing out of bounds
    const unsigned long long j = idx[i];
    if (j != ~0ull) atomicAdd(out + j, delta[i]);
  }
}

// upsample cuda_sgemm.cu:1274-1296: N = batch*channels planes, out is (h*stride) x (w*stride)
__global__ void __launch_bounds__(kT) k_upsample_fwd(int64_t total, int h, int w, int stride, float scale,
                                                     const float *__restrict__ in, float *__restrict__ out) {
  const int W2 = w * stride, H2 = h * stride;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += nthreads) {
    const int x = (int)(o % W2);
    const int64_t t = o / W2;
    const int y = (int)(t % H2);
    const int64_t c = t / H2;
    out[o] = scale * __ldg(in + (c * h + y / stride) * w + x / stride);
  }
}
// backward: in[c][y][x] (+)= scale * sum over the stride x stride block of out (gather form, no atomics)
__global__ void __launch_bounds__(kT) k_upsample_bwd(int64_t total, int h, int w, int stride, float scale, float *in,
                                                     const float *__restrict__ out, int zero) {
  const int W2 = w * stride;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int x = (int)(i % w);
    const int64_t t = i / w;
    const int y = (int)(t % h);
    const int64_t c = t / h;
    float acc = zero ? 0.f : in[i];
    for (int dy = 0; dy < stride; ++dy)
      for (int dx = 0; dx < stride; ++dx)
        acc += scale * __ldg(out + ((c * h + y) * stride + dy) * (int64_t)W2 + x * stride + dx);
    in[i] = acc;
  }
}

// softmaxBatch cuda_sgemm.cu:918-952: one thread per (batch, group) row, temperature scaled
__global__ void k_softmax(int64_t batch, int64_t groups, const float *in, int64_t n, int64_t batch_size,
                          int64_t group_size, int64_t stride, float temp, float *out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= batch * groups) return;
  const int64_t b = r / groups, g = r % groups;
  const float *p = in + b * batch_size + g * group_size;
  float *q = out + b * batch_size + g * group_size;
  float largest = p[0];
  for (int64_t i = 1; i < n; ++i) largest = fmaxf(largest, p[i * stride]);
  float sum = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float e = expf((p[i * stride] - largest) / temp);
    sum += e;
    q[i * stride] = e;
  }
  for (int64_t i = 0; i < n; ++i) q[i * stride] /= sum;
}

// out[b][cOff + c][:] = in[b][c][:]
__global__ void __launch_bounds__(kT) k_concat(int64_t total, int64_t cb, int64_t outCb, int64_t offCb,
                                               const float *__restrict__ in, float *__restrict__ out) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int64_t b = i / cb, r = i - b * cb;
    out[b * outCb + offCb + r] = in[i];
  }
}
__global__ void __launch_bounds__(kT) k_concat4(int64_t total4, int64_t cb4, int64_t outCb4, int64_t offCb4,
                                                const float4 *__restrict__ in, float4 *__restrict__ out) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += nthreads) {
    const int64_t b = i / cb4, r = i - b * cb4;
    out[b * outCb4 + offCb4 + r] = in[i];
  }
}

// ---- TYoloLayer.getDetections' candidate test (nyololayer.pas:414-418: `if objectness > thresh`) on the device: one thread
// per (image, anchor, cell) reads the objectness plane; a candidate's (classes + 5) entries (stride wh in the planar
// output) are appended to a compact record list.  The host decodes the records exactly as it decodes the full output
// (same floats), so only a few KB instead of the whole output tensor cross PCIe.
//   cand[0] = number of candidates found (may exceed maxCand: the list then holds the first maxCand to arrive)
//   record r at cand + 4 + r*(classes + 9): {tag, image, anchor, cell} as int32, then the classes + 5 entries
__global__ void __launch_bounds__(256) k_yolo_candidates(int64_t total, int n, int entries, int wh, const float *__restrict__ out,
                                                         float thresh, int tag, int maxCand, float *__restrict__ cand) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t % wh);
    const int64_t ba = t / wh;   // image * n + anchor
    const float *p = out + ba * entries * wh + i;
    const float obj = __ldg(p + 4 * (int64_t)wh);
    if (!(obj > thresh)) continue;
    const unsigned slot = atomicAdd(reinterpret_cast<unsigned *>(cand), 1u);
    if (slot >= (unsigned)maxCand) continue;
    float *r = cand + 4 + (int64_t)slot * (entries + 4);
    reinterpret_cast<int *>(r)[0] = tag;
    reinterpret_cast<int *>(r)[1] = (int)(ba / n);
    reinterpret_cast<int *>(r)[2] = (int)(ba % n);
    reinterpret_cast<int *>(r)[3] = i;
    for (int e = 0; e < entries; ++e) r[4 + e] = __ldg(p + (int64_t)e * wh);
  }
}


extern "C" {

int32_t tnn_im2col(tnn_context *ctx, int64_t C, int64_t H, int64_t W, int64_t kh, int64_t kw, int64_t ph, int64_t pw,
                   int64_t sy, int64_t sx, int64_t dy, int64_t dx, const float *im, int64_t imOffset, float *col,
                   int64_t colOffset) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(sy > 0 && sx > 0 && dy > 0 && dx > 0 && kh > 0 && kw > 0, "bad conv geometry");
  const int64_t OH = conv_out(H, ph, kh, sy, dy), OW = conv_out(W, pw, kw, sx, dx);
  const int64_t total = C * kh * kw * OH * OW;
  if (total <= 0) return 0;
  k_im2col<<<grid1d(total), kT, 0, ctx->stream>>>(total, (int)C, (int)H, (int)W, (int)kh, (int)kw, (int)ph, (int)pw,
                                                 (int)sy, (int)sx, (int)dy, (int)dx, (int)OH, (int)OW, im + imOffset,
                                                 col + colOffset);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_col2im(tnn_context *ctx, int64_t C, int64_t H, int64_t W, int64_t kh, int64_t kw, int64_t ph, int64_t pw,
                   int64_t sy, int64_t sx, int64_t dy, int64_t dx, const float *col, int64_t colOffset, float *im,
                   int64_t imOffset) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(sy > 0 && sx > 0 && dy > 0 && dx > 0 && kh > 0 && kw > 0, "bad conv geometry");
  const int64_t OH = conv_out(H, ph, kh, sy, dy), OW = conv_out(W, pw, kw, sx, dx);
  const int64_t total = C * H * W;
  if (total <= 0) return 0;
  k_col2im<<<grid1d(total), kT, 0, ctx->stream>>>(total, (int)C, (int)H, (int)W, (int)kh, (int)kw, (int)ph, (int)pw,
                                                 (int)sy, (int)sx, (int)dy, (int)dx, (int)OH, (int)OW,
                                                 col + colOffset, im + imOffset);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_forward_maxpool(tnn_context *ctx, int64_t aBatch, int64_t outC, int64_t outH, int64_t outW,
                            const float *input, int64_t c, int64_t h, int64_t w, int64_t stride_x, int64_t stride_y,
                            int64_t padding, int64_t kernelSize, uint64_t *indexes, float *output) {
  TNN_CHECK_CTX(ctx);
  (void)c;
  const int64_t total = aBatch * outC * outH * outW;
  if (total <= 0) return 0;
  k_maxpool_fwd<<<grid1d(total), kT, 0, ctx->stream>>>(total, (int)outH, (int)outW, input, (int)h, (int)w,
                                                      (int)stride_x, (int)stride_y, (int)padding, (int)kernelSize,
                                                      (unsigned long long *)indexes, output);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_backward_maxpool(tnn_context *ctx, int64_t aBatch, int64_t outC, int64_t outH, int64_t outW,
                             float *output, const uint64_t *indexes, const float *delta) {
  TNN_CHECK_CTX(ctx);
  const int64_t total = aBatch * outC * outH * outW;
  if (total <= 0) return 0;
  k_maxpool_bwd<<<grid1d(total), kT, 0, ctx->stream>>>(total, output, (const unsigned long long *)indexes, delta);
  TNN_LAUNCHED(ctx);
  return 0;
}

// nncuda.pas:1267-1286: M = batch*channels, out is [M][outHeight*stride][outWidth*stride] where
// outHeight/outWidth are the LOW-resolution sizes (nupsamplelayer.pas:196-201 passes h, w)
int32_t tnn_upsample(tnn_context *ctx, int64_t aBatch, int64_t aChannels, int64_t outHeight, int64_t outWidth,
                     float *in, int64_t stride, int32_t isForward, float scale, float *out, int32_t zeroIn) {
  TNN_CHECK_CTX(ctx);
  const int64_t planes = aBatch * aChannels;
  if (isForward) {
    const int64_t total = planes * outHeight * stride * outWidth * stride;
    if (total <= 0) return 0;
    k_upsample_fwd<<<grid1d(total), kT, 0, ctx->stream>>>(total, (int)outHeight, (int)outWidth, (int)stride, scale, in,
                                                         out);
  } else {
    const int64_t total = planes * outHeight * outWidth;
    if (total <= 0) return 0;
    k_upsample_bwd<<<grid1d(total), kT, 0, ctx->stream>>>(total, (int)outHeight, (int)outWidth, (int)stride, scale, in,
                                                         out, zeroIn);
  }
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_softmax_batch(tnn_context *ctx, int64_t N, const float *input, int64_t iOffset, int64_t batch,
                          int64_t batch_size, int64_t groups, int64_t group_size, int64_t stride, float temp,
                          float *output, int64_t oOffset) {
  TNN_CHECK_CTX(ctx);
  const int64_t rows = batch * groups;
  if (rows <= 0) return 0;
  k_softmax<<<(int)cdiv64(rows, 64), 64, 0, ctx->stream>>>(batch, groups, input + iOffset, N, batch_size, group_size,
                                                          stride, temp, output + oOffset);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_concat_channels(tnn_context *ctx, int64_t batch, int64_t c, int64_t blockSize, const float *in,
                            int64_t outChannels, int64_t cOffset, float *out) {
  TNN_CHECK_CTX(ctx);
  const int64_t cb = c * blockSize, outCb = outChannels * blockSize, offCb = cOffset * blockSize;
  const int64_t total = batch * cb;
  if (total <= 0) return 0;
  const bool vec = (cb % 4 == 0) && (outCb % 4 == 0) && (offCb % 4 == 0) && ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
  if (vec)
    k_concat4<<<grid1d(total / 4), kT, 0, ctx->stream>>>(total / 4, cb / 4, outCb / 4, offCb / 4, (const float4 *)in,
                                                        (float4 *)out);
  else
    k_concat<<<grid1d(total), kT, 0, ctx->stream>>>(total, cb, outCb, offCb, in, out);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_yolo_candidates(tnn_context *ctx, int64_t batch, int64_t n, int64_t classes, int64_t wh, const float *out,
                            float thresh, int32_t tag, int64_t maxCand, float *cand) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(out != nullptr && cand != nullptr && maxCand > 0 && maxCand < (1ll << 30) && wh < (1ll << 31) && n > 0,
              "bad candidate list arguments");
  const int64_t total = batch * n * wh;
  if (total <= 0) return 0;
  k_yolo_candidates<<<grid1d(total), kT, 0, ctx->stream>>>(total, (int)n, (int)(classes + 5), (int)wh, out, thresh, tag,
                                                          (int)maxCand, cand);
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // extern "C"
