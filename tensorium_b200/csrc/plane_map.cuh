// plane_map.cuh — per-channel elementwise driver for NCHW tensors viewed as [planes][blockSize]
// with channel = plane % channels.  One warp per (plane, 512-element chunk) work item, 128-bit
// accesses in the body, scalar head/tail so arbitrary element offsets stay legal.
// Up to two extra inputs and two extra outputs share the indexing of `d`.
#pragma once
#include "common.cuh"

namespace planemap {

constexpr int kThreads = 256;
constexpr int kChunk = 512;

// F: void f(float &d, float in0, float in1, float &out0, float &out1, int c)
template <int NIN, int NOUT, bool VEC, typename F>
__global__ void __launch_bounds__(kThreads)
k_plane(float *d, const float *in0, const float *in1, float *out0, float *out1, int64_t planes, int64_t blockSize,
        int64_t channels, int64_t chunksPerPlane, F f) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t items = planes * chunksPerPlane;
  for (int64_t it = warp0; it < items; it += nwarps) {
    const int64_t plane = it / chunksPerPlane;
    const int64_t chunk = it - plane * chunksPerPlane;
    const int c = (int)(plane % channels);
    const int64_t base = plane * blockSize;
    const int64_t lo = chunk * kChunk;
    const int64_t hi = min(lo + (int64_t)kChunk, blockSize);
    auto scalar_at = [&](int64_t j) {
      const int64_t g = base + j;
      float dv = d[g], o0 = 0.f, o1 = 0.f;
      f(dv, NIN > 0 ? in0[g] : 0.f, NIN > 1 ? in1[g] : 0.f, o0, o1, c);
      d[g] = dv;
      if (NOUT > 0) out0[g] = o0;
      if (NOUT > 1) out1[g] = o1;
    };
    if (!VEC) {
      for (int64_t j = lo + lane; j < hi; j += 32) scalar_at(j);
      continue;
    }
    const int mis = (int)((((uintptr_t)(d + base + lo)) & 15) >> 2);
    const int head = mis ? 4 - mis : 0;
    if (lane < head && lo + lane < hi) scalar_at(lo + lane);
    const int64_t i = lo + head;
    for (int64_t j = i + lane * 4; j + 3 < hi; j += 128) {
      const int64_t g = base + j;
      float4 dv = *reinterpret_cast<float4 *>(d + g);
      float4 a = NIN > 0 ? *reinterpret_cast<const float4 *>(in0 + g) : make_float4(0, 0, 0, 0);
      float4 b = NIN > 1 ? *reinterpret_cast<const float4 *>(in1 + g) : make_float4(0, 0, 0, 0);
      float4 o0 = make_float4(0, 0, 0, 0), o1 = make_float4(0, 0, 0, 0);
      f(dv.x, a.x, b.x, o0.x, o1.x, c);
      f(dv.y, a.y, b.y, o0.y, o1.y, c);
      f(dv.z, a.z, b.z, o0.z, o1.z, c);
      f(dv.w, a.w, b.w, o0.w, o1.w, c);
      *reinterpret_cast<float4 *>(d + g) = dv;
      if (NOUT > 0) *reinterpret_cast<float4 *>(out0 + g) = o0;
      if (NOUT > 1) *reinterpret_cast<float4 *>(out1 + g) = o1;
    }
    if (hi > i) {
      const int64_t t0 = i + (((hi - i) >> 2) << 2);
      if (t0 + lane < hi) scalar_at(t0 + lane);
    }
  }
}

template <int NIN, int NOUT, typename F>
int launch(tnn_context *ctx, float *d, const float *in0, const float *in1, float *out0, float *out1, int64_t total,
           int64_t channels, int64_t batch, F f) {
  if (total <= 0) return 0;
  TNN_REQUIRE(channels > 0 && batch > 0, "bad channel/batch count");
  const int64_t blockSize = total / (channels * batch);
  TNN_REQUIRE(blockSize > 0 && blockSize * channels * batch == total, "tensor sizes do not align");
  const int64_t planes = channels * batch;
  const int64_t chunks = cdiv64(blockSize, kChunk);
  int64_t blocks = cdiv64(planes * chunks, kThreads / 32);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const uintptr_t pd = (uintptr_t)d;
  bool vec = true;
  if (NIN > 0) vec = vec && (((pd ^ (uintptr_t)in0) & 15) == 0);
  if (NIN > 1) vec = vec && (((pd ^ (uintptr_t)in1) & 15) == 0);
  if (NOUT > 0) vec = vec && (((pd ^ (uintptr_t)out0) & 15) == 0);
  if (NOUT > 1) vec = vec && (((pd ^ (uintptr_t)out1) & 15) == 0);
  if (vec)
    k_plane<NIN, NOUT, true, F><<<(int)blocks, kThreads, 0, ctx->stream>>>(d, in0, in1, out0, out1, planes, blockSize,
                                                                          channels, chunks, f);
  else
    k_plane<NIN, NOUT, false, F><<<(int)blocks, kThreads, 0, ctx->stream>>>(d, in0, in1, out0, out1, planes,
                                                                           blockSize, channels, chunks, f);
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // namespace planemap
