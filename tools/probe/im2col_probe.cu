// Standalone probe: TMA im2col-mode loads (cp.async.bulk.tensor.4d...im2col) of an NHWC fp32 tensor, checked on the
// host against the convolution's im2col definition (sim2Col, reference source/ntensors.pas:11415-11491, with the
// channel index innermost).  One load = 128 consecutive output pixels (linearised over n, oh, ow) x 32 channels of
// ONE kernel tap, landed as a K-major SWIZZLE_128B tile (row = pixel, 128 bytes = 32 channels).
//   usage: im2col_probe W H C N k pad stride dil pitch
// Loads: every tap, channel blocks 0 and last, tiles starting at pixel 0, mid-row, across an image boundary and
// past the end of the batch (zero fill expected).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Load { int c0, w0, h0, n0, offw, offh; };
__global__ void k(const __grid_constant__ CUtensorMap tm, float *out, const Load *loads, int nloads, int bytes) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  for (int i = 0; i < nloads; ++i) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int j = threadIdx.x; j < 4096; j += blockDim.x) ((float *)smem)[j] = -7.0f;  // poison
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const Load L = loads[i];
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
          "[%6], {%7, %8};" ::"r"(smem_u32(smem)),
          "l"(reinterpret_cast<uint64_t>(&tm)), "r"(L.c0), "r"(L.w0), "r"(L.h0), "r"(L.n0), "r"(smem_u32(&bar)),
          "h"((uint16_t)L.offw), "h"((uint16_t)L.offh)
          : "memory");
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      }
      asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 4096; j += blockDim.x) out[(size_t)i * 4096 + j] = ((float *)smem)[j];
    __syncthreads();
  }
}
int main(int argc, char **argv) {
  if (argc < 10) { printf("usage: im2col_probe W H C N k pad stride dil pitch\n"); return 1; }
  const int W = atoi(argv[1]), H = atoi(argv[2]), C = atoi(argv[3]), N = atoi(argv[4]), ks = atoi(argv[5]);
  const int pad = atoi(argv[6]), st = atoi(argv[7]), dil = atoi(argv[8]), pitch = atoi(argv[9]);
  const int OH = (H + 2 * pad - (dil * (ks - 1) + 1)) / st + 1, OW = (W + 2 * pad - (dil * (ks - 1) + 1)) / st + 1;
  const long M = (long)N * OH * OW;
  size_t n = (size_t)W * H * N * pitch;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 65521);
  float *d, *out;
  cudaMalloc(&d, n * 4);
  cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q);
  if (!p) { printf("cuTensorMapEncodeIm2col not available\n"); return 1; }
  EncodeIm2colFn enc = (EncodeIm2colFn)p;
  alignas(64) CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * W * 4, (cuuint64_t)pitch * W * H * 4};
  int lower[2] = {-pad, -pad};
  int upper[2] = {pad - (ks - 1) * dil, pad - (ks - 1) * dil};
  cuuint32_t estr[4] = {1, (cuuint32_t)st, (cuuint32_t)st, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, lower, upper, 32, 128, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  if (getenv("CLEAR_BIT21")) reinterpret_cast<uint64_t *>(&tm)[1] &= ~(1llu << 21);
  // the loads
  std::vector<Load> loads;
  std::vector<long> m0s = {0, 37, (long)OH * OW - 50, M - 128 > 0 ? M - 100 : 0};
  std::vector<int> c0s = {0, ((C - 1) / 32) * 32};
  std::vector<long> lm0;
  std::vector<int> ltap;
  for (long m0 : m0s)
    for (int tap = 0; tap < ks * ks; ++tap)
      for (int c0 : c0s) {
        if (m0 < 0 || m0 >= M) continue;
        const int n0 = (int)(m0 / (OH * OW)), rem = (int)(m0 % (OH * OW)), oh0 = rem / OW, ow0 = rem % OW;
        loads.push_back(Load{c0, ow0 * st - pad, oh0 * st - pad, n0, (tap % ks) * dil, (tap / ks) * dil});
        lm0.push_back(m0);
        ltap.push_back(tap);
      }
  Load *dl;
  cudaMalloc(&dl, loads.size() * sizeof(Load));
  cudaMemcpy(dl, loads.data(), loads.size() * sizeof(Load), cudaMemcpyHostToDevice);
  cudaMalloc(&out, loads.size() * 16384);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024);
  k<<<1, 128, 16384 + 1024>>>(tm, out, dl, (int)loads.size(), 16384);
  cudaError_t e = cudaDeviceSynchronize();
  printf("W%d H%d C%d N%d k%d pad%d s%d d%d pitch%d (OH %d OW %d M %ld) %zu loads: %s\n", W, H, C, N, ks, pad, st, dil,
         pitch, OH, OW, M, loads.size(), cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<float> o(loads.size() * 4096);
  cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  long bad = 0, poison = 0;
  for (size_t i = 0; i < loads.size(); ++i) {
    const int tap = ltap[i], kr = tap / ks, kc = tap % ks;
    long lbad = 0;
    for (int row = 0; row < 128; ++row) {
      const long m = lm0[i] + row;
      for (int ci = 0; ci < 32; ++ci) {
        float want = 0.f;
        if (m < M) {
          const int nn = (int)(m / (OH * OW)), rem = (int)(m % (OH * OW)), oh = rem / OW, ow = rem % OW;
          const int ih = oh * st - pad + kr * dil, iw = ow * st - pad + kc * dil, cc = loads[i].c0 + ci;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W && cc < C) want = h[(((size_t)nn * H + ih) * W + iw) * pitch + cc];
        }
        const int chunk = ci >> 2;
        const float got = o[i * 4096 + row * 32 + ((chunk ^ (row & 7)) << 2) + (ci & 3)];
        if (got == -7.0f) ++poison;
        if (got != want) { ++lbad; if (lbad <= 2) printf("   load %zu (m0 %ld tap %d c0 %d) row %d ci %d: got %g want %g\n", i, lm0[i], tap, loads[i].c0, row, ci, got, want); }
      }
    }
    bad += lbad;
  }
  printf("  mismatches: %ld of %zu (unwritten %ld)\n", bad, loads.size() * 4096, poison);
  return bad ? 3 : 0;
}
