// Standalone probe: TMA 4-D tile loads of an NCHW fp32 tensor at arbitrary (possibly unaligned / negative) inner
// coordinates.  usage: tma_probe W H C N TW TH w0 h0 nloads dw   (load i uses w = w0 + (i%3)*dw, c = (i*32)%C)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float *out, int w0, int h0, int nloads, int dw, int C) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[64];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 64; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < nloads; ++i) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(4096) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
              "r"(smem_u32(smem + i * 4096)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(w0 + (i % 3) * dw), "r"(h0), "r"((i * 32) % C), "r"(0),
          "r"(smem_u32(&bar[i]))
          : "memory");
    }
    for (int i = 0; i < nloads; ++i) {
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(&bar[i])), "r"(0) : "memory");
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nloads * 1024; i += blockDim.x) out[i] = ((float *)smem)[i];
}
int main(int argc, char **argv) {
  int W = atoi(argv[1]), H = atoi(argv[2]), C = atoi(argv[3]), N = atoi(argv[4]), TW = atoi(argv[5]), TH = atoi(argv[6]);
  int w0 = atoi(argv[7]), h0 = atoi(argv[8]), nloads = atoi(argv[9]), dw = atoi(argv[10]);
  int sw = argc > 11 ? atoi(argv[11]) : 1;
  size_t n = (size_t)W * H * C * N;
  std::vector<float> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 65521);
  float *d, *out;
  cudaMalloc(&d, n * 4);
  cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, (size_t)nloads * 4096);
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  alignas(64) CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)TW, (cuuint32_t)TH, 32, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, nloads * 4096 + 1024);
  k<<<1, 128, nloads * 4096 + 1024>>>(tm, out, w0, h0, nloads, dw, C);
  cudaError_t e = cudaDeviceSynchronize();
  printf("W%d H%d C%d TW%d TH%d w0=%d h0=%d nloads=%d dw=%d sw=%d: %s\n", W, H, C, TW, TH, w0, h0, nloads, dw, sw, cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<float> o((size_t)nloads * 1024);
  cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  // verify against the expected swizzled layout: [k(32)][pixel(32)], 32B chunks of a 128B row XOR (k&3)
  long bad = 0;
  for (int i = 0; i < nloads; ++i) {
    int w = w0 + (i % 3) * dw, c0 = (i * 32) % C;
    for (int kk = 0; kk < 32; ++kk)
      for (int px = 0; px < 32; ++px) {
        int dh = px / TW, dwp = px % TW;
        int hh = h0 + dh, ww = w + dwp, cc = c0 + kk;
        float want = (hh >= 0 && hh < H && ww >= 0 && ww < W && cc < C) ? h[((size_t)cc * H + hh) * W + ww] : 0.f;
        int chunk = px >> 3;
        int off = kk * 32 + (sw ? ((chunk ^ (kk & 3)) << 3) : (chunk << 3)) + (px & 7);
        if (o[(size_t)i * 1024 + off] != want) ++bad;
      }
  }
  printf("  mismatches: %ld of %d\n", bad, nloads * 1024);
  return 0;
}
