// umma_conv.cuh — tcgen05 (5th-gen tensor core) contraction kernel for sm_100a.
//
// One persistent, warp-specialised kernel template serves the implicit-GEMM convolution forward and plain SGEMM:
//   D[m][n] (TMEM, fp32, 128 lanes x BN columns, double buffered) = sum_k A(m,k) * B(n,k)
//
//   warps 0-3   epilogue   tcgen05.ld (32 lanes x 32 columns) -> scale/bias/activation -> coalesced stores;
//                          overlaps the next tile's main loop through the second TMEM accumulator
//   warps 4-11  producers  read global memory directly (im2col on the fly — no column buffer exists), keep
//                          kPrefetch k-blocks of loads in flight in registers, split fp32 into TF32 hi/lo and
//                          write the UMMA canonical swizzled shared-memory tiles
//   warp 12     MMA        one thread issues tcgen05.mma.cta_group::1.kind::tf32 (3 per k-step in 3xTF32)
//   warp 13     TMA        one thread streams the pre-split weight tiles (cp.async.bulk.tensor, SWIZZLE_128B)
//
// Every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (m fastest, so co-resident CTAs share weights
// in L2); the smem ring and its mbarrier phases run continuously across tiles.
//
// 3xTF32: hi = x with the low 13 mantissa bits cleared (exactly a TF32 number), lo = x - hi (exact in fp32);
// D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi.  TF32 mode stores cvt.rna.tf32 of the operand (the tensor core itself
// would truncate, which biases every product toward zero).
//
// Shared-memory operand layouts (bytes; every tile is 1024B aligned):
//   K-major  (rows = M or N index, 32 floats = 128B per row):   row*128 + ((chunk16 ^ (row&7))<<4)
//            descriptor: SWIZZLE_128B, SBO = 1024 (8 rows), advance 32B per UMMA_K=8 step
//   MN-major (32 consecutive M/N elements = 128B per k row, 4 k rows per 512B atom; the only MN-major layout
//            the hardware takes for 32-bit operands is SWIZZLE_128B_BASE32B = Swizzle<2,5,2>: 32-byte chunks of a
//            row XOR (k&3); measured on B200: plain SWIZZLE_128B MN-major tf32 yields zeros):
//            ((k>>2)*(R/32) + (mn>>5))*512 + (k&3)*128 + ((((mn&31)>>3) ^ (k&3))<<5) + (mn&7)*4
//            descriptor: SWIZZLE_128B_BASE32B, LBO = 512 (next 32-element atom along M/N),
//            SBO = (R/32)*512 (next 4-row k group), advance two SBO per UMMA_K=8 step
#pragma once
#include <type_traits>
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gemm_simt.cuh"

#ifndef TNN_TAIL_TILES
#define TNN_TAIL_TILES 1  // 0 compiles the half-width tail tiles out (A/B builds)
#endif
#ifndef TNN_PIXEL_PREFETCH
#define TNN_PIXEL_PREFETCH 3  // k-blocks of im2col loads kept in flight per producer thread
#endif

#ifndef TNN_MMA_ELECT
#define TNN_MMA_ELECT (lane == 0)
#endif
namespace umma {

constexpr int kNotTaken = -100;
constexpr int BM = 128;
constexpr int BK = 32;
constexpr int kEpilogueWarps = 4;                      // warps 0-3 (warp % 4 = TMEM lane quadrant)
constexpr int kProducerWarps = 8;                      // warps 4-11
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kFirstProducerWarp = kEpilogueWarps;
constexpr int kMmaWarp = kEpilogueWarps + kProducerWarps;  // 12 (+ TMEM alloc/dealloc)
constexpr int kTmaWarp = kMmaWarp + 1;                     // 13
constexpr int kThreads = (kTmaWarp + 1) * 32;              // 448
constexpr int kMaxStages = 6;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA: global (tensor map, coords {c0 = k, c1 = row}) -> swizzled smem tile, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap *tmap, int c0, int c1, int c2, int c3,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
      "[%6];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr), "r"(parity)
      : "memory");
}
// one lane polls the mbarrier, the other 31 park at the warp barrier (no issue slots burnt on spinning)
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}
// same, for waits that are normally long (epilogue waiting for a whole tile, producers waiting for a free slot):
// back off between polls so the polling lane does not steal issue slots from the producer warps of its scheduler
__device__ __forceinline__ void mbar_wait_warp_relaxed(uint64_t *bar, uint32_t parity, unsigned ns) {
  if ((threadIdx.x & 31) == 0) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (true) {
      asm volatile(
          "{\n.reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
          "selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(done)
          : "r"(addr), "r"(parity)
          : "memory");
      if (done) break;
      __nanosleep(ns);
    }
  }
  __syncwarp();
}
// one lane of a CONVERGED warp; unlike `lane == 0` the compiler knows a single thread runs the guarded block, so
// the uniform-datapath operands of UTCMMA / UTMALDG need no per-value "waterfall" loop
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred P1;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P1;\n}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// four columns, no wait: callers batch several and then call tmem_ld_wait()
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), version 1;
// layout_type: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (MN-major 32-bit tiles)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, M=128, N=n
__host__ __device__ inline uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
template <bool SPLIT>
__device__ __forceinline__ void sts_elem(uint8_t *hi, uint8_t *lo, uint32_t off, float v) {
  if (SPLIT) {
    const float h = tf32_hi(v);
    *reinterpret_cast<float *>(hi + off) = h;
    *reinterpret_cast<float *>(lo + off) = v - h;
  } else {
    *reinterpret_cast<float *>(hi + off) = tf32_rna(v);
  }
}
// byte offset of element (mn, k) in an MN-major tile with R rows (R % 32 == 0)
__device__ __forceinline__ uint32_t mn_off(int R, int mn, int k) {
  return (uint32_t)((((k >> 2) * (R >> 5) + (mn >> 5)) << 9) + ((k & 3) << 7) + (((((mn & 31) >> 3) ^ (k & 3))) << 5) +
                    ((mn & 7) << 2));
}
// store one 16-byte chunk (4 consecutive k) of a K-major tile
template <bool SPLIT>
__device__ __forceinline__ void sts_k(uint8_t *hi, uint8_t *lo, int row, int chunk, float4 v) {
  const int off = (row << 7) + ((chunk ^ (row & 7)) << 4);
  if (SPLIT) {
    float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
    *reinterpret_cast<float4 *>(hi + off) = h;
    *reinterpret_cast<float4 *>(lo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  } else {
    *reinterpret_cast<float4 *>(hi + off) = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
  }
}

// ------------------------------------------------------------------ operand loaders
// A loader stages one [R rows][32 k] operand tile per k-block with the 256 producer threads in two phases, so
// global-memory latency overlaps the pipeline (kPrefetch k-blocks of loads are in flight per thread):
//   TileCtx tile_ctx(R, mn0, tid)              per-thread constants of one output tile (row decode etc.)
//   void    load(tctx, regs, kb, tid)          issue ALL global loads of k-block kb into registers
//   void    store<SPLIT>(regs, hi, lo, R, tid) split (3xTF32) and write the swizzled smem tile
//   kMnMajor   smem layout it writes (decides the descriptor);  kTma: tile is written by the TMA warp instead

// matrix whose k index is contiguous: elem(mn,k) = p[mn*ld + k]   (A of NN/NT, B of NT)
struct KContigLoader {
  static constexpr bool kTmaA = false;
  static constexpr bool kMnMajor = false;
  static constexpr bool kTma = false;
  static constexpr int kPrefetch = 1;
  const float *p;
  int64_t ld;
  int64_t rows, K;   // valid extent
  int64_t K1, zfold; // fold-batch-into-k: k = b*K1 + kk -> p + b*zfold (K1 % 32 == 0), else K1 = K
  bool vec_ok;       // rows 16B aligned and K1 % 4 == 0
  struct TileCtx { int R; int64_t mn0; };
  struct Regs { float4 v[8]; };
  __device__ __forceinline__ void fill_lut(int *, int) const {}
  __device__ __forceinline__ TileCtx tile_ctx(int R, int64_t mn0, int, int) const { return TileCtx{R, mn0}; }
  __device__ __forceinline__ void load(TileCtx &c, Regs &r, int64_t kb, int tid, const int *) const {
    const int64_t k0 = kb * BK;
    const int64_t b = k0 / K1;
    const int64_t kk0 = k0 - b * K1;
    const float *base = p + b * zfold;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * kProducerThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < c.R * 8) {
        const int row = e >> 3, chunk = e & 7;
        const int64_t mn = c.mn0 + row;
        const int64_t kk = kk0 + chunk * 4;
        if (mn < rows) {
          const float *q = base + mn * ld + kk;
          if (vec_ok && kk + 3 < K1) {
            v = __ldg(reinterpret_cast<const float4 *>(q));
          } else {
            if (kk + 0 < K1) v.x = __ldg(q + 0);
            if (kk + 1 < K1) v.y = __ldg(q + 1);
            if (kk + 2 < K1) v.z = __ldg(q + 2);
            if (kk + 3 < K1) v.w = __ldg(q + 3);
          }
        }
      }
      r.v[i] = v;
    }
  }
  template <bool SPLIT>
  __device__ __forceinline__ void store(const Regs &r, uint8_t *hi, uint8_t *lo, int R, int tid) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * kProducerThreads;
      if (e < R * 8) sts_k<SPLIT>(hi, lo, e >> 3, e & 7, r.v[i]);
    }
  }
};

// matrix whose mn index is contiguous: elem(mn,k) = p[k*ld + mn]   (A of TN, B of NN/TN)
struct MnContigLoader {
  static constexpr bool kTmaA = false;
  static constexpr bool kMnMajor = true;
  static constexpr bool kTma = false;
  static constexpr int kPrefetch = 1;
  const float *p;
  int64_t ld;
  int64_t rows, K;
  int64_t K1, zfold;
  struct TileCtx { int R, sh; int64_t mn0; };
  struct Regs { float v[32]; };
  __device__ __forceinline__ void fill_lut(int *, int) const {}
  __device__ __forceinline__ TileCtx tile_ctx(int R, int64_t mn0, int, int) const { return TileCtx{R, 31 - __clz(R), mn0}; }
  __device__ __forceinline__ void load(TileCtx &c, Regs &r, int64_t kb, int tid, const int *) const {
    const int64_t k0 = kb * BK;
    const int64_t b = k0 / K1;
    const int64_t kk0 = k0 - b * K1;
    const float *base = p + b * zfold;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int e = tid + i * kProducerThreads;  // R is a power of two >= 32: a warp covers 32 rows of one k
      float v = 0.f;
      if (e < c.R * BK) {
        const int row = e & (c.R - 1), k = e >> c.sh;
        const int64_t mn = c.mn0 + row;
        if (mn < rows && kk0 + k < K1) v = __ldg(base + (kk0 + k) * ld + mn);
      }
      r.v[i] = v;
    }
  }
  template <bool SPLIT>
  __device__ __forceinline__ void store(const Regs &r, uint8_t *hi, uint8_t *lo, int R, int tid) const {
    const int sh = 31 - __clz(R);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int e = tid + i * kProducerThreads;
      if (e < R * BK) sts_elem<SPLIT>(hi, lo, mn_off(R, e & (R - 1), e >> sh), r.v[i]);
    }
  }
};

// implicit im2col: row = output pixel (b, oh, ow) folded over the batch.
// elem = in[b][c][oh*sy - ph + kr*dy][ow*sx - pw + kc*dx] or 0     (sim2Col ntensors.pas:11415-11491)
// The contraction index is re-ordered (the weight copy is ours, see k_prepare_weights) as
//     k = ((cg*kh*kw + tap)*16 + ci),   c = cg*16 + ci,   tap = kr*kw + kc
// so the 16 consecutive k one thread stages per k-block share ONE kernel tap and walk 16 channels: the
// in-bounds test is one bit of a per-tile tap mask, and each load is `pointer += H*W; predicated LDG`.
// Channels are padded to a multiple of 16 (zero weights, loads predicated off).
struct Im2colPixelLoader {
  static constexpr bool kTmaA = false;
  static constexpr bool kMnMajor = true;
  static constexpr bool kTma = false;
  static constexpr int kPrefetch = TNN_PIXEL_PREFETCH;
  const float *in;
  simt::ConvGeom g;
  int64_t rows;  // batch*OH*OW
  int HW, kk;    // H*W, kh*kw
  struct TileCtx {
    const float *base;           // &in[b][0][ih0][iw0]
    unsigned long long tapmask;  // bit (kr*kw+kc): tap inside the image for this pixel (0 if the row is padding)
    int cg, tap;                 // (channel group, tap) of the NEXT k-block this thread loads (walks forward)
  };
  struct Regs { float v[16]; };
  // lut[t] = element offset of kernel tap t inside a channel plane
  __device__ __forceinline__ void fill_lut(int *lut, int tid_all) const {
    if (tid_all < kk) {
      const int kr = tid_all / g.kw, kc = tid_all - kr * g.kw;
      lut[tid_all] = kr * g.dy * g.W + kc * g.dx;
    }
  }
  __device__ __noinline__ TileCtx tile_ctx(int, int64_t mn0, int tid, int kb0) const {
    TileCtx c;
    {
      const int gi = kb0 * 2 + (tid >> 7);
      c.cg = gi / kk;
      c.tap = gi - c.cg * kk;
    }
    const int row = tid & 127;
    const int64_t mn = mn0 + row;
    const bool valid = mn < rows;
    const int P = g.OH * g.OW;
    const int pix = valid ? (int)mn : 0;
    const int b = pix / P, p = pix - b * P;
    const int oh = p / g.OW, ow = p - oh * g.OW;
    const int ih0 = oh * g.sy - g.ph, iw0 = ow * g.sx - g.pw;
    c.base = in + (int64_t)b * g.inImg + (int64_t)ih0 * g.W + iw0;
    unsigned long long m = 0;
    if (valid) {
      int t = 0;
      for (int kr = 0; kr < g.kh; ++kr)
        for (int kc = 0; kc < g.kw; ++kc, ++t)
          if ((unsigned)(ih0 + kr * g.dy) < (unsigned)g.H && (unsigned)(iw0 + kc * g.dx) < (unsigned)g.W) m |= 1ull << t;
    }
    c.tapmask = m;
    return c;
  }
  // loads the thread's 16 channels of the current (cg, tap) and advances the walk by one k-block (two groups)
  __device__ __forceinline__ void load(TileCtx &cx, Regs &r, int64_t, int, const int *lut) const {
    const int c0 = cx.cg * 16;
    const bool on = (cx.tapmask >> cx.tap) & 1ull;
    const float *q = cx.base + (int64_t)c0 * HW + lut[cx.tap];
    const int nc = on ? g.C - c0 : 0;  // channels of this group that exist (may exceed 16, may be <= 0)
    const uint32_t hw = (uint32_t)HW;
    uint32_t off = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      r.v[j] = (j < nc) ? __ldg(q + off) : 0.f;
      off += hw;
    }
    cx.tap += 2;
    if (cx.tap >= kk) { cx.tap -= kk; ++cx.cg; if (cx.tap >= kk) { cx.tap -= kk; ++cx.cg; } }
  }
  template <bool SPLIT>
  __device__ __forceinline__ void store(const Regs &r, uint8_t *hi, uint8_t *lo, int, int tid) const {
    const int row = tid & 127, khalf = tid >> 7;
    // mn_off for R = 128 and k = khalf*16 + j: per-thread base + compile-time terms
    const uint32_t rowpart = (uint32_t)((row >> 5) << 9) + (uint32_t)((row & 7) << 2) + (uint32_t)khalf * 4u * 2048u;
    uint32_t soff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) soff[q] = rowpart + (uint32_t)(q << 7) + (uint32_t)(((((row & 31) >> 3) ^ q)) << 5);
#pragma unroll
    for (int j = 0; j < 16; ++j) sts_elem<SPLIT>(hi, lo, soff[j & 3] + (uint32_t)((j >> 2) * 2048), r.v[j]);
  }
};

// k extent of one pipeline stage: 32 unless the A loader says otherwise
template <typename L, typename = void>
struct StageK { static constexpr int v = BK; };
template <typename L>
struct StageK<L, std::void_t<decltype(L::kKB)>> { static constexpr int v = L::kKB; };

// Same gather with 16-k stages: one (channel group, tap) per stage, each thread stages 8 of its 16 channels.  The
// A tile is k-outermost (2 KB per 4 k x 128 pixels), so a 16-k stage is simply the first half of the 32-k layout;
// half-size stages put four of them in flight where 3xTF32 at BN = 256 only has room for two 32-k stages.
struct Im2colPixelLoader16 : Im2colPixelLoader {
  static constexpr int kKB = 16;
  static constexpr int kPrefetch = 2 * TNN_PIXEL_PREFETCH;
  struct Regs { float v[8]; };
  __device__ __noinline__ TileCtx tile_ctx(int r, int64_t mn0, int tid, int kb0) const {
    TileCtx c = Im2colPixelLoader::tile_ctx(r, mn0, tid, 0);
    c.cg = kb0 / kk;
    c.tap = kb0 - c.cg * kk;
    return c;
  }
  __device__ __forceinline__ void load(TileCtx &cx, Regs &r, int64_t, int tid, const int *lut) const {
    const int c0 = cx.cg * 16 + (tid >> 7) * 8;
    const bool on = (cx.tapmask >> cx.tap) & 1ull;
    const float *q = cx.base + (int64_t)c0 * HW + lut[cx.tap];
    const int nc = on ? g.C - c0 : 0;
    const uint32_t hw = (uint32_t)HW;
    uint32_t off = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r.v[j] = (j < nc) ? __ldg(q + off) : 0.f;
      off += hw;
    }
    if (++cx.tap == kk) { cx.tap = 0; ++cx.cg; }
  }
  template <bool SPLIT>
  __device__ __forceinline__ void store(const Regs &r, uint8_t *hi, uint8_t *lo, int, int tid) const {
    const int row = tid & 127, khalf = tid >> 7;
    const uint32_t rowpart = (uint32_t)((row >> 5) << 9) + (uint32_t)((row & 7) << 2) + (uint32_t)khalf * 2u * 2048u;
    uint32_t soff[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) soff[q] = rowpart + (uint32_t)(q << 7) + (uint32_t)(((((row & 31) >> 3) ^ q)) << 5);
#pragma unroll
    for (int j = 0; j < 8; ++j) sts_elem<SPLIT>(hi, lo, soff[j & 3] + (uint32_t)((j >> 2) * 2048), r.v[j]);
  }
};

// weights pre-split into a zero-padded K-major copy [F][Kpad] (hi and lo planes) and fetched by TMA
struct TmaWeightsLoader {
  static constexpr bool kTmaA = false;
  static constexpr bool kMnMajor = false;
  static constexpr bool kTma = true;
  static constexpr int kPrefetch = 3;
  struct TileCtx {};
  struct Regs {};
  __device__ __forceinline__ void fill_lut(int *, int) const {}
  __device__ __forceinline__ TileCtx tile_ctx(int, int64_t, int, int) const { return TileCtx{}; }
  __device__ __forceinline__ void load(TileCtx &, Regs &, int64_t, int, const int *) const {}
  template <bool SPLIT>
  __device__ __forceinline__ void store(const Regs &, uint8_t *, uint8_t *, int, int) const {}
};

// TF32 mode: the producers, not the tensor pipe, are the bound of the wide layers, and the general load form costs
// ~8 instructions per element.  Full channel groups of taps inside the image take an unpredicated path with one
// pointer bump per load: 0.438 -> 0.400 ms at 256->512 @26^2.  (In 3xTF32 the same code LOSES 6 %, and its mere
// presence costs 2 % — hence a separate loader type that only the TF32 instantiation uses.)
struct Im2colPixelLoaderFast : Im2colPixelLoader {
  __device__ __forceinline__ void load(TileCtx &cx, Regs &r, int64_t, int, const int *lut) const {
    const int c0 = cx.cg * 16;
    const bool on = (cx.tapmask >> cx.tap) & 1ull;
    const float *q = cx.base + (int64_t)c0 * HW + lut[cx.tap];
    const int nc = on ? g.C - c0 : 0;
    if (nc >= 16) {
      const int64_t hw = HW;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        r.v[j] = __ldg(q);
        q += hw;
      }
    } else {
      const uint32_t hw = (uint32_t)HW;
      uint32_t off = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        r.v[j] = (j < nc) ? __ldg(q + off) : 0.f;
        off += hw;
      }
    }
    cx.tap += 2;
    if (cx.tap >= kk) { cx.tap -= kk; ++cx.cg; if (cx.tap >= kk) { cx.tap -= kk; ++cx.cg; } }
  }
};

// Activations fetched by TMA (stride-1 convolutions whose rows are 16-byte aligned and whose kernel columns
// shift by multiples of 4 pixels — in practice the 1x1 convolutions; see conv_patch for the measured reason).  The 128 rows of a tile are
// four independent "atoms" of 32 output pixels, each a TH x TW patch (TH*TW = 32) of one image; for kernel tap
// (kr,kc) the im2col rows of an atom are the same patch of the INPUT shifted by (kr*dy - ph, kc*dx - pw), which
// one cp.async.bulk.tensor.4d box {TW, TH, 32, 1} of the NCHW tensor {W, H, C, N} delivers — padding comes for
// free as the TMA's out-of-bounds zero fill, and the box lands as [channel k][pixel], i.e. a run of the 512-byte
// (4 k rows x 32 pixels) MN-major atoms (SWIZZLE_128B_ATOM_32B == the UMMA SWIZZLE_128B_BASE32B pattern).  A k-block is
// 32 channels of ONE tap:  k = (tap*CB + cb)*32 + ci,  c = cb*32 + ci.  The atoms of a tile are 4 KB apart, so
// the descriptor uses LBO = 4096 (next atom along M) and SBO = 512 (next 4-row k group).
// The producer warps no longer gather: they only derive the 3xTF32 low part (or round to TF32) in shared memory.
struct TmaPatchLoader {
  static constexpr bool kTmaA = true;
  static constexpr bool kMnMajor = true;
  static constexpr bool kTma = false;
  static constexpr int kPrefetch = 1;
  int TW, TH, WT, HT;   // patch shape; patches per image row / column
  int atomsPerImg;      // WT*HT
  int numAtoms;         // batch*atomsPerImg
  int batch;
  int ph, pw, kw, dy, dx;
  int CB;               // 32-channel blocks per tap
  struct TileCtx {};
  struct Regs {};
  __device__ __forceinline__ void fill_lut(int *, int) const {}
  __device__ __forceinline__ TileCtx tile_ctx(int, int64_t, int, int) const { return TileCtx{}; }
  __device__ __forceinline__ void load(TileCtx &, Regs &, int64_t, int, const int *) const {}
  template <bool SPLIT>
  __device__ __forceinline__ void store(const Regs &, uint8_t *, uint8_t *, int, int) const {}
  // image and top-left output pixel of an atom (atoms past the end map to image `batch`: the TMA zero-fills)
  __device__ __forceinline__ void atom_origin(int atom, int &n, int &oh0, int &ow0) const {
    n = atom / atomsPerImg;
    const int r = atom - n * atomsPerImg;
    const int hy = r / WT;
    oh0 = hy * TH;
    ow0 = (r - hy * WT) * TW;
    if (atom >= numAtoms) n = batch;
  }
};

// 16-channel stages (CB counts 16-channel blocks): atoms of 16 k x 32 pixels, 2 KB apart
struct TmaPatchLoader16 : TmaPatchLoader {
  static constexpr int kKB = 16;
};

// ------------------------------------------------------------------ epilogues
// activations other than linear / leaky: one out-of-line copy (the switch is large; inlining it into the unrolled
// epilogue blew the kernel up to 300 KB of SASS)
static __device__ __noinline__ float act_out_of_line(float v, int a) { return tnn_act(v, a); }

// Row context per thread (its TMEM lane = row m), then one call per column.
// EXTRA = false: the lean instantiation for plain forward convolutions (no residual, no accumulation) — the kernel
// is sensitive to epilogue code size, so the rarely used paths live in their own instantiation
template <bool EXTRA>
struct ConvEpilogueT {  // out[b][f][p] = act(acc*scale[f] + bias[f])
  float *out;
  const float *scale, *bias;
  int P;
  int64_t outImg;
  int act;
  int64_t rows, cols;
  bool accumulate;  // out += value (dgrad adds into the upstream delta, like the reference's col2im)
  // TW != 0: rows are patches of TH x TW output pixels (1 << ashift rows each: 32-row TmaPatchLoader atoms or the
  // 128-row patches of the halo kernel), else row m = pixel m of the batch-folded image list
  int TW, TH, WT, atomsPerImg, OH, OW, batch, ashift;
  const float *residual;  // optional: added after the activation (the shortcut layer that follows), same shape as out
  struct Row { float *q; };
  __device__ __forceinline__ Row row(int64_t m, bool &valid) const {
    if (TW != 0) {
      const int atom = (int)(m >> ashift), j = (int)(m & ((1 << ashift) - 1));
      const int n = atom / atomsPerImg;
      const int r = atom - n * atomsPerImg;
      const int hy = r / WT;
      const int jh = j / TW;
      const int oh = hy * TH + jh, ow = (r - hy * WT) * TW + (j - jh * TW);
      valid = n < batch && oh < OH && ow < OW;
      return Row{valid ? out + (int64_t)n * outImg + oh * OW + ow : out};
    }
    valid = m < rows;
    const int64_t mm = valid ? m : 0;
    const int b = (int)(mm / P), p = (int)(mm - (int64_t)b * P);
    return Row{out + (int64_t)b * outImg + p};
  }
  // 32 consecutive columns starting at n; `cnt` of them are valid.  Each lane fetches the scale/bias of ONE column
  // (coalesced) into the warp's 64-float shared scratch; the unrolled loop reads them back four at a time as
  // broadcast LDS.128 (shuffles cost two MIO slots per column, the pipe the stores need).  The activation switch
  // stays outside the loops.
  __device__ __forceinline__ void store_chunk(const Row &r, int64_t n, const uint32_t (&acc)[32], int cnt,
                                              bool rowvalid, float *scratch) const {
    const int lane = threadIdx.x & 31;
    __syncwarp();  // the previous chunk's reads of the scratch are done
    scratch[lane] = (scale && lane < cnt) ? __ldg(scale + n + lane) : 1.0f;
    scratch[32 + lane] = (bias && lane < cnt) ? __ldg(bias + n + lane) : 0.0f;
    __syncwarp();
    float *q = r.q + n * P;
    const float *rq = residual ? residual + (q - out) : nullptr;
    const bool leaky = act == TNN_ACT_LEAKY || act == TNN_ACT_REVLEAKY;
    const bool plain = !EXTRA || (rq == nullptr && !accumulate);  // the common case gets a loop without loads
    if ((leaky || act == TNN_ACT_LINEAR) && plain && rowvalid && cnt == 32) {
      // the hot loop: full chunk, nothing to read — FMUL, FADD, FMUL, FMNMX, STG and the pointer bump per column
      // (leaky(v) = max(v, 0.1 v); linear: max(v, v))
      const float neg = leaky ? 0.1f : 1.0f;
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) {
        const float4 s4 = *reinterpret_cast<const float4 *>(scratch + j4);
        const float4 b4 = *reinterpret_cast<const float4 *>(scratch + 32 + j4);
        const float s[4] = {s4.x, s4.y, s4.z, s4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float v = __fadd_rn(__fmul_rn(__uint_as_float(acc[j4 + u]), s[u]), b[u]);
          *q = fmaxf(v, __fmul_rn(neg, v));
          q += P;
        }
      }
    } else if (EXTRA && (leaky || act == TNN_ACT_LINEAR)) {
      // residual add (the shortcut that follows) and/or accumulation into the output (dgrad): fetch the operands of
      // 16 columns FIRST so the loads are in flight together — two memory round trips per chunk instead of 32
      const float neg = leaky ? 0.1f : 1.0f;
      const bool both = rq != nullptr && accumulate;
#pragma unroll
      for (int h = 0; h < 32; h += 16) {
        float ex[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool on = rowvalid && (h + j) < cnt;
          const float *src = rq ? rq + (int64_t)(h + j) * P : q + (int64_t)(h + j) * P;
          ex[j] = (on && (rq || accumulate)) ? *src : 0.f;   // the residual if there is one, else the old output
        }
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const float4 s4 = *reinterpret_cast<const float4 *>(scratch + h + j4);
          const float4 b4 = *reinterpret_cast<const float4 *>(scratch + 32 + h + j4);
          const float s[4] = {s4.x, s4.y, s4.z, s4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = h + j4 + u;
            float v = __fadd_rn(__fmul_rn(__uint_as_float(acc[j]), s[u]), b[u]);
            v = fmaxf(v, __fmul_rn(neg, v));
            v = __fadd_rn(v, ex[j4 + u]);                       // + residual, or + old output
            if (rowvalid && j < cnt) {
              float *dst = q + (int64_t)j * P;
              *dst = both ? __fadd_rn(*dst, v) : v;             // residual AND accumulate: re-read the old value
            }
          }
        }
      }
    } else {
#pragma unroll
      for (int j4 = 0; j4 < 32; j4 += 4) {
        const float4 s4 = *reinterpret_cast<const float4 *>(scratch + j4);
        const float4 b4 = *reinterpret_cast<const float4 *>(scratch + 32 + j4);
        const float s[4] = {s4.x, s4.y, s4.z, s4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j4 + u;
          const float v = __fadd_rn(__fmul_rn(__uint_as_float(acc[j]), s[u]), b[u]);
          if (rowvalid && j < cnt) {
            float o = act_out_of_line(v, act);
            if (rq) o = __fadd_rn(o, __ldg(rq + (int64_t)j * P));
            q[(int64_t)j * P] = o;
          }
        }
      }
    }
  }
};
struct GemmEpilogue {  // C[m][n] = alpha*acc + beta*C
  float *C;
  int64_t ldc;
  float alpha, beta;
  int64_t rows, cols;
  struct Row { float *q; };
  __device__ __forceinline__ Row row(int64_t m, bool &valid) const {
    valid = m < rows;
    return Row{C + (valid ? m : 0) * ldc};
  }
  __device__ __forceinline__ void store_chunk(const Row &r, int64_t n, const uint32_t (&acc)[32], int cnt,
                                              bool rowvalid, float *) const {
    if (!rowvalid) return;
    float *q = r.q + n;
    if (beta == 0.f) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cnt) q[j] = alpha * __uint_as_float(acc[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cnt) q[j] = fmaf(alpha, __uint_as_float(acc[j]), beta * q[j]);
    }
  }
};
struct GemmAtomicEpilogue {  // split-k: C += alpha*acc
  float *C;
  int64_t ldc;
  float alpha, beta;
  int64_t rows, cols;
  struct Row { float *q; };
  __device__ __forceinline__ Row row(int64_t m, bool &valid) const {
    valid = m < rows;
    return Row{C + (valid ? m : 0) * ldc};
  }
  __device__ __forceinline__ void store_chunk(const Row &r, int64_t n, const uint32_t (&acc)[32], int cnt,
                                              bool rowvalid, float *) const {
    if (!rowvalid) return;
    float *q = r.q + n;
    if (cnt == 32 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
      // 16-byte vector reductions: a quarter of the instructions and of the L2 reduction operations (each lane owns
      // a different row, so every scalar red is its own sector operation)
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + j), "f"(alpha * __uint_as_float(acc[j])),
                     "f"(alpha * __uint_as_float(acc[j + 1])), "f"(alpha * __uint_as_float(acc[j + 2])),
                     "f"(alpha * __uint_as_float(acc[j + 3]))
                     : "memory");
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < cnt) atomicAdd(q + j, alpha * __uint_as_float(acc[j]));
  }
};

// ------------------------------------------------------------------ the kernel
struct SharedCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t araw[kMaxStages];  // TmaPatchLoader: the raw fp32 A tile of the stage has landed
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  int lut[64];     // A-loader scratch (per-tap offsets of the im2col gather)
  int lut_b[192];  // B-loader scratch (wgrad: per-tap {offset, dh, dw})
  alignas(16) float epi_sb[kEpilogueWarps][64];  // per epilogue warp: scale[32] | bias[32] of the current chunk
};

struct TileSched {  // tiles are numbered m fastest, then n, then k-split (32-bit: divisions stay cheap and small)
  int mTiles, nTiles, kSplits, kBlocks, kBlocksPerSplit, numTiles;
  // Tiles [wideTiles, numTiles) are half-width: when the last wave would fill less than half of the SMs, each of its
  // tiles is split in two along n (TMA-fed B only), so that wave costs half a tile instead of a whole one.
  int wideTiles;
  __device__ __forceinline__ int half(int t) const { return (TNN_TAIL_TILES && t >= wideTiles) ? 1 + ((t - wideTiles) & 1) : 0; }
  __device__ __noinline__ void decode(int t, int &mt, int &nt, int &kb0, int &kb1) const {
    if (TNN_TAIL_TILES && t >= wideTiles) t = wideTiles + ((t - wideTiles) >> 1);
    const unsigned r = (unsigned)t / (unsigned)mTiles;
    mt = t - (int)r * mTiles;
    const unsigned ks = r / (unsigned)nTiles;
    nt = (int)r - (int)ks * nTiles;
    kb0 = (int)ks * kBlocksPerSplit;
    kb1 = min(kBlocks, kb0 + kBlocksPerSplit);
  }
};

// BN = runtime tile width (multiple of 16, <= 256; a power of two >= 32 when B is MN-major).  SPLIT = 3xTF32.
// accCols = TMEM columns per accumulator buffer (2 buffers are allocated).  tmBhi/tmBlo: tensor maps of the
// pre-split weights (LB::kTma only).
template <bool SPLIT, typename LA, typename LB, typename EPI>
__global__ void __launch_bounds__(kThreads, 1)
k_umma(LA la, LB lb, EPI epi, TileSched sched, int BN, int stages, uint32_t accCols,
       const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
       const __grid_constant__ CUtensorMap tmA) {
  static_assert(!LA::kTmaA || LB::kTma, "TMA activations are paired with TMA weights");
  extern __shared__ uint8_t smem_raw[];
  __shared__ SharedCtl ctl;
  // 1024B-aligned operand ring
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the shared address space
  // and emits LDS / STS instead of generic LD / ST for every access below)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int KB = StageK<LA>::v;
  static_assert(KB == BK || (LB::kTma && LA::kMnMajor), "16-k stages: MN-major A (gathered or TMA patches) + TMA weights only");
  const int aBytes = BM * KB * 4;
  const int bBytes = BN * KB * 4;
  const int stageBytes = (aBytes + bBytes) * (SPLIT ? 2 : 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&ctl.full[s], kProducerWarps + (LB::kTma ? 1 : 0));
      mbar_init(&ctl.empty[s], 1);
      mbar_init(&ctl.araw[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl.tmem_full[s], 1);
      mbar_init(&ctl.tmem_empty[s], kEpilogueWarps);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&ctl.tmem_base, accCols * 2);
  la.fill_lut(ctl.lut, threadIdx.x);
  lb.fill_lut(ctl.lut_b, threadIdx.x);
  if (LB::kTma && warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmBhi);
    if (SPLIT) tma_prefetch_desc(&tmBlo);
    if (LA::kTmaA) tma_prefetch_desc(&tmA);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl.tmem_base;

  if (LA::kTmaA && warp >= kFirstProducerWarp && warp < kMmaWarp) {
    // ===================== TMA-fed A: derive the operand the tensor core needs, in shared memory ===========
    // 3xTF32: the raw fp32 tile IS the high operand (kind::tf32 reads the top 19 bits); write lo = x - trunc(x).
    // TF32: round the tile in place (cvt.rna) so products are not biased toward zero.
    const int tid = threadIdx.x - kFirstProducerWarp * 32;
    int64_t remaining = 0;
    for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x) {
      int mt, nt, k0, k1;
      sched.decode(t, mt, nt, k0, k1);
      remaining += k1 - k0;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (; remaining > 0; --remaining) {
      mbar_wait_warp(&ctl.araw[stage], phase);
      uint8_t *sa = smem + (size_t)stage * stageBytes;
      uint8_t *sa_lo = sa + aBytes;
#pragma unroll
      for (int i = 0; i < aBytes / (16 * kProducerThreads); ++i) {
        const int off = (i * kProducerThreads + tid) * 16;
        const float4 v = *reinterpret_cast<const float4 *>(sa + off);
        if (SPLIT) {
          *reinterpret_cast<float4 *>(sa_lo + off) =
              make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
        } else {
          *reinterpret_cast<float4 *>(sa + off) = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.full[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= kFirstProducerWarp && warp < kMmaWarp) {
    // ===================== producers: global -> registers -> (split) -> swizzled smem =====================
    constexpr int D = LA::kPrefetch;
    const int tid = threadIdx.x - kFirstProducerWarp * 32;
    typename LA::Regs ra[D];
    typename LB::Regs rb[D];
    // load cursor (runs D k-blocks ahead of the store cursor)
    int ltile = blockIdx.x, lkb = 0, lkb1 = 0, lmt = 0, lnt = 0;
    typename LA::TileCtx ca;
    typename LB::TileCtx cb;
    bool lvalid = ltile < sched.numTiles;
    if (lvalid) {
      sched.decode(ltile, lmt, lnt, lkb, lkb1);
      ca = la.tile_ctx(BM, (int64_t)lmt * BM, tid, lkb);
      cb = lb.tile_ctx(BN, (int64_t)lnt * BN, tid, lkb);
    }
    auto issue_load = [&](typename LA::Regs &xa, typename LB::Regs &xb) {
      if (!lvalid) return;
      la.load(ca, xa, lkb, tid, ctl.lut);
      if (!LB::kTma) lb.load(cb, xb, lkb, tid, ctl.lut_b);
      if (++lkb == lkb1) {
        ltile += gridDim.x;
        lvalid = ltile < sched.numTiles;
        if (lvalid) {
          sched.decode(ltile, lmt, lnt, lkb, lkb1);
          ca = la.tile_ctx(BM, (int64_t)lmt * BM, tid, lkb);
          cb = lb.tile_ctx(BN, (int64_t)lnt * BN, tid, lkb);
        }
      }
    };
    // store cursor
    int64_t remaining = 0;  // k-blocks this CTA stages in total
    for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x) {
      int mt, nt, k0, k1;
      sched.decode(t, mt, nt, k0, k1);
      remaining += k1 - k0;
    }
    int stage = 0;
    uint32_t phase = 0;
    int prime = D;  // the first D passes only issue loads (fills the register slots)
    while (remaining > 0) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (remaining > 0) {
          if (prime == 0) {
            mbar_wait_warp(&ctl.empty[stage], phase ^ 1);
            uint8_t *sa = smem + (size_t)stage * stageBytes;
            uint8_t *sa_lo = sa + aBytes;
            uint8_t *sb = sa + (SPLIT ? 2 : 1) * aBytes;
            uint8_t *sb_lo = sb + bBytes;
            la.template store<SPLIT>(ra[d], sa, sa_lo, BM, tid);
            if (!LB::kTma) lb.template store<SPLIT>(rb[d], sb, sb_lo, BN, tid);
            fence_proxy_async();  // make generic-proxy smem writes visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctl.full[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1; }
            --remaining;
          } else {
            --prime;
          }
          issue_load(ra[d], rb[d]);  // (re)fill this register slot: D k-blocks of loads stay in flight
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (one elected thread) =====================
    const uint32_t smem_base = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x, ++it) {
      int mt, nt, kb0, kb1;
      sched.decode(t, mt, nt, kb0, kb1);
      const uint32_t idesc = make_idesc(sched.half(t) ? BN >> 1 : BN, LA::kMnMajor, LB::kMnMajor);
      const int as = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait_warp(&ctl.tmem_empty[as], (use & 1) ^ 1);  // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_warp(&ctl.full[stage], phase);
        tc_fence_after();
        if (TNN_MMA_ELECT) {
          const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
          const uint32_t sa_lo = sa + aBytes;
          const uint32_t sb = sa + (SPLIT ? 2 : 1) * aBytes;
          const uint32_t sb_lo = sb + bBytes;
#pragma unroll
          for (int ks = 0; ks < KB / 8; ++ks) {
            // MN-major: LBO 512, SBO (R/32)*512, one UMMA_K=8 step = two 4-row k groups.  K-major: SBO 1024, +32B.
            const uint32_t aoff = LA::kTmaA ? ks * 1024 : LA::kMnMajor ? ks * 2 * (BM / 32) * 512 : ks * 32;
            const uint32_t boff = LB::kMnMajor ? ks * 2 * (BN / 32) * 512 : ks * 32;
            const uint32_t a_lbo = LA::kTmaA ? KB * 128 : LA::kMnMajor ? 512 : 16;
            const uint32_t a_sbo = LA::kTmaA ? 512 : LA::kMnMajor ? (BM / 32) * 512 : 1024;
            // K-major B rows are 128 B (SWIZZLE_128B) or, with 16-k stages, 64 B (SWIZZLE_64B: 8-row groups of 512 B)
            const uint32_t b_lbo = LB::kMnMajor ? 512 : 16, b_sbo = LB::kMnMajor ? (BN / 32) * 512 : (KB == BK ? 1024 : 512);
            const uint32_t a_lt = LA::kMnMajor ? 1u : 2u, b_lt = LB::kMnMajor ? 1u : (KB == BK ? 2u : 4u);
            const uint64_t ah = make_desc(sa + aoff, a_lbo, a_sbo, a_lt);
            const uint64_t bh = make_desc(sb + boff, b_lbo, b_sbo, b_lt);
            const uint32_t acc = (kb != kb0 || ks != 0) ? 1u : 0u;
            if (SPLIT) {
              const uint64_t al = make_desc(sa_lo + aoff, a_lbo, a_sbo, a_lt);
              const uint64_t bl = make_desc(sb_lo + boff, b_lbo, b_sbo, b_lt);
              mma_tf32(tmem_d, al, bh, idesc, acc);  // small terms first
              mma_tf32(tmem_d, ah, bl, idesc, 1u);
              mma_tf32(tmem_d, ah, bh, idesc, 1u);
            } else {
              mma_tf32(tmem_d, ah, bh, idesc, acc);
            }
          }
          mma_commit(&ctl.empty[stage]);                     // frees the smem slot when the MMAs have read it
          if (kb == kb1 - 1) mma_commit(&ctl.tmem_full[as]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kTmaWarp) {
    // ===================== TMA producer of the weight tiles (one thread) =====================
    if (LB::kTma && lane == 0) {
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x) {
        int mt, nt, kb0, kb1;
        sched.decode(t, mt, nt, kb0, kb1);
        const int n0 = nt * BN + (sched.half(t) == 2 ? BN >> 1 : 0);
        int an[4], ah[4], aw[4];  // TmaPatchLoader: image and input origin (tap 0) of the tile's four atoms
        int tap = 0, cb = 0, kr = 0, kc = 0;
        if constexpr (LA::kTmaA) {
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            la.atom_origin(mt * 4 + a, an[a], ah[a], aw[a]);
            ah[a] -= la.ph;
            aw[a] -= la.pw;
          }
          tap = kb0 / la.CB;
          cb = kb0 - tap * la.CB;
          kr = tap / la.kw;
          kc = tap - kr * la.kw;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&ctl.empty[stage], phase ^ 1);
          if constexpr (LA::kTmaA) {
            const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
            mbar_arrive_expect_tx(&ctl.araw[stage], (uint32_t)aBytes);
#pragma unroll
            for (int a = 0; a < 4; ++a)
              tma_load_4d(sa + a * (KB * 128), &tmA, aw[a] + kc * la.dx, ah[a] + kr * la.dy, cb * KB, an[a], &ctl.araw[stage]);
            if (++cb == la.CB) {
              cb = 0;
              if (++kc == la.kw) { kc = 0; ++kr; }
            }
          }
          const uint32_t sb = smem_base + (uint32_t)stage * stageBytes + (SPLIT ? 2 : 1) * aBytes;
          mbar_arrive_expect_tx(&ctl.full[stage], (uint32_t)bBytes * (SPLIT ? 2 : 1));
          // a half-width tile loads the full box from its own first row: the MMA reads the first BN/2 rows of it
          tma_load_2d(sb, &tmBhi, (int)(kb * KB), n0, &ctl.full[stage]);
          if (SPLIT) tma_load_2d(sb + bBytes, &tmBlo, (int)(kb * KB), n0, &ctl.full[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps 0-3: TMEM -> registers -> global =====================
    const int quad = warp;  // == warp % 4: the TMEM lane quadrant this warp may read
    int it = 0;
    for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x, ++it) {
      int mt, nt, kb0, kb1;
      sched.decode(t, mt, nt, kb0, kb1);
      const int as = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      const int64_t m = (int64_t)mt * BM + quad * 32 + lane;
      const int hf = sched.half(t);
      const int bn = hf ? BN >> 1 : BN;  // width of this tile
      const int64_t n0 = (int64_t)nt * BN + (hf == 2 ? BN >> 1 : 0);
      bool mvalid;
      typename EPI::Row rc = epi.row(m, mvalid);
      mbar_wait_warp_relaxed(&ctl.tmem_full[as], use & 1, 64);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols + ((uint32_t)(quad * 32) << 16);
      const int nChunks = (bn + 31) >> 5;  // 32-column chunks; chunk i starts at column 32*i (< accCols)
      for (int ci = 0; ci < nChunks; ++ci) {
        const int c0 = ci * 32;
        uint32_t r[32];
        tmem_ld32(tmem_d + (uint32_t)c0, r);
        if (ci == nChunks - 1) {  // all TMEM reads of this warp are done: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl.tmem_empty[as]);
        }
        {
          int64_t cnt = min((int64_t)(bn - c0), epi.cols - (n0 + c0));  // warp-uniform
          if (cnt > 32) cnt = 32;
          if (cnt > 0) epi.store_chunk(rc, n0 + c0, r, (int)cnt, mvalid, ctl.epi_sb[quad]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, accCols * 2);
}

// ------------------------------------------------------------------ weight preparation (TMA path)
// hi/lo (3xTF32) or rna-rounded (TF32) zero-padded copy of W[F][C*kk] -> [F][Kpad] in the loader's k order:
// k = (cg*kk + tap)*16 + ci  <-  W[f][(cg*16+ci)*kk + tap]
static __global__ void __launch_bounds__(256) k_prepare_weights(const float *__restrict__ w, int64_t F, int C, int kk,
                                                         int64_t Kpad, float *__restrict__ hi, float *__restrict__ lo,
                                                         int split) {
  const int64_t total = F * Kpad;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int64_t f = i / Kpad;
    const int k = (int)(i - f * Kpad);
    const int ci = k & 15, gt = k >> 4;
    const int cg = gt / kk, tap = gt - cg * kk;
    const int c = cg * 16 + ci;
    const float v = (c < C) ? __ldg(w + (f * C + c) * kk + tap) : 0.f;
    if (split) {
      const float h = tf32_hi(v);
      hi[i] = h;
      lo[i] = v - h;
    } else {
      hi[i] = tf32_rna(v);
    }
  }
}

// dgrad as a forward convolution of delta: rows = input channel c, contraction over (filter group, flipped tap, f16)
// out[c][(fg*kk + tapv)*16 + fi] = W[fg*16+fi][c][kk-1-tapv]      (W is [F][C*kk])
static __global__ void __launch_bounds__(256) k_prepare_weights_dgrad(const float *__restrict__ w, int F, int C, int kk,
                                                               int64_t Kpad, float *__restrict__ hi,
                                                               float *__restrict__ lo, int split) {
  const int64_t total = (int64_t)C * Kpad;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int64_t c = i / Kpad;
    const int k = (int)(i - c * Kpad);
    const int fi = k & 15, gt = k >> 4;
    const int fg = gt / kk, tapv = gt - fg * kk;
    const int f = fg * 16 + fi;
    const float v = (f < F) ? __ldg(w + ((int64_t)f * C + c) * kk + (kk - 1 - tapv)) : 0.f;
    if (split) {
      const float h = tf32_hi(v);
      hi[i] = h;
      lo[i] = v - h;
    } else {
      hi[i] = tf32_rna(v);
    }
  }
}

// TmaPatchLoader k order, forward and dgrad in one kernel:  k = (tap*CB + cb)*32 + ci,  channel = cb*32 + ci
//   forward: out[f][k] = W[f][channel][tap]                         (rows = F, channels = C)
//   dgrad:   out[c][k] = W[channel][c][kk-1-tap]                    (rows = C, channels = F; flipped taps)
static __global__ void __launch_bounds__(256) k_prepare_weights_patch(const float *__restrict__ w, int rows, int chans, int kk,
                                                               int CB, int dgrad, float *__restrict__ hi,
                                                               float *__restrict__ lo, int split) {
  const int Kpad = kk * CB * 32;
  const int64_t total = (int64_t)rows * Kpad;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int r = (int)(i / Kpad);
    const int k = (int)(i - (int64_t)r * Kpad);
    const int ci = k & 31, g = k >> 5;
    const int tap = g / CB, cb = g - tap * CB;
    const int ch = cb * 32 + ci;
    float v = 0.f;
    if (ch < chans)
      v = dgrad ? __ldg(w + ((int64_t)ch * rows + r) * kk + (kk - 1 - tap)) : __ldg(w + ((int64_t)r * chans + ch) * kk + tap);
    if (split) {
      const float h = tf32_hi(v);
      hi[i] = h;
      lo[i] = v - h;
    } else {
      hi[i] = tf32_rna(v);
    }
  }
}

// wgrad B operand: rows n = (c, tap) of im2col(in)^T, k = (image b, output pixel p); K-major tile (32 pixels of one
// image per k-block; requires OH*OW % 32 == 0).  lane = pixel, warp w stages rows w, w+8, ...
struct Im2colRowsLoader {
  static constexpr bool kTmaA = false;
  static constexpr bool kMnMajor = false;
  static constexpr bool kTma = false;
  static constexpr int kPrefetch = 1;
  const float *in;
  simt::ConvGeom g;
  int64_t rows;  // C*kh*kw
  int HW, kk, P;
  struct TileCtx { int R; int n0; };
  struct Regs { float v[32]; };
  // lut[3t..3t+2] = {offset inside a channel plane, dh, dw} of tap t  (kk <= 64)
  __device__ __forceinline__ void fill_lut(int *lut, int tid_all) const {
    if (tid_all < kk) {
      const int kr = tid_all / g.kw, kc = tid_all - kr * g.kw;
      lut[3 * tid_all + 0] = (kr * g.dy - g.ph) * g.W + (kc * g.dx - g.pw);
      lut[3 * tid_all + 1] = kr * g.dy - g.ph;
      lut[3 * tid_all + 2] = kc * g.dx - g.pw;
    }
  }
  __device__ __forceinline__ TileCtx tile_ctx(int R, int64_t mn0, int, int) const { return TileCtx{R, (int)mn0}; }
  __device__ __forceinline__ void load(TileCtx &cx, Regs &r, int64_t kb, int tid, const int *lut) const {
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t k0 = kb * BK;
    const int b = (int)(k0 / P);
    const int p = (int)(k0 - (int64_t)b * P) + lane;
    const int oh = p / g.OW, ow = p - oh * g.OW;
    const int ih0 = oh * g.sy, iw0 = ow * g.sx;
    const float *base = in + (int64_t)b * g.inImg + (int64_t)ih0 * g.W + iw0;
    int n = cx.n0 + warp;
    int c = n / kk, tap = n - c * kk;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float v = 0.f;
      if (warp + 8 * i < cx.R && n < rows) {
        const int dh = lut[3 * tap + 1], dw = lut[3 * tap + 2];
        if ((unsigned)(ih0 + dh) < (unsigned)g.H && (unsigned)(iw0 + dw) < (unsigned)g.W)
          v = __ldg(base + (int64_t)c * HW + lut[3 * tap]);
      }
      r.v[i] = v;
      n += 8;
      tap += 8;
      while (tap >= kk) { tap -= kk; ++c; }
    }
  }
  template <bool SPLIT>
  __device__ __forceinline__ void store(const Regs &r, uint8_t *hi, uint8_t *lo, int R, int tid) const {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int row = warp + 8 * i;
      if (row < R)
        sts_elem<SPLIT>(hi, lo, (uint32_t)((row << 7) + ((((lane >> 2) ^ (row & 7))) << 4) + ((lane & 3) << 2)), r.v[i]);
    }
  }
};

// ------------------------------------------------------------------ host side
struct Plan {
  int BN, stages;
  uint32_t accCols;
  size_t smemBytes;
  TileSched sched;
  int grid;
};

inline uint32_t pow2_cols(int bn) {
  uint32_t c = 32;
  while ((int)c < bn) c <<= 1;
  return c;
}

inline Plan make_plan(tnn_context *ctx, int64_t M, int64_t N, int64_t K, bool split, bool bMnMajor, bool allowSplitK) {
  Plan p;
  int bn;
  if (N >= 256) bn = 256;
  else bn = (int)(((N + 15) / 16) * 16);
  if (bMnMajor) {  // MN-major B tiles need whole 32-element atoms
    bn = (bn <= 32) ? 32 : (bn <= 64) ? 64 : (bn <= 128) ? 128 : 256;
  }
  if (bn < 16) bn = 16;
  p.BN = bn;
  p.accCols = pow2_cols(bn);
  const size_t stageBytes = (size_t)(BM + bn) * BK * 4 * (split ? 2 : 1);
  const size_t budget = 200 * 1024;
  int st = (int)(budget / stageBytes);
  if (st > kMaxStages) st = kMaxStages;
  if (st < 2) st = 2;
  p.stages = st;
  p.smemBytes = stageBytes * st + 1024;
  TileSched &s = p.sched;
  s.kBlocks = (int)cdiv64(K, BK);
  if (s.kBlocks < 1) s.kBlocks = 1;
  s.mTiles = (int)cdiv64(M, BM);
  s.nTiles = (int)cdiv64(N, bn);
  const int64_t sms = ctx->prop.multiProcessorCount;
  int64_t ks = 1;
  if (allowSplitK) {
    const int64_t tiles = (int64_t)s.mTiles * s.nTiles;
    if (tiles < sms && s.kBlocks >= 16) {
      ks = sms / tiles;
      const int64_t maxs = s.kBlocks / 8;
      if (ks > maxs) ks = maxs;
      if (ks < 1) ks = 1;
    }
  }
  s.kBlocksPerSplit = (int)cdiv64(s.kBlocks, ks);
  s.kSplits = (int)cdiv64(s.kBlocks, s.kBlocksPerSplit);
  s.numTiles = s.mTiles * s.nTiles * s.kSplits;
  s.wideTiles = s.numTiles;
  p.grid = (int)(s.numTiles < sms ? s.numTiles : sms);
  return p;
}

// Split the tiles of a last wave that fills at most half of the SMs in two along n (callers with TMA-fed B only).
inline void split_tail_wave(tnn_context *ctx, Plan &p) {
  TileSched &s = p.sched;
  const int sms = ctx->prop.multiProcessorCount;
  if (!TNN_TAIL_TILES || !ctx->tail_split || s.kSplits != 1 || (p.BN % 32) != 0) return;
  const int rem = s.numTiles % sms;
  if (rem == 0 || 2 * rem > sms) return;
  s.wideTiles = s.numTiles - rem;
  s.numTiles = s.wideTiles + 2 * rem;
  p.grid = s.numTiles < sms ? s.numTiles : sms;
}

template <bool SPLIT, typename LA, typename LB, typename EPI>
int launch(tnn_context *ctx, const Plan &p, LA la, LB lb, EPI epi, const CUtensorMap *tmHi = nullptr,
           const CUtensorMap *tmLo = nullptr, const CUtensorMap *tmA = nullptr) {
  auto kern = k_umma<SPLIT, LA, LB, EPI>;
  TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
  alignas(64) CUtensorMap dummy;
  memset(&dummy, 0, sizeof(dummy));
  kern<<<p.grid, kThreads, p.smemBytes, ctx->stream>>>(la, lb, epi, p.sched, p.BN, p.stages, p.accCols,
                                                       tmHi ? *tmHi : dummy, tmLo ? *tmLo : dummy, tmA ? *tmA : dummy);
  TNN_LAUNCHED(ctx);
  return 0;
}

inline bool gemm_supported(tnn_context *, int64_t M, int64_t N, int64_t K) {
  // tiny problems are latency bound: leave them to the CUDA-core kernel.  So are thin ones: with fewer than 16 rows or
  // columns the 128 x BN tile is more than 7/8 padding, and a contraction shorter than one k-block cannot use the TMA-fed
  // kernel — the register-staged tensor kernel needed 54 / 60 us for the 10-class connected layer's two backward products
  // (10 x 1024 x 128 and 128 x 1024 x 10) in the deepCIFAR10 step, the CUDA-core kernel a few us
  return M * N * K >= (1ll << 18) && K >= BK && M >= 16 && N >= 16;
}

template <typename LA, typename LB>
int gemm_launch(tnn_context *ctx, LA la, LB lb, int64_t M, int64_t N, int64_t Ktot, float alpha, float beta, float *C,
                int64_t ldc, bool allowSplitK) {
  const bool split = tnn_math_is_split(ctx);
  // split-K needs C += partial sums: natural for beta = 1; for beta = 0 a contiguous C is zeroed first (few-tile GEMMs
  // like the connected layers' batch x outputs products would otherwise run on a handful of SMs)
  const bool zeroFirst = beta == 0.0f && ldc == N;
  Plan p = make_plan(ctx, M, N, Ktot, split, LB::kMnMajor, allowSplitK && (beta == 1.0f || zeroFirst));
  if (p.sched.kSplits > 1) {
    if (beta == 0.0f) TNN_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), ctx->stream));
    GemmAtomicEpilogue epi{C, ldc, alpha, beta, M, N};
    return split ? launch<true>(ctx, p, la, lb, epi) : launch<false>(ctx, p, la, lb, epi);
  }
  GemmEpilogue epi{C, ldc, alpha, beta, M, N};
  return split ? launch<true>(ctx, p, la, lb, epi) : launch<false>(ctx, p, la, lb, epi);
}

// C = beta*C + alpha*op(A)op(B); batch > 1 either loops (strideC != 0) or is folded into k (strideC == 0, beta == 1)
inline int gemm(tnn_context *ctx, bool tA, bool tB, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                int64_t lda, int64_t sA, const float *B, int64_t ldb, int64_t sB, float beta, float *C, int64_t ldc,
                int64_t sC, int batch, bool foldBatchIntoK) {
  if (foldBatchIntoK && batch > 1 && (K % BK) != 0) return kNotTaken;  // k-blocks would straddle images
  const int nz = foldBatchIntoK ? 1 : batch;
  for (int z = 0; z < nz; ++z) {
    const float *Az = A + (foldBatchIntoK ? 0 : z * sA);
    const float *Bz = B + (foldBatchIntoK ? 0 : z * sB);
    float *Cz = C + (foldBatchIntoK ? 0 : z * sC);
    const int64_t Ktot = foldBatchIntoK ? K * batch : K;
    const int64_t K1 = K;
    int rc;
    auto vec_ok = [&](const float *p, int64_t ld, int64_t zs) {
      return (((uintptr_t)p) & 15) == 0 && (ld % 4) == 0 && (zs % 4) == 0 && (K1 % 4) == 0;
    };
    if (!tA && tB) {  // A k-contig, B k-contig
      KContigLoader la{Az, lda, M, Ktot, K1, foldBatchIntoK ? sA : 0, vec_ok(Az, lda, sA)};
      KContigLoader lb{Bz, ldb, N, Ktot, K1, foldBatchIntoK ? sB : 0, vec_ok(Bz, ldb, sB)};
      rc = gemm_launch(ctx, la, lb, M, N, Ktot, alpha, beta, Cz, ldc, true);
    } else if (!tA && !tB) {  // A k-contig, B n-contig
      KContigLoader la{Az, lda, M, Ktot, K1, foldBatchIntoK ? sA : 0, vec_ok(Az, lda, sA)};
      MnContigLoader lb{Bz, ldb, N, Ktot, K1, foldBatchIntoK ? sB : 0};
      rc = gemm_launch(ctx, la, lb, M, N, Ktot, alpha, beta, Cz, ldc, true);
    } else if (tA && !tB) {  // A m-contig, B n-contig
      MnContigLoader la{Az, lda, M, Ktot, K1, foldBatchIntoK ? sA : 0};
      MnContigLoader lb{Bz, ldb, N, Ktot, K1, foldBatchIntoK ? sB : 0};
      rc = gemm_launch(ctx, la, lb, M, N, Ktot, alpha, beta, Cz, ldc, true);
    } else {  // A m-contig, B k-contig
      MnContigLoader la{Az, lda, M, Ktot, K1, foldBatchIntoK ? sA : 0};
      KContigLoader lb{Bz, ldb, N, Ktot, K1, foldBatchIntoK ? sB : 0, vec_ok(Bz, ldb, sB)};
      rc = gemm_launch(ctx, la, lb, M, N, Ktot, alpha, beta, Cz, ldc, true);
    }
    if (rc) return rc;
  }
  return 0;
}

// cuTensorMapEncodeTiled, fetched once through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// K-major fp32 matrix [rows][Kpad] -> SWIZZLE_128B tiles of 32 k x boxRows rows (OOB rows read as zero)
inline int encode_weights_map(CUtensorMap *tm, const float *base, int64_t rows, int64_t Kpad, int boxRows) {
  EncodeTiledFn enc = get_encode_tiled();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

// same matrix as tiles of 16 k x boxRows rows (64-byte rows: SWIZZLE_64B)
inline int encode_weights_map16(CUtensorMap *tm, const float *base, int64_t rows, int64_t Kpad, int boxRows) {
  EncodeTiledFn enc = get_encode_tiled();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * 4};
  cuuint32_t box[2] = {16, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeTiled (16-k weights) failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

// NCHW activations {W, H, C, N}; box = one atom of one k-block: {TW, TH, 32, 1} = 4 KB
inline int encode_patch_map(CUtensorMap *tm, const float *base, int64_t W, int64_t H, int64_t C, int64_t N, int TW,
                            int TH, int kb = BK) {
  EncodeTiledFn enc = get_encode_tiled();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)kb, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeTiled (activations) failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

// Stride-1 convolution (forward, or dgrad posed as one) with TMA-fed activations; kNotTaken when the geometry
// does not fit (caller falls back to the gathering producer).  `v` is the forward problem actually computed:
// in = [batch][v.C][v.H][v.W], out = [batch][v.F][v.OH][v.OW].
inline int conv_patch(tnn_context *ctx, const simt::ConvGeom &v, int batch, const float *in, const float *weights,
                      bool dgradWeights, const float *scale, const float *bias, int act, float *out, bool accumulate,
                      const float *residual = nullptr) {
  if (v.sy != 1 || v.sx != 1) return kNotTaken;
  const int kk = v.kh * v.kw;
  // a 1x1 / no-padding convolution has no 2-D structure: view every image as one row of H*W pixels
  const bool flat = kk == 1 && v.ph == 0 && v.pw == 0;
  const int64_t We = flat ? (int64_t)v.H * v.W : v.W, He = flat ? 1 : v.H;
  const int64_t OWe = flat ? We : v.OW, OHe = flat ? 1 : v.OH;
  if ((We % 4) != 0 || v.C < 16 || (((uintptr_t)in) & 15) != 0) return kNotTaken;
  // Measured on B200 (tools/probe/tma_probe.cu): a tiled TMA load whose innermost start coordinate is not a
  // multiple of 16 bytes raises "illegal instruction", in or out of bounds.  Patch columns start at multiples of
  // TW >= 8, so every kernel column must shift by a multiple of 4 pixels: true for 1-wide kernels (all 1x1
  // convolutions), not for the +-1 column taps of a 3x3 kernel — those stay on the gathering producer.
  if (!flat)
    for (int kc = 0; kc < v.kw; ++kc)
      if (((kc * v.dx - v.pw) % 4) != 0) return kNotTaken;
  if (We >= (1ll << 31) || (int64_t)batch * v.C * v.H * v.W >= (1ll << 40)) return kNotTaken;
  // patch shape: the TW in {32,16,8} that wastes the fewest rows
  int TW = 0;
  double best = 0;
  for (int tw = 32; tw >= 8; tw >>= 1) {
    const int th = 32 / tw;
    const double u = (double)OWe / (double)(cdiv64(OWe, tw) * tw) * (double)OHe / (double)(cdiv64(OHe, th) * th);
    if (u > best + 1e-9) { best = u; TW = tw; }
  }
  if (best < 0.75) return kNotTaken;
  const int TH = 32 / TW;
  const int WT = (int)cdiv64(OWe, TW), HT = (int)cdiv64(OHe, TH);
  const int64_t numAtoms = (int64_t)batch * WT * HT;
  if (numAtoms >= (1ll << 29)) return kNotTaken;
  const int CB = (v.C + 31) / 32;
  const int64_t M = cdiv64(numAtoms, 4) * BM;  // rows incl. patch padding
  const int64_t N = v.F;
  const int64_t K = (int64_t)kk * CB * 32;
  const bool split = tnn_math_is_split(ctx);
  Plan p = make_plan(ctx, M, N, K, split, false, false);
  split_tail_wave(ctx, p);
  float *whi, *wlo;
  bool fresh;
  int rc = tnn_weight_buffer(ctx, weights, dgradWeights ? 2 : 1, (size_t)N * K, &whi, &wlo, &fresh);
  if (rc) return rc;
  if (fresh) {
    int64_t blocks = cdiv64(N * K, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_prepare_weights_patch<<<(int)blocks, 256, 0, ctx->stream>>>(weights, (int)N, v.C, kk, CB, dgradWeights ? 1 : 0, whi,
                                                                 wlo, split ? 1 : 0);
    TNN_LAUNCHED(ctx);
  }
  alignas(64) CUtensorMap tmHi, tmLo, tmA;
  if (split && ctx->patch_kb16) {
    // 16-channel stages: the same weight copy read as 16-k boxes, twice the stages in flight (what lifted the TMA-fed
    // SGEMM from 166 to 221 TFLOP/s).  No gain here (97.3 vs 96.6 TFLOP/s over the 28 YOLOv3 1x1 layers): the small-C
    // layers sit on the HBM roofline, the wide ones are not stage-depth bound -> opt-in (TNN_PATCH_KB16=1)
    const size_t stageBytes = (size_t)(BM + p.BN) * 16 * 4 * 2;
    int st = (int)((200 * 1024) / stageBytes);
    p.stages = st > kMaxStages ? kMaxStages : st;
    p.smemBytes = stageBytes * p.stages + 1024;
    p.sched.kBlocks = p.sched.kBlocksPerSplit = (int)(K / 16);
    rc = encode_weights_map16(&tmHi, whi, N, K, p.BN);
    if (rc) return rc;
    rc = encode_weights_map16(&tmLo, wlo, N, K, p.BN);
    if (rc) return rc;
    rc = encode_patch_map(&tmA, in, We, He, v.C, batch, TW, TH, 16);
    if (rc) return rc;
    TmaPatchLoader16 la16{{TW, TH, WT, HT, WT * HT, (int)numAtoms, batch, flat ? 0 : v.ph, flat ? 0 : v.pw, v.kw, v.dy, v.dx,
                           2 * CB}};
    if (accumulate || residual) {
      ConvEpilogueT<true> epi{out, scale, bias, v.OH * v.OW, v.outImg, act, M, N, accumulate,
                              TW, TH, WT, WT * HT, (int)OHe, (int)OWe, batch, 5, residual};
      return launch<true>(ctx, p, la16, TmaWeightsLoader{}, epi, &tmHi, &tmLo, &tmA);
    }
    ConvEpilogueT<false> epi{out, scale, bias, v.OH * v.OW, v.outImg, act, M, N, false,
                             TW, TH, WT, WT * HT, (int)OHe, (int)OWe, batch, 5, nullptr};
    return launch<true>(ctx, p, la16, TmaWeightsLoader{}, epi, &tmHi, &tmLo, &tmA);
  }
  rc = encode_weights_map(&tmHi, whi, N, K, p.BN);
  if (rc) return rc;
  rc = encode_weights_map(&tmLo, split ? wlo : whi, N, K, p.BN);
  if (rc) return rc;
  rc = encode_patch_map(&tmA, in, We, He, v.C, batch, TW, TH);
  if (rc) return rc;
  TmaPatchLoader la{TW, TH, WT, HT, WT * HT, (int)numAtoms, batch, flat ? 0 : v.ph, flat ? 0 : v.pw, v.kw, v.dy, v.dx, CB};
  TmaWeightsLoader lb{};
  if (accumulate || residual) {
    ConvEpilogueT<true> epi{out, scale, bias, v.OH * v.OW, v.outImg, act, M, N, accumulate,
                            TW, TH, WT, WT * HT, (int)OHe, (int)OWe, batch, 5, residual};
    return split ? launch<true>(ctx, p, la, lb, epi, &tmHi, &tmLo, &tmA) : launch<false>(ctx, p, la, lb, epi, &tmHi, &tmLo, &tmA);
  }
  ConvEpilogueT<false> epi{out, scale, bias, v.OH * v.OW, v.outImg, act, M, N, false,
                           TW, TH, WT, WT * HT, (int)OHe, (int)OWe, batch, 5, nullptr};
  return split ? launch<true>(ctx, p, la, lb, epi, &tmHi, &tmLo, &tmA) : launch<false>(ctx, p, la, lb, epi, &tmHi, &tmLo, &tmA);
}

// umma_halo.cuh
int conv_halo(tnn_context *ctx, const simt::ConvGeom &v, int batch, const float *in, const float *weights,
              bool dgradWeights, const float *scale, const float *bias, int act, float *out, bool accumulate,
              const float *residual = nullptr);

// conv forward: D[pixel][filter]; A = im2col gather (MN-major, producer warps), B = weights via TMA (K-major)
inline int conv_forward(tnn_context *ctx, const simt::ConvGeom &g, int batch, const float *in, const float *weights,
                        const float *scale, const float *bias, int act, float *out, const float *residual = nullptr) {
  if (ctx->tma_activations) {
    ctx->last_conv_kernel = 3;
    int rc = conv_patch(ctx, g, batch, in, weights, false, scale, bias, act, out, false, residual);
    if (rc != kNotTaken) return rc;
    ctx->last_conv_kernel = 4;
    rc = conv_halo(ctx, g, batch, in, weights, false, scale, bias, act, out, false, residual);
    if (rc != kNotTaken) return rc;
  }
  ctx->last_conv_kernel = 2;
  const int64_t M = (int64_t)batch * g.OH * g.OW;
  const int64_t N = g.F;
  const int kk = g.kh * g.kw;
  const int64_t K = (int64_t)((g.C + 15) / 16) * kk * 16;  // contraction length in the (cg, tap, ci) order
  if (kk > 64 || g.inImg >= (1ll << 31)) return kNotTaken;  // tap mask is 64 bits, offsets 32 bits
  const bool split = tnn_math_is_split(ctx);
  Plan p = make_plan(ctx, M, N, K, split, false, false);
  // no split_tail_wave here: a half-width tile costs this kernel a whole A gather (measured: 17.50 vs 17.35 ms/step)
  // 16-k stages keep twice as many in flight in the same shared memory.  That is what lifts the TMA-fed SGEMM by 30 %,
  // but here the producers pay every stage's wait/fence/arrive twice: 173 vs 195 TFLOP/s on the YOLOv3 layers -> opt-in
  const bool kb16 = split && residual == nullptr && ctx->gather_kb16;
  if (kb16) {
    const size_t stageBytes = (size_t)(BM + p.BN) * 16 * 4 * 2;
    int st = (int)((200 * 1024) / stageBytes);
    p.stages = st > kMaxStages ? kMaxStages : st;
    p.smemBytes = stageBytes * p.stages + 1024;
    p.sched.kBlocks = p.sched.kBlocksPerSplit = (int)(K / 16);
  }
  const int64_t Kpad = kb16 ? K : (int64_t)p.sched.kBlocks * BK;
  float *whi, *wlo;
  bool fresh;
  int rc = tnn_weight_buffer(ctx, weights, kb16 ? 9 : 3, (size_t)N * Kpad, &whi, &wlo, &fresh);
  if (rc) return rc;
  if (fresh) {
    int64_t blocks = cdiv64(N * Kpad, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_prepare_weights<<<(int)blocks, 256, 0, ctx->stream>>>(weights, N, g.C, kk, Kpad, whi, wlo, split ? 1 : 0);
    TNN_LAUNCHED(ctx);
  }
  alignas(64) CUtensorMap tmHi, tmLo;
  rc = kb16 ? encode_weights_map16(&tmHi, whi, N, Kpad, p.BN) : encode_weights_map(&tmHi, whi, N, Kpad, p.BN);
  if (rc) return rc;
  rc = kb16 ? encode_weights_map16(&tmLo, wlo, N, Kpad, p.BN) : encode_weights_map(&tmLo, split ? wlo : whi, N, Kpad, p.BN);
  if (rc) return rc;
  if (kb16) {
    Im2colPixelLoader16 la16{{in, g, M, g.H * g.W, kk}};
    ConvEpilogueT<false> epi{out, scale, bias, g.OH * g.OW, g.outImg, act, M, N, false, 0, 0, 0, 0, 0, 0, 0, 0, nullptr};
    return launch<true>(ctx, p, la16, TmaWeightsLoader{}, epi, &tmHi, &tmLo);
  }
  Im2colPixelLoader la{in, g, M, g.H * g.W, kk};
  Im2colPixelLoaderFast laf{{in, g, M, g.H * g.W, kk}};
  TmaWeightsLoader lb{};
  if (residual) {
    ConvEpilogueT<true> epi{out, scale, bias, g.OH * g.OW, g.outImg, act, M, N, false, 0, 0, 0, 0, 0, 0, 0, 0, residual};
    return split ? launch<true>(ctx, p, la, lb, epi, &tmHi, &tmLo) : launch<false>(ctx, p, laf, lb, epi, &tmHi, &tmLo);
  }
  ConvEpilogueT<false> epi{out, scale, bias, g.OH * g.OW, g.outImg, act, M, N, false, 0, 0, 0, 0, 0, 0, 0, 0, nullptr};
  return split ? launch<true>(ctx, p, la, lb, epi, &tmHi, &tmLo) : launch<false>(ctx, p, laf, lb, epi, &tmHi, &tmLo);
}

// dgrad (nConvolutionLayer.pas:1250-1275: TN gemm + col2im) for stride 1 / dilation 1:
//   in_delta[b][c][ih][iw] += sum_{f,kr,kc} W[f][c][kr][kc] * delta[b][f][ih + ph - kr][iw + pw - kc]
// = a forward convolution of delta (F channels, OHxOW) with the flipped, transposed weights and padding
// (kh-1-ph, kw-1-pw), accumulated into in_delta by the epilogue.
inline int conv_backward_data(tnn_context *ctx, const simt::ConvGeom &g, int batch, const float *weights,
                              const float *delta, float *in_delta) {
  const int kk = g.kh * g.kw;
  if (g.sy != 1 || g.sx != 1 || g.dy != 1 || g.dx != 1 || kk > 64) return kNotTaken;
  if (g.kh - 1 - g.ph < 0 || g.kw - 1 - g.pw < 0 || g.outImg >= (1ll << 31)) return kNotTaken;
  simt::ConvGeom v = g;  // the virtual forward problem
  v.C = g.F; v.H = g.OH; v.W = g.OW; v.F = g.C;
  v.ph = g.kh - 1 - g.ph; v.pw = g.kw - 1 - g.pw;
  v.OH = g.H; v.OW = g.W;
  v.inImg = g.outImg; v.outImg = g.inImg;
  if (ctx->tma_activations) {
    int rc = conv_patch(ctx, v, batch, delta, weights, true, nullptr, nullptr, TNN_ACT_LINEAR, in_delta, true);
    if (rc != kNotTaken) return rc;
    rc = conv_halo(ctx, v, batch, delta, weights, true, nullptr, nullptr, TNN_ACT_LINEAR, in_delta, true);
    if (rc != kNotTaken) return rc;
  }
  const int64_t M = (int64_t)batch * v.OH * v.OW;
  const int64_t N = v.F;
  const int64_t K = (int64_t)((v.C + 15) / 16) * kk * 16;
  const bool split = tnn_math_is_split(ctx);
  Plan p = make_plan(ctx, M, N, K, split, false, false);
  // no split_tail_wave here: a half-width tile costs this kernel a whole A gather (measured: 17.50 vs 17.35 ms/step)
  const int64_t Kpad = (int64_t)p.sched.kBlocks * BK;
  float *whi, *wlo;
  bool fresh;
  int rc = tnn_weight_buffer(ctx, weights, 4, (size_t)N * Kpad, &whi, &wlo, &fresh);
  if (rc) return rc;
  if (fresh) {
    int64_t blocks = cdiv64(N * Kpad, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_prepare_weights_dgrad<<<(int)blocks, 256, 0, ctx->stream>>>(weights, g.F, g.C, kk, Kpad, whi, wlo, split ? 1 : 0);
    TNN_LAUNCHED(ctx);
  }
  alignas(64) CUtensorMap tmHi, tmLo;
  rc = encode_weights_map(&tmHi, whi, N, Kpad, p.BN);
  if (rc) return rc;
  rc = encode_weights_map(&tmLo, split ? wlo : whi, N, Kpad, p.BN);
  if (rc) return rc;
  Im2colPixelLoader la{delta, v, M, v.H * v.W, kk};
  Im2colPixelLoaderFast laf{{delta, v, M, v.H * v.W, kk}};
  TmaWeightsLoader lb{};
  ConvEpilogueT<true> epi{in_delta, nullptr, nullptr, v.OH * v.OW, v.outImg, TNN_ACT_LINEAR, M, N, true, 0, 0, 0, 0, 0, 0, 0, 0, nullptr};
  return split ? launch<true>(ctx, p, la, lb, epi, &tmHi, &tmLo) : launch<false>(ctx, p, laf, lb, epi, &tmHi, &tmLo);
}

// wgrad (nConvolutionLayer.pas:1203-1232: im2col + NT gemm, beta = 1, batch accumulated):
//   weight_updates[f][(c,kr,kc)] += sum_{b,p} delta[b][f][p] * im2col(in_b)[(c,kr,kc)][p]
// M = F (rows of delta, K-major), N = C*kh*kw (gathered, K-major), K = batch*OH*OW split across CTAs (atomic adds).
inline int conv_backward_filter(tnn_context *ctx, const simt::ConvGeom &g, int batch, const float *in,
                                const float *delta, float *weight_updates) {
  const int kk = g.kh * g.kw;
  const int P = g.OH * g.OW;
  if ((P % BK) != 0 || kk > 64) return kNotTaken;  // a k-block must not straddle two images
  const int64_t M = g.F, N = (int64_t)g.C * kk, Ktot = (int64_t)batch * P;
  const bool vec = ((((uintptr_t)delta) & 15) == 0) && (P % 4 == 0);
  KContigLoader la{delta, P, M, Ktot, P, g.outImg, vec};
  Im2colRowsLoader lb{in, g, N, g.H * g.W, kk, P};
  return gemm_launch(ctx, la, lb, M, N, Ktot, 1.0f, 1.0f, weight_updates, N, true);
}

}  // namespace umma
