// conv_nhwc.cu — implicit-GEMM convolution on channels-innermost (NHWC) activations, every operand moved by TMA.
//
// Replaces, for a graph that keeps its activations channels-innermost between layers, the per-image im2col +
// gemmStridedBatched + forwardBias | batch-norm + ActivateArray sequence of TConvolutionalLayer.forwardGPU
// (reference: source/nConvolutionLayer.pas:1040-1122).  No column buffer exists and no thread gathers:
//
//   D[pixel m][filter n] (TMEM, fp32) = sum over (tap, channel) A(m, tap, c) * W(n, tap, c)
//
//   A  activations [B][H][W][pitch] fp32: ONE cp.async.bulk.tensor.4d ... im2col load per k-block delivers the
//      128 consecutive output pixels of a tile (linearised over image, row, column — rows and images are crossed
//      by the TMA unit's own traversal) x KB channels of one kernel tap, padding and the batch tail zero-filled,
//      as a K-major SWIZZLE_128B (KB = 32) / SWIZZLE_64B (KB = 16) tile: exactly what tcgen05.mma reads.
//   W  weights: the pre-split [F][tap][Cpad] copy (hi and lo TF32 planes, k_prepare_weights_patch) by tiled TMA.
//   D  epilogue: tcgen05.ld -> *scale[f] + bias[f] -> activation -> swizzled staging tile in shared memory ->
//      cp.async.bulk.tensor.2d store (UTMASTG) of 32 pixels x 32 filters; rows past the last pixel and filters
//      past F are clipped by the TMA unit.  Head layers write NCHW instead (the yolo layer and the host decode
//      read planes), with the coalesced per-plane epilogue of umma_conv.cuh.
//
// Warp roles (448 threads, one persistent CTA per SM): 0-3 epilogue, 4-11 derive the 3xTF32 low half of the A
// tile in shared memory (lo = x - trunc_tf32(x); the raw tile IS the high half because kind::tf32 reads the top
// 19 bits), 12 MMA issue, 13 TMA.  In TF32 mode nobody touches the tiles: TMA -> mbarrier -> MMA.
//
// BF16x3 mode (TNN_MATH_BF16X3): the same pipeline with every fp32 operand split into two bf16 numbers,
// hi = bf16_rn(x), lo = bf16_rn(x - hi) (x = hi + lo to 2^-18 relative), and D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi
// issued as tcgen05.mma kind::f16 — bf16 products are exact in fp32 and the accumulation is the same fp32 TMEM
// accumulator, so the result stays inside the 1e-4 gate of 3xTF32 while the tensor pipe does half the work
// (3 bf16 MMAs = 1.5 TF32-MMA equivalents per k instead of 3): the 3xTF32 kernels run at the board's power cap, so
// fewer tensor FLOPs per output is what makes them faster.  The producer warps convert the landed fp32 A tile into
// the two bf16 K-major tiles (64-byte rows, SWIZZLE_64B); the weights are pre-split bf16 planes fetched by TMA.
#include "umma_conv.cuh"
#include "conv_direct.cuh"

namespace nhwc {
using namespace umma;

typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t,
                                   const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeIm2colFn get_encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeIm2colFn)p;
  }
  return fn;
}

// im2col-mode load: coordinates {c, w, h, n} of the tile's first pixel at tap (0,0), tap offsets {offw, offh}
__device__ __forceinline__ void tma_load_im2col(uint32_t dst_smem, const CUtensorMap *tmap, int c, int w, int h, int n,
                                                uint32_t offw, uint32_t offh, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
      "[%6], {%7, %8};" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c), "r"(w), "r"(h), "r"(n), "r"(smem_u32(bar)), "h"((uint16_t)offw),
      "h"((uint16_t)offh)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tmap, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// D[tmem] (+)= A * B with bf16 operands (kind::f16), fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor: D = F32, A = B = BF16, K-major both, M = 128, N = n
__host__ __device__ inline uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// two floats -> packed bf16x2 (round to nearest even), low half = first argument
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
// split 8 consecutive floats into 8 bf16 high parts and 8 bf16 low parts (each one 16-byte chunk)
__device__ __forceinline__ void split_bf16x8(const float4 &a, const float4 &b, uint4 &hi, uint4 &lo) {
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
    const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xFFFF0000u);
    l[i] = pack_bf16x2(x[2 * i] - h0, x[2 * i + 1] - h1);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

enum { MODE_TF32 = 0, MODE_3X = 1, MODE_BF16X3 = 2 };

struct Geo {
  int OW, OHW;             // output row length, output pixels per image
  int sx, sy, pw, ph, dx, dy, kw;
  int cBlocks;             // k-blocks per kernel tap (Cpad / KB)
  int M;                   // batch * OH * OW
};

// epilogue that writes channels-innermost rows through the TMA store (scale/bias/activation like ConvEpilogueT)
struct NhwcEpilogue {
  static constexpr bool kTmaStore = true;
  const float *scale, *bias;
  int act;
  int64_t rows, cols;  // M, F
  // optional: the shortcut layer that follows (TAddLayer.forwardGPU naddlayer.pas:885-910, linear activation):
  // value + residual[row][filter], rows of `resPitch` floats — one fp32 add after the activation, like the separate layer
  const float *residual;
  int64_t resPitch;
  // ... or the same operand as pre-split planes (a graph that keeps every tensor between convolutions only as planes):
  // residual = hi + lo, the value the planes hold; resSplitPitch / resSplitPlane in bf16 elements
  const uint16_t *resHi;
  int64_t resSplitPitch, resSplitPlane;
  // what to write: fp32 rows (tmOut) and / or the pre-split bf16 planes the next convolution's TMA reads directly
  // (tmOutHi / tmOutLo: hi = bf16_rn(v), lo = bf16_rn(v - hi) — the very split the converter warps would derive)
  bool f32Out, splitOut;
};
// NCHW planes (head layers): the epilogue of the NCHW kernels
struct NchwEpilogue : ConvEpilogueT<false> {
  static constexpr bool kTmaStore = false;
  // yoloEntries > 0: the yolo layer that follows a head convolution, folded in (TYoloLayer.forwardGPU
  // nyololayer.pas:1081-1099 = tnn_yolo_activate): filter f is entry e = f % yoloEntries of its anchor; logistic on
  // every entry but the raw w / h (e = 2, 3), x / y then mapped v*yoloA + yoloB — the same expressions, so the same bits
  int yoloEntries;
  float yoloA, yoloB;
  __device__ __forceinline__ void store_chunk(const Row &r, int64_t n, const uint32_t (&acc)[32], int cnt, bool rowvalid,
                                              float *scratch) const {
    if (yoloEntries <= 0) {
      ConvEpilogueT<false>::store_chunk(r, n, acc, cnt, rowvalid, scratch);
      return;
    }
    const int lane = threadIdx.x & 31;
    __syncwarp();
    scratch[lane] = (scale && lane < cnt) ? __ldg(scale + n + lane) : 1.0f;
    scratch[32 + lane] = (bias && lane < cnt) ? __ldg(bias + n + lane) : 0.0f;
    __syncwarp();
    float *q = r.q + n * P;
    int e = (int)(n % yoloEntries);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float v = __fadd_rn(__fmul_rn(__uint_as_float(acc[j]), scratch[j]), scratch[32 + j]);   // (linear head convolution)
      if (e != 2 && e != 3) {
        float t = 1.0f / (1.0f + expf(-v));
        if (e < 2 && yoloA != 1.0f) t = __fadd_rn(__fmul_rn(t, yoloA), yoloB);
        v = t;
      }
      if (rowvalid && j < cnt) q[(int64_t)j * P] = v;
      if (++e == yoloEntries) e = 0;
    }
  }
};

struct Ctl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t araw[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  alignas(16) float epi_sb[kEpilogueWarps][64];   // (a second set of epilogue warps keeps its scratch behind its staging tiles)
};

// The eight converter warps work as kConvGroups independent groups, group g taking k-blocks g, g + kConvGroups, ...:
// a group's wait -> LDS -> convert -> STS -> fence -> arrive chain is latency, not issue, bound, and with all warps on
// the same stage nothing overlaps it.
// Measured over the YOLOv3 layers (batch 64): 3xTF32 19.3 ms with two groups vs 20.5 with one; BF16x3 13.5 ms with one
// group vs 14.8 with two (its conversion is twice the instructions per byte, so a half-size group falls behind).
template <int MODE>
struct ConvGroups {
  static constexpr int kGroups = MODE == MODE_3X ? 2 : 1;
  static constexpr int kWarps = kProducerWarps / kGroups;
  static constexpr int kThreads = kWarps * 32;
};
constexpr int kStageTile = 32 * 32 * 4;  // one epilogue staging tile: 32 pixels x 32 filters
constexpr int kStageBufs = 2;            // per epilogue warp

// ---- device pieces shared by the single-CTA kernel and the CTA-pair kernel
// producers: derive the second operand plane(s) of the landed fp32 A tile `sa` (aBytes = 128 rows x KB floats)
// NT = threads that share the tile (a converter group)
template <int MODE, int KB, int NT>
__device__ __forceinline__ void convert_stage(uint8_t *sa, int tid) {
  constexpr bool BF = MODE == MODE_BF16X3;
  constexpr int aBytes = BM * KB * 4;
  if constexpr (BF) {
    // One unit = 8 consecutive k of one row: two 16-byte fp32 chunks in, one 16-byte bf16 chunk out per plane.
    // KB = 32: raw rows of 128 bytes, chunk c at c ^ (r & 7) (SWIZZLE_128B); bf16 rows of 64 bytes, chunk j at
    //          j ^ ((r >> 1) & 3) (SWIZZLE_64B).
    // KB = 16: raw rows of 64 bytes, chunk c at c ^ ((r >> 1) & 3) (SWIZZLE_64B); bf16 rows of 32 bytes, chunk j
    //          at j ^ ((r >> 2) & 1) (SWIZZLE_32B).
    uint8_t *shi = sa + aBytes, *slo = shi + aBytes / 2;
    constexpr int UPR = KB / 8;  // units per row
#pragma unroll
    for (int i = 0; i < (BM * UPR) / NT; ++i) {
      const int u = i * NT + tid;
      const int r = u / UPR, j = u % UPR;
      const int rsw = KB == 32 ? (r & 7) : ((r >> 1) & 3);
      const uint8_t *rowp = sa + r * (KB * 4);
      const float4 a = *reinterpret_cast<const float4 *>(rowp + (((2 * j) ^ rsw) << 4));
      const float4 b = *reinterpret_cast<const float4 *>(rowp + (((2 * j + 1) ^ rsw) << 4));
      uint4 hi, lo;
      split_bf16x8(a, b, hi, lo);
      const int osw = KB == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1);
      const int o = r * (KB * 2) + ((j ^ osw) << 4);
      *reinterpret_cast<uint4 *>(shi + o) = hi;
      *reinterpret_cast<uint4 *>(slo + o) = lo;
    }
  } else {
#pragma unroll
    for (int i = 0; i < aBytes / (16 * NT); ++i) {
      const int off = (i * NT + tid) * 16;
      const float4 v = *reinterpret_cast<const float4 *>(sa + off);
      *reinterpret_cast<float4 *>(sa + aBytes + off) =
          make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
    }
  }
}

// epilogue of one tile for one warp (32 accumulator rows): TMEM -> scale/bias/activation -> TMA store (channels-innermost
// rows) or per-plane stores (NCHW).  `release()` is called by lane 0 once the warp's last TMEM read has completed.
struct StoreRows {   // the default store: 32 consecutive pixel rows of the 2-D {F, M} map
  int row0;
  __device__ __forceinline__ void operator()(const CUtensorMap *tm, uint32_t src, int n) const { tma_store_2d(tm, src, n, row0); }
};
template <typename EPI, typename RELEASE, typename STORE>
__device__ __forceinline__ void drain_tile(const EPI &epi, uint32_t tmem_d, int bn, int64_t n0, int row0, float *scratch,
                                           uint8_t *stg, int &sbuf, const CUtensorMap *tmOut, const CUtensorMap *tmOutHi,
                                           const CUtensorMap *tmOutLo, RELEASE release, STORE store, int ci0 = 0,
                                           int ciStep = 1) {
  // ci0 / ciStep: the warp drains chunks ci0, ci0 + ciStep, ... of the tile (two sets of epilogue warps share a tile
  // chunk by chunk); `release` follows the warp's LAST TMEM read — at once if it has no chunk in this tile
  const int lane = threadIdx.x & 31;
  const int nChunks = (bn + 31) >> 5;
  if (ci0 >= nChunks) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) release();
    return;
  }
  if constexpr (EPI::kTmaStore) {
    const bool leaky = epi.act == TNN_ACT_LEAKY || epi.act == TNN_ACT_REVLEAKY;
    const bool simple = leaky || epi.act == TNN_ACT_LINEAR;
    const float neg = leaky ? 0.1f : 1.0f;
    for (int ci = ci0; ci < nChunks; ci += ciStep) {
      const int c0 = ci * 32;
      uint32_t r[32];
      tmem_ld32(tmem_d + (uint32_t)c0, r);
      if (ci + ciStep >= nChunks) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release();
      }
      int64_t cnt = min((int64_t)(bn - c0), epi.cols - (n0 + c0));  // warp-uniform
      if (cnt > 32) cnt = 32;
      if (cnt <= 0 || row0 >= epi.rows) continue;
      const int64_t n = n0 + c0;
      // the shortcut's other operand: this thread's 32 consecutive filters of its pixel row, fetched before the staging
      // wait so that the loads are in flight while the scale / bias scratch is filled
      // (fp32 rows: res[j] = filters 4j .. 4j+3; planes: res[q] = the hi halves of filters 8q .. 8q+7, res[4 + q] the lo)
      uint4 res[8];
      const bool resSplit = epi.resHi != nullptr;
      const bool addRes = epi.residual != nullptr || resSplit;
      if (addRes) {
        const int64_t row = (int64_t)row0 + lane;
        const bool rowOk = row < epi.rows;
        if (resSplit) {
          const uint4 *rh = reinterpret_cast<const uint4 *>(epi.resHi + (rowOk ? row : 0) * epi.resSplitPitch + n);
          const uint4 *rl = reinterpret_cast<const uint4 *>(epi.resHi + epi.resSplitPlane + (rowOk ? row : 0) * epi.resSplitPitch + n);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool ok = rowOk && 8 * q < cnt;
            res[q] = ok ? __ldg(rh + q) : make_uint4(0u, 0u, 0u, 0u);
            res[4 + q] = ok ? __ldg(rl + q) : make_uint4(0u, 0u, 0u, 0u);
          }
        } else {
          const uint4 *rq = reinterpret_cast<const uint4 *>(epi.residual + (rowOk ? row : 0) * epi.resPitch + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) res[j] = (rowOk && 4 * j < cnt) ? __ldg(rq + j) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
      __syncwarp();
      scratch[lane] = (epi.scale && lane < cnt) ? __ldg(epi.scale + n + lane) : 1.0f;
      scratch[32 + lane] = (epi.bias && lane < cnt) ? __ldg(epi.bias + n + lane) : 0.0f;
      // staging: fp32 rows take one 4 KB tile (SWIZZLE_128B), the two bf16 planes half a tile each (64-byte rows,
      // SWIZZLE_64B).  One of the two -> the warp's two tiles alternate; both -> they are used together.
      const bool both = epi.f32Out && epi.splitOut;
      if (lane == 0) {
        if (both) bulk_wait_read<0>();
        else bulk_wait_read<kStageBufs - 1>();  // the staging tile written two chunks ago has been read
      }
      __syncwarp();
      uint8_t *tile = both ? stg : stg + sbuf * kStageTile;
      uint8_t *tileHi = both ? stg + kStageTile : tile;
      uint8_t *tileLo = tileHi + kStageTile / 2;
      uint8_t *rowp = tile + lane * 128;
#pragma unroll
      for (int j8 = 0; j8 < 32; j8 += 8) {
        float o[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 s4 = *reinterpret_cast<const float4 *>(scratch + j8 + 4 * h);
          const float4 b4 = *reinterpret_cast<const float4 *>(scratch + 32 + j8 + 4 * h);
          const float s[4] = {s4.x, s4.y, s4.z, s4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float v = __fadd_rn(__fmul_rn(__uint_as_float(r[j8 + 4 * h + u]), s[u]), b[u]);
            o[4 * h + u] = simple ? fmaxf(v, __fmul_rn(neg, v)) : act_out_of_line(v, epi.act);
          }
          if (addRes) {
            float rr[4];
            if (resSplit) {
              // two bf16 per word, the even filter in the low half (pack_bf16x2): value = hi + lo, exact in fp32
              const uint4 qh = res[j8 >> 3], ql = res[4 + (j8 >> 3)];
              const uint32_t wh0 = h ? qh.z : qh.x, wh1 = h ? qh.w : qh.y, wl0 = h ? ql.z : ql.x, wl1 = h ? ql.w : ql.y;
              rr[0] = __fadd_rn(__uint_as_float(wh0 << 16), __uint_as_float(wl0 << 16));
              rr[1] = __fadd_rn(__uint_as_float(wh0 & 0xFFFF0000u), __uint_as_float(wl0 & 0xFFFF0000u));
              rr[2] = __fadd_rn(__uint_as_float(wh1 << 16), __uint_as_float(wl1 << 16));
              rr[3] = __fadd_rn(__uint_as_float(wh1 & 0xFFFF0000u), __uint_as_float(wl1 & 0xFFFF0000u));
            } else {
              const uint4 q = res[(j8 >> 2) + h];
              rr[0] = __uint_as_float(q.x); rr[1] = __uint_as_float(q.y); rr[2] = __uint_as_float(q.z); rr[3] = __uint_as_float(q.w);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) o[4 * h + u] = __fadd_rn(o[4 * h + u], rr[u]);
          }
        }
        const float4 oa = make_float4(o[0], o[1], o[2], o[3]), ob = make_float4(o[4], o[5], o[6], o[7]);
        if (epi.f32Out) {
          // SWIZZLE_128B staging: 16-byte chunk index XOR (row & 7) — conflict-free and what the TMA store expects
          *reinterpret_cast<float4 *>(rowp + ((((j8 >> 2) ^ (lane & 7))) << 4)) = oa;
          *reinterpret_cast<float4 *>(rowp + (((((j8 >> 2) + 1) ^ (lane & 7))) << 4)) = ob;
        }
        if (epi.splitOut) {
          uint4 hi, lo;
          split_bf16x8(oa, ob, hi, lo);
          const int o16 = lane * 64 + (((j8 >> 3) ^ ((lane >> 1) & 3)) << 4);  // SWIZZLE_64B
          *reinterpret_cast<uint4 *>(tileHi + o16) = hi;
          *reinterpret_cast<uint4 *>(tileLo + o16) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (epi.f32Out) store(tmOut, smem_u32(tile), (int)n);
        if (epi.splitOut) {
          store(tmOutHi, smem_u32(tileHi), (int)n);
          store(tmOutLo, smem_u32(tileLo), (int)n);
        }
        bulk_commit();
      }
      sbuf ^= 1;
    }
  } else {
    const int64_t m = (int64_t)row0 + lane;
    bool mvalid;
    typename EPI::Row rc = epi.row(m, mvalid);
    for (int ci = ci0; ci < nChunks; ci += ciStep) {
      const int c0 = ci * 32;
      uint32_t r[32];
      tmem_ld32(tmem_d + (uint32_t)c0, r);
      if (ci + ciStep >= nChunks) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release();
      }
      int64_t cnt = min((int64_t)(bn - c0), epi.cols - (n0 + c0));
      if (cnt > 32) cnt = 32;
      if (cnt > 0) epi.store_chunk(rc, n0 + c0, r, (int)cnt, mvalid, scratch);
    }
  }
}

// Programmatic dependent launch: a kernel launched with programmatic stream serialization may start while the previous
// kernel of the stream is still running; `pdl_wait` blocks until that kernel has completed and its writes are visible.
// Everything before it (barrier initialisation, TMEM allocation, descriptor prefetch, the cluster rendezvous) overlaps
// the previous layer's last, partly filled wave.  `pdl_trigger` lets the NEXT kernel's CTAs be scheduled as soon as SMs
// free up.  Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ASPLIT (BF16x3 only): the activations arrive PRE-SPLIT — two bf16 planes (hi, lo) written by the producing layer's
// epilogue — so both A tiles land by TMA in the tensor-core layout and nobody converts: the TF32-mode pipeline
// (TMA -> mbarrier -> MMA) with the BF16x3 arithmetic.  tmA = hi plane, tmA2 = lo plane.
template <int MODE, int KB, bool ASPLIT, typename EPI>
__global__ void __launch_bounds__(kThreads, 1)
k_conv_nhwc(EPI epi, TileSched sched, Geo g, int BN, int stages, uint32_t accCols, int epiSets,
            const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
            const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOutHi,
            const __grid_constant__ CUtensorMap tmOutLo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Ctl ctl;
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the shared address space
  // and emits LDS / STS instead of generic LD / ST for every access below)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr bool SPLIT = MODE != MODE_TF32;   // two B planes
  constexpr bool BF = MODE == MODE_BF16X3;
  constexpr bool CONVERT = SPLIT && !ASPLIT;  // producers derive the second plane(s) of the landed A tile
  static_assert(KB == 32 || KB == 16, "stages are 32 or 16 channels deep");
  static_assert(!ASPLIT || (BF && KB == 32), "pre-split activations: BF16x3, 32-channel stages");
  constexpr int aBytes = BM * KB * 4;                 // the landed fp32 A tile (= the two landed bf16 tiles of ASPLIT)
  const int bBytes = BN * KB * (BF ? 2 : 4);          // one weight plane of the stage
  // stage layout: 3xTF32  A(=hi) | A_lo | B_hi | B_lo;   TF32  A | B;   BF16x3  A raw | A_hi | A_lo | B_hi | B_lo (bf16 tiles)
  //               BF16x3 pre-split  A_hi | A_lo | B_hi | B_lo
  const int bOff = BF ? (ASPLIT ? 1 : 2) * aBytes : (SPLIT ? 2 : 1) * aBytes;
  const int stageBytes = bOff + bBytes * (SPLIT ? 2 : 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&ctl.full[s], ConvGroups<MODE>::kWarps);
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
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    if (ASPLIT) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmBhi);
    if (SPLIT) tma_prefetch_desc(&tmBlo);
    if (EPI::kTmaStore) tma_prefetch_desc(&tmOut);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  pdl_wait();   // nothing above reads or writes global memory
  const uint32_t tmem_base = ctl.tmem_base;

  // epiSets == 2 (TMA-only modes, narrow tiles): warps 4-7 are a second set of epilogue warps; here the two sets take
  // ALTERNATE TILES — set s drains accumulator s — because a narrow tile has a single 32-filter chunk to drain
  // (the CTA-pair kernel's sets alternate chunks within a tile)
  const bool secondEpi = !CONVERT && epiSets == 2 && warp >= kEpilogueWarps && warp < 2 * kEpilogueWarps;
  if (warp >= kFirstProducerWarp && warp < kMmaWarp && !secondEpi) {
    // ===================== split modes on fp32 activations: second plane(s) of the landed A tile =====================
    if (CONVERT) {
      constexpr int kConvGroups = ConvGroups<MODE>::kGroups, kConvGroupWarps = ConvGroups<MODE>::kWarps;
      const int group = (warp - kFirstProducerWarp) / kConvGroupWarps;
      const int tid = threadIdx.x - (kFirstProducerWarp + group * kConvGroupWarps) * 32;
      int64_t total = 0;
      for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x) {
        int mt, nt, k0, k1;
        sched.decode(t, mt, nt, k0, k1);
        total += k1 - k0;
      }
      for (int64_t n = group; n < total; n += kConvGroups) {
        const int stage = (int)(n % stages);
        const uint32_t phase = (uint32_t)((n / stages) & 1);
        mbar_wait_warp(&ctl.araw[stage], phase);
        uint8_t *sa = smem + (size_t)stage * stageBytes;
        convert_stage<MODE, KB, ConvGroups<MODE>::kThreads>(sa, tid);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl.full[stage]);
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
      const uint32_t idesc = BF ? make_idesc_bf16(sched.half(t) ? BN >> 1 : BN)
                                : make_idesc(sched.half(t) ? BN >> 1 : BN, false, false);
      const int as = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait_warp(&ctl.tmem_empty[as], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_warp(CONVERT ? &ctl.full[stage] : &ctl.araw[stage], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
          const uint32_t sb = sa + bOff;
          // K-major rows of KB floats: 8-row groups are 8*KB*4 bytes apart, one UMMA_K = 8 step = +32 bytes
          constexpr uint32_t kSbo = KB == 32 ? 1024 : 512, kLt = KB == 32 ? 2u : 4u;
          if constexpr (BF) {
            // bf16 tiles: rows of KB*2 bytes (SWIZZLE_64B for 32 k, SWIZZLE_32B for 16 k), 8-row groups KB*16 bytes apart;
            // UMMA_K = 16 -> +32 bytes per step
            const uint32_t shi = ASPLIT ? sa : sa + aBytes, slo = shi + aBytes / 2;
            constexpr uint32_t bSbo = KB * 16, bLt = KB == 32 ? 4u : 6u;
#pragma unroll
            for (int ks = 0; ks < KB / 16; ++ks) {
              const uint64_t ah = make_desc(shi + ks * 32, 16, bSbo, bLt), al = make_desc(slo + ks * 32, 16, bSbo, bLt);
              const uint64_t bh = make_desc(sb + ks * 32, 16, bSbo, bLt), bl = make_desc(sb + bBytes + ks * 32, 16, bSbo, bLt);
              const uint32_t acc = (kb != kb0 || ks != 0) ? 1u : 0u;
              mma_bf16(tmem_d, al, bh, idesc, acc);  // small terms first
              mma_bf16(tmem_d, ah, bl, idesc, 1u);
              mma_bf16(tmem_d, ah, bh, idesc, 1u);
            }
          } else
#pragma unroll
          for (int ks = 0; ks < KB / 8; ++ks) {
            const uint64_t ah = make_desc(sa + ks * 32, 16, kSbo, kLt);
            const uint64_t bh = make_desc(sb + ks * 32, 16, kSbo, kLt);
            const uint32_t acc = (kb != kb0 || ks != 0) ? 1u : 0u;
            if (SPLIT) {
              const uint64_t al = make_desc(sa + aBytes + ks * 32, 16, kSbo, kLt);
              const uint64_t bl = make_desc(sb + bBytes + ks * 32, 16, kSbo, kLt);
              mma_tf32(tmem_d, al, bh, idesc, acc);  // small terms first
              mma_tf32(tmem_d, ah, bl, idesc, 1u);
              mma_tf32(tmem_d, ah, bh, idesc, 1u);
            } else {
              mma_tf32(tmem_d, ah, bh, idesc, acc);
            }
          }
          mma_commit(&ctl.empty[stage]);
          if (kb == kb1 - 1) mma_commit(&ctl.tmem_full[as]);
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kTmaWarp) {
    // ===================== TMA producer of both operand tiles (one thread) =====================
    if (lane == 0) {
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x) {
        int mt, nt, kb0, kb1;
        sched.decode(t, mt, nt, kb0, kb1);
        const int n0 = nt * BN + (sched.half(t) == 2 ? BN >> 1 : 0);
        // first pixel of the tile -> (image, row, column); the TMA unit walks on from there
        const int m0 = mt * BM;
        const int img = m0 / g.OHW;
        const int rem = m0 - img * g.OHW;
        const int oh0 = rem / g.OW;
        const int ow0 = rem - oh0 * g.OW;
        const int cw = ow0 * g.sx - g.pw, ch = oh0 * g.sy - g.ph;
        int tap = kb0 / g.cBlocks;
        int cb = kb0 - tap * g.cBlocks;
        int kr = tap / g.kw;
        int kc = tap - kr * g.kw;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&ctl.empty[stage], phase ^ 1);
          const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
          const uint32_t sb = sa + bOff;
          mbar_arrive_expect_tx(&ctl.araw[stage], (uint32_t)(aBytes + bBytes * (SPLIT ? 2 : 1)));
          tma_load_im2col(sa, &tmA, cb * KB, cw, ch, img, (uint32_t)(kc * g.dx), (uint32_t)(kr * g.dy), &ctl.araw[stage]);
          if (ASPLIT)
            tma_load_im2col(sa + aBytes / 2, &tmA2, cb * KB, cw, ch, img, (uint32_t)(kc * g.dx), (uint32_t)(kr * g.dy),
                            &ctl.araw[stage]);
          tma_load_2d(sb, &tmBhi, kb * KB, n0, &ctl.araw[stage]);
          if (SPLIT) tma_load_2d(sb + bBytes, &tmBlo, kb * KB, n0, &ctl.araw[stage]);
          if (++cb == g.cBlocks) {
            cb = 0;
            if (++kc == g.kw) { kc = 0; ++kr; }
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps 0-3 (and 4-7 as a second set) =====================
    const int quad = warp & 3, set = warp >> 2;
    int it = 0;
    uint8_t *stg = smem + (size_t)stages * stageBytes + (size_t)warp * (kStageBufs * kStageTile);
    int sbuf = 0;
    float *scratch = set == 0 ? ctl.epi_sb[quad]
                              : reinterpret_cast<float *>(smem + (size_t)stages * stageBytes +
                                                          (size_t)2 * kEpilogueWarps * (kStageBufs * kStageTile)) + quad * 64;
    for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x, ++it) {
      if (epiSets == 2 && (it & 1) != set) continue;   // the other set's accumulator
      int mt, nt, kb0, kb1;
      sched.decode(t, mt, nt, kb0, kb1);
      const int as = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      const int hf = sched.half(t);
      const int bn = hf ? BN >> 1 : BN;
      const int64_t n0 = (int64_t)nt * BN + (hf == 2 ? BN >> 1 : 0);
      const int row0 = mt * BM + quad * 32;  // first pixel row of this warp
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols + ((uint32_t)(quad * 32) << 16);
      mbar_wait_warp_relaxed(&ctl.tmem_full[as], use & 1, 64);
      tc_fence_after();
      drain_tile(epi, tmem_d, bn, n0, row0, scratch, stg, sbuf, &tmOut, &tmOutHi, &tmOutLo,
                 [&] { mbar_arrive(&ctl.tmem_empty[as]); }, StoreRows{row0});
    }
    if (EPI::kTmaStore && lane == 0) bulk_wait_all();  // the staging tiles must outlive their stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, accCols * 2);
}

// =====================================================================================================================
// CTA-pair version (tcgen05.mma.cta_group::2).  The single-CTA kernel is bound by shared-memory bandwidth, not by the
// tensor pipe: per 32-deep k-block in TF32 mode the TMA unit writes 48 KB (16 A + 32 B) and the MMAs read 48 KB —
// 96 KB against the ~64 KB the SM moves in the 512 cycles four 128x256x8 MMAs take.  Two CTAs of one cluster (the two
// SMs of a TPC) compute a 256-pixel x BN-filter tile together: each CTA lands the A rows of ITS 128 pixels and only
// HALF of the weight tile (BN/2 filter rows); the pair's tensor cores read both B halves.  Per CTA and k-block the
// weight fill and the weight reads halve: 32 + 32 KB in TF32 mode — the tensor pipe becomes the bound.
//
// Protocol (rank = %cluster_ctarank, leader = rank 0):
//   araw[s]   (split modes) per CTA: its fp32 A tile has landed; the producer warps derive the second plane(s)
//   full[s]   the LEADER's barrier: every TMA load of both CTAs that the MMA reads directly signals it
//             (cp.async.bulk.tensor ... cta_group::2 — the completion goes to the barrier at the same offset in the
//             leader CTA), plus 8 local + 8 remote producer-warp arrivals in the split modes
//   empty[s]  one per CTA, released together by the leader's multicast tcgen05.commit
//   tmem_full[a] per CTA (multicast commit); tmem_empty[a] the leader's: 4 local + 4 remote epilogue-warp arrivals
//   The leader's elected thread issues every MMA (M = 256 over the pair); each CTA's epilogue drains its own 128 lanes.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (CTA-scope release)
// semantics: the data the peer's tensor core will read was made visible by fence.proxy.async before this arrive; a
// cluster-scope release here costs ~0.9 us per stage and serialises the producer warps (measured: 155 vs 330 TFLOP/s).
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t rank) {
  asm volatile(
      "{\n.reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <bool BF>
__device__ __forceinline__ void mma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  if (BF)
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// when the MMAs issued so far complete: arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma2_commit(uint64_t *bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// tile -> local smem, completion bytes on the LEADER's barrier (the peer bit of the barrier address cleared)
__device__ __forceinline__ void tma2_load_2d(uint32_t dst_smem, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], "
      "[%4];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar) & 0xFEFFFFFFu)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col(uint32_t dst_smem, const CUtensorMap *tmap, int c, int w, int h, int n,
                                                 uint32_t offw, uint32_t offh, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], "
      "[%1, {%2, %3, %4, %5}], [%6], {%7, %8};" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c), "r"(w), "r"(h), "r"(n), "r"(smem_u32(bar) & 0xFEFFFFFFu),
      "h"((uint16_t)offw), "h"((uint16_t)offh)
      : "memory");
}

// BN = filters per pair tile (multiple of 32, <= 256): each CTA holds BN/2 weight rows.  mTiles counts 128-pixel tiles.
template <int MODE, int KB, bool ASPLIT, typename EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_conv_nhwc_pair(EPI epi, int mTiles, int nTiles, int kBlocks, Geo g, int BN, int stages, int epiSets,
                 const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOutHi,
                 const __grid_constant__ CUtensorMap tmOutLo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Ctl ctl;
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the shared address space
  // and emits LDS / STS instead of generic LD / ST for every access below)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr bool SPLIT = MODE != MODE_TF32;
  constexpr bool BF = MODE == MODE_BF16X3;
  constexpr bool CONVERT = SPLIT && !ASPLIT;
  static_assert(!ASPLIT || (BF && KB == 32), "pre-split activations: BF16x3, 32-channel stages");
  constexpr int aBytes = BM * KB * 4;
  const int bBytes = (BN >> 1) * KB * (BF ? 2 : 4);   // this CTA's half of one weight plane
  const int bOff = BF ? (ASPLIT ? 1 : 2) * aBytes : (SPLIT ? 2 : 1) * aBytes;
  const int stageBytes = bOff + bBytes * (SPLIT ? 2 : 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, numPairs = gridDim.x >> 1;
  const int mPairs = (mTiles + 1) >> 1;
  const int numTiles = mPairs * nTiles;   // pair tiles: 256 pixels x BN filters, pixel-pair index fastest
  constexpr uint32_t kAccStride = 256;    // TMEM columns between the two accumulator buffers

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&ctl.full[s], CONVERT ? 2 * ConvGroups<MODE>::kWarps + 1 : 1);
      mbar_init(&ctl.empty[s], 1);
      mbar_init(&ctl.araw[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl.tmem_full[s], 1);
      mbar_init(&ctl.tmem_empty[s], 2 * kEpilogueWarps * epiSets);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc2(&ctl.tmem_base, 512);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    if (ASPLIT) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmBhi);
    if (SPLIT) tma_prefetch_desc(&tmBlo);
    if (EPI::kTmaStore) tma_prefetch_desc(&tmOut);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before anyone signals across the pair
  tc_fence_after();
  pdl_trigger();
  pdl_wait();   // nothing above reads or writes global memory
  const uint32_t tmem_base = ctl.tmem_base;

  // TMA-only modes leave warps 4-11 without work: with epiSets == 2 warps 4-7 (TMEM lane quadrants 0-3 again) form a
  // second set of epilogue warps and the two sets drain a tile's 32-filter chunks alternately.  The four epilogue warps
  // are one warp per scheduler, so every latency of a chunk's chain is exposed: on a 1x1 layer on planes ncu showed them
  // in the chunk body 84 % of the time at ~0.2 instructions per cycle — layers with short main loops are epilogue-bound.
  const bool secondEpi = !CONVERT && epiSets == 2 && warp >= kEpilogueWarps && warp < 2 * kEpilogueWarps;
  if (warp >= kFirstProducerWarp && warp < kMmaWarp && !secondEpi) {
    // ===================== split modes on fp32 activations: second plane(s) of this CTA's A tile =====================
    if (CONVERT) {
      constexpr int kConvGroups = ConvGroups<MODE>::kGroups, kConvGroupWarps = ConvGroups<MODE>::kWarps;
      const int group = (warp - kFirstProducerWarp) / kConvGroupWarps;
      const int tid = threadIdx.x - (kFirstProducerWarp + group * kConvGroupWarps) * 32;
      int64_t total = 0;
      for (int t = pair; t < numTiles; t += numPairs) total += kBlocks;
      for (int64_t n = group; n < total; n += kConvGroups) {
        const int stage = (int)(n % stages);
        const uint32_t phase = (uint32_t)((n / stages) & 1);
        mbar_wait_warp(&ctl.araw[stage], phase);
        convert_stage<MODE, KB, ConvGroups<MODE>::kThreads>(smem + (size_t)stage * stageBytes, tid);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&ctl.full[stage]);
          else mbar_arrive_cluster(&ctl.full[stage], 0);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer: the leader's elected thread drives both tensor cores =====================
    if (rank == 0) {
      // D = F32, K-major operands, N = BN, M = 256 over the pair
      const uint32_t idesc = (1u << 4) | ((BF ? 1u : 2u) << 7) | ((BF ? 1u : 2u) << 10) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = pair; t < numTiles; t += numPairs, ++it) {
        const int as = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait_warp(&ctl.tmem_empty[as], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * kAccStride;
        for (int kb = 0; kb < kBlocks; ++kb) {
          mbar_wait_warp(&ctl.full[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
            const uint32_t sb = sa + bOff;
            if constexpr (BF) {
              const uint32_t shi = ASPLIT ? sa : sa + aBytes, slo = shi + aBytes / 2;
              constexpr uint32_t bSbo = KB * 16, bLt = KB == 32 ? 4u : 6u;
#pragma unroll
              for (int ks = 0; ks < KB / 16; ++ks) {
                const uint64_t ah = make_desc(shi + ks * 32, 16, bSbo, bLt), al = make_desc(slo + ks * 32, 16, bSbo, bLt);
                const uint64_t bh = make_desc(sb + ks * 32, 16, bSbo, bLt), bl = make_desc(sb + bBytes + ks * 32, 16, bSbo, bLt);
                const uint32_t acc = (kb != 0 || ks != 0) ? 1u : 0u;
                mma2<true>(tmem_d, al, bh, idesc, acc);
                mma2<true>(tmem_d, ah, bl, idesc, 1u);
                mma2<true>(tmem_d, ah, bh, idesc, 1u);
              }
            } else {
              constexpr uint32_t kSbo = KB == 32 ? 1024 : 512, kLt = KB == 32 ? 2u : 4u;
#pragma unroll
              for (int ks = 0; ks < KB / 8; ++ks) {
                const uint64_t ah = make_desc(sa + ks * 32, 16, kSbo, kLt);
                const uint64_t bh = make_desc(sb + ks * 32, 16, kSbo, kLt);
                const uint32_t acc = (kb != 0 || ks != 0) ? 1u : 0u;
                if (SPLIT) {
                  const uint64_t al = make_desc(sa + aBytes + ks * 32, 16, kSbo, kLt);
                  const uint64_t bl = make_desc(sb + bBytes + ks * 32, 16, kSbo, kLt);
                  mma2<false>(tmem_d, al, bh, idesc, acc);
                  mma2<false>(tmem_d, ah, bl, idesc, 1u);
                  mma2<false>(tmem_d, ah, bh, idesc, 1u);
                } else {
                  mma2<false>(tmem_d, ah, bh, idesc, acc);
                }
              }
            }
            mma2_commit(&ctl.empty[stage]);
            if (kb == kBlocks - 1) mma2_commit(&ctl.tmem_full[as]);
          }
          __syncwarp();
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kTmaWarp) {
    // ===================== TMA: this CTA's A tile and its half of the weight tile =====================
    if (lane == 0) {
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      // bytes the leader's full[] barrier expects per stage: the loads of BOTH CTAs that signal it
      const uint32_t fullTx = 2u * (uint32_t)(bBytes * (SPLIT ? 2 : 1) + (CONVERT ? 0 : aBytes));
      for (int t = pair; t < numTiles; t += numPairs) {
        const int mp = t % mPairs, nt = t / mPairs;
        const int n0 = nt * BN + (int)rank * (BN >> 1);
        const int m0 = (mp * 2 + (int)rank) * BM;   // past the last pixel for the odd tile out: the TMA zero-fills
        const int img = m0 / g.OHW;
        const int rem = m0 - img * g.OHW;
        const int oh0 = rem / g.OW;
        const int ow0 = rem - oh0 * g.OW;
        const int cw = ow0 * g.sx - g.pw, ch = oh0 * g.sy - g.ph;
        int cb = 0, kr = 0, kc = 0;
        for (int kb = 0; kb < kBlocks; ++kb) {
          mbar_wait(&ctl.empty[stage], phase ^ 1);
          const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
          const uint32_t sb = sa + bOff;
          if (rank == 0) mbar_arrive_expect_tx(&ctl.full[stage], fullTx);
          if (CONVERT) {
            mbar_arrive_expect_tx(&ctl.araw[stage], (uint32_t)aBytes);
            tma_load_im2col(sa, &tmA, cb * KB, cw, ch, img, (uint32_t)(kc * g.dx), (uint32_t)(kr * g.dy), &ctl.araw[stage]);
          } else {
            tma2_load_im2col(sa, &tmA, cb * KB, cw, ch, img, (uint32_t)(kc * g.dx), (uint32_t)(kr * g.dy), &ctl.full[stage]);
            if (ASPLIT)
              tma2_load_im2col(sa + aBytes / 2, &tmA2, cb * KB, cw, ch, img, (uint32_t)(kc * g.dx), (uint32_t)(kr * g.dy),
                               &ctl.full[stage]);
          }
          tma2_load_2d(sb, &tmBhi, kb * KB, n0, &ctl.full[stage]);
          if (SPLIT) tma2_load_2d(sb + bBytes, &tmBlo, kb * KB, n0, &ctl.full[stage]);
          if (++cb == g.cBlocks) {
            cb = 0;
            if (++kc == g.kw) { kc = 0; ++kr; }
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: each CTA drains its own 128 TMEM lanes =====================
    const int quad = warp & 3, set = warp >> 2;
    int it = 0;
    uint8_t *stg = smem + (size_t)stages * stageBytes + (size_t)warp * (kStageBufs * kStageTile);
    int sbuf = 0;
    float *scratch = set == 0 ? ctl.epi_sb[quad]
                              : reinterpret_cast<float *>(smem + (size_t)stages * stageBytes +
                                                          (size_t)2 * kEpilogueWarps * (kStageBufs * kStageTile)) + quad * 64;
    for (int t = pair; t < numTiles; t += numPairs, ++it) {
      const int mp = t % mPairs, nt = t / mPairs;
      const int as = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int64_t n0 = (int64_t)nt * BN;
      const int row0 = (mp * 2 + (int)rank) * BM + quad * 32;
      const uint32_t tmem_d = tmem_base + (uint32_t)as * kAccStride + ((uint32_t)(quad * 32) << 16);
      mbar_wait_warp_relaxed(&ctl.tmem_full[as], use & 1, 64);
      tc_fence_after();
      drain_tile(epi, tmem_d, BN, n0, row0, scratch, stg, sbuf, &tmOut, &tmOutHi, &tmOutLo, [&] {
        if (rank == 0) mbar_arrive(&ctl.tmem_empty[as]);
        else mbar_arrive_cluster(&ctl.tmem_empty[as], 0);
      }, StoreRows{row0}, set, epiSets);
    }
    if (EPI::kTmaStore && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512);
}

// bf16 big/small planes of the weights in the kernel's k order:  k = (tap*CB + cb)*32 + ci,  channel = cb*32 + ci
//   hi[f][k] = bf16_rn(W[f][channel][tap]),  lo[f][k] = bf16_rn(W - hi)        (zero for channel >= C)
static __global__ void __launch_bounds__(256) k_prepare_weights_bf16(const float *__restrict__ w, int rows, int chans, int kk,
                                                                     int CB, uint16_t *__restrict__ hi,
                                                                     uint16_t *__restrict__ lo) {
  const int Kpad = kk * CB * 32;
  const int64_t total = (int64_t)rows * Kpad;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int r = (int)(i / Kpad);
    const int k = (int)(i - (int64_t)r * Kpad);
    const int ci = k & 31, g = k >> 5;
    const int tap = g / CB, cb = g - tap * CB;
    const int ch = cb * 32 + ci;
    const float v = ch < chans ? __ldg(w + ((int64_t)r * chans + ch) * kk + tap) : 0.f;
    const uint32_t h = pack_bf16x2(v, 0.f) & 0xFFFFu;
    const float hf = __uint_as_float(h << 16);
    hi[i] = (uint16_t)h;
    lo[i] = (uint16_t)(pack_bf16x2(v - hf, 0.f) & 0xFFFFu);
  }
}

// bf16 matrix [rows][Kpad] (k contiguous) -> tiles of 32 k x boxRows rows, 64-byte rows: SWIZZLE_64B
static int encode_weights_map_bf16(CUtensorMap *tm, const uint16_t *base, int64_t rows, int64_t Kpad, int boxRows, int kb) {
  EncodeTiledFn enc = get_encode_tiled();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
  cuuint32_t box[2] = {(cuuint32_t)kb, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeTiled (bf16 weights) failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

// activations [B][H][W][pitch] as the im2col-mode tensor {C, W, H, N}: box = 128 pixels x kb channels of one tap
// (bf16 = one plane of a pre-split tensor: 2-byte elements, 64-byte rows for 32 channels -> SWIZZLE_64B)
static int encode_im2col_map(CUtensorMap *tm, const void *base, const simt::ConvGeom &g, int batch, int64_t pitch, int kb,
                             bool bf16 = false) {
  EncodeIm2colFn enc = get_encode_im2col();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeIm2col is not available from the driver");
  const cuuint64_t es = bf16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * g.W * es, (cuuint64_t)pitch * g.W * g.H * es};
  // bounding box of the tap-(0,0) position of an output pixel (the convention of the im2col mode: lower corner
  // = -padding, upper corner = padding - (k-1)*dilation, measured from the far edge)
  int lower[2] = {-g.pw, -g.ph};
  int upper[2] = {g.pw - (g.kw - 1) * g.dx, g.ph - (g.kh - 1) * g.dy};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.sx, (cuuint32_t)g.sy, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims,
                   strides, lower, upper, (cuuint32_t)kb, BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kb * es == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeIm2col failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

// output rows [M][pitch] (channels innermost): 32 x 32 store boxes, SWIZZLE_128B staging
// (bf16 = one plane of a pre-split tensor: the same boxes with 64-byte rows, SWIZZLE_64B staging)
static int encode_out_map(CUtensorMap *tm, void *base, int64_t M, int64_t F, int64_t pitch, bool bf16 = false) {
  EncodeTiledFn enc = get_encode_tiled();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)F, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeTiled (NHWC output) failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

struct Maps {
  alignas(64) CUtensorMap a, a2, bhi, blo, out, outHi, outLo;
};

// launch configuration with programmatic stream serialization (see pdl_wait) unless TNN_PDL=0
static cudaLaunchConfig_t pdl_config(tnn_context *ctx, int grid, int threads, size_t smemBytes, cudaLaunchAttribute *at) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smemBytes;
  cfg.stream = ctx->stream;
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = ctx->pdl ? 1 : 0;
  return cfg;
}

template <int MODE, int KB, bool ASPLIT, typename EPI>
static int launch_nhwc(tnn_context *ctx, const Plan &p, const Geo &geo, int epiSets, EPI epi, const Maps &m) {
  auto kern = k_conv_nhwc<MODE, KB, ASPLIT, EPI>;
  TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
  cudaLaunchAttribute at[1];
  cudaLaunchConfig_t cfg = pdl_config(ctx, p.grid, kThreads, p.smemBytes, at);
  TNN_CUDA(cudaLaunchKernelEx(&cfg, kern, epi, p.sched, geo, p.BN, p.stages, p.accCols, epiSets, m.a, m.a2, m.bhi, m.blo, m.out,
                              m.outHi, m.outLo));
  TNN_LAUNCHED(ctx);
  return 0;
}

template <int MODE, int KB, bool ASPLIT, typename EPI>
static int launch_nhwc_pair(tnn_context *ctx, int grid, size_t smemBytes, int mTiles, int nTiles, int kBlocks, const Geo &geo,
                            int BN, int stages, int epiSets, EPI epi, const Maps &m) {
  auto kern = k_conv_nhwc_pair<MODE, KB, ASPLIT, EPI>;
  TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  cudaLaunchAttribute at[1];
  cudaLaunchConfig_t cfg = pdl_config(ctx, grid, kThreads, smemBytes, at);
  TNN_CUDA(cudaLaunchKernelEx(&cfg, kern, epi, mTiles, nTiles, kBlocks, geo, BN, stages, epiSets, m.a, m.a2, m.bhi, m.blo, m.out,
                              m.outHi, m.outLo));
  TNN_LAUNCHED(ctx);
  return 0;
}

// One tensor of the convolution: fp32 (rows of channels or planes) and / or the pre-split bf16 planes.
struct SplitRef {
  const uint16_t *hi;       // nullptr: not present
  int64_t pitch, plane;     // bf16 elements between pixels / between the hi and the lo plane
};

// in: fp32 rows (inSplit.hi == nullptr) or pre-split planes.  Outputs: `out` fp32 (rows of channels, or planes when
// outNchw) and / or outSplit (pre-split planes, rows of channels only).
static int conv_forward(tnn_context *ctx, const simt::ConvGeom &g, int batch, const float *in, int64_t inPitch,
                        SplitRef inSplit, const float *weights, const float *scale, const float *bias, int act, float *out,
                        int64_t outPitch, bool outNchw, SplitRef outSplit, const float *residual = nullptr,
                        int64_t resPitch = 0, SplitRef resSplit = SplitRef{nullptr, 0, 0}, int yoloEntries = 0,
                        float yoloScaleXY = 1.0f) {
  const int kk = g.kh * g.kw;
  const bool aSplit = inSplit.hi != nullptr;
  if (aSplit) {
    TNN_REQUIRE(ctx->math_mode == TNN_MATH_BF16X3, "pre-split (bf16 hi/lo) activations are the BF16x3 mode's operands");
    TNN_REQUIRE((inSplit.pitch % 8) == 0 && inSplit.pitch >= g.C && (inSplit.plane % 8) == 0 &&
                    (((uintptr_t)inSplit.hi) & 15) == 0,
                "pre-split input needs a pixel pitch and a plane stride that are multiples of 8 and a 16-byte aligned base");
  } else {
    TNN_REQUIRE((inPitch % 4) == 0 && inPitch >= g.C && (((uintptr_t)in) & 15) == 0,
                "NHWC input needs a pixel pitch that is a multiple of 4 floats and a 16-byte aligned base");
  }
  TNN_REQUIRE(g.sx <= 8 && g.sy <= 8 && g.pw <= 128 && g.ph <= 128 && (g.kw - 1) * g.dx - g.pw <= 128 &&
                  (g.kh - 1) * g.dy - g.ph <= 128, "geometry outside the TMA im2col limits");
  const int64_t M = (int64_t)batch * g.OH * g.OW;
  const int64_t N = g.F;
  TNN_REQUIRE(M < (1ll << 31) - BM, "too many output pixels");
  const bool bf = ctx->math_mode == TNN_MATH_BF16X3;
  const bool split = tnn_math_is_split(ctx);
  const int CB = (g.C + 31) / 32;       // 32-channel blocks per tap (the weight copy's k order)
  const int64_t K = (int64_t)kk * CB * 32;
  // 3xTF32 stages carry four fp32 tiles: half-depth (16-channel) stages keep four of them in flight
  // (bf16 stages are half the bytes per channel: 32-channel stages measured 13.5 vs 15.3 ms over the YOLOv3 layers)
  // a layer with at most 16 channels (the image layer) moves half the zero padding per tap with 16-channel stages
  const int KB = aSplit ? 32 : split ? (ctx->nhwc_kb ? ctx->nhwc_kb : ((bf && g.C > 16) ? 32 : 16)) : 32;
  Plan p = make_plan(ctx, M, N, K, split, false, false);
  // Wave quantisation: with several filter tiles per pixel tile the width of a tile is a free parameter.  172 tiles of
  // 256 filters on 74 CTA pairs are 2.3 waves = 3; 215 tiles of 224 filters are 2.9 waves = 3 of cheaper tiles.  A tile's
  // cost is modelled as 1/4 fixed (activation traffic, epilogue set-up) + 3/4 proportional to its width.
  if (ctx->nhwc_pair && ctx->nhwc_bn_auto && N > 256 && p.BN == 256) {
    const int64_t pairs = ctx->prop.multiProcessorCount / 2, mPairs = (p.sched.mTiles + 1) / 2;
    double best = 0;
    int bestBN = 256;
    for (int bn = 256; bn >= 160; bn -= 32) {
      const int64_t tiles = mPairs * cdiv64(N, bn);
      if (tiles < pairs) continue;
      const double cost = (double)cdiv64(tiles, pairs) * (0.25 + 0.75 * bn / 256.0);
      if (best == 0 || cost < best * 0.97) { best = cost; bestBN = bn; }   // narrower only for a clear (3 %) gain
    }
    if (bestBN != 256) {
      p.BN = bestBN;
      p.sched.nTiles = (int)cdiv64(N, bestBN);
      p.sched.numTiles = p.sched.mTiles * p.sched.nTiles * p.sched.kSplits;
      p.sched.wideTiles = p.sched.numTiles;
    }
  }
  // half-width tail tiles: the TMA store writes whole 32-filter boxes, so a half tile must be a multiple of 32 wide
  // CTA pairs (cta_group::2) halve the weight traffic through shared memory — the bound of the single-CTA kernel.
  // Taken when the tile is wide enough to split (each CTA holds BN/2 >= 32 weight rows) and there are pair tiles for
  // every TPC; TNN_NHWC_PAIR=0 keeps the single-CTA kernel (A/B runs).
  const int sms = ctx->prop.multiProcessorCount;
  const int64_t pairTiles = (int64_t)((p.sched.mTiles + 1) / 2) * p.sched.nTiles;
  const bool pairMode = ctx->nhwc_pair && (p.BN % 32) == 0 && p.BN >= 64 && pairTiles >= sms / 2;
  const int bnCta = pairMode ? p.BN / 2 : p.BN;   // weight rows a CTA stages
  if (!pairMode && (outNchw || (p.BN % 64) == 0)) split_tail_wave(ctx, p);
  const size_t stageBytes = bf ? (size_t)BM * KB * (aSplit ? 4 : 8) + (size_t)bnCta * KB * 4
                               : (size_t)(BM + bnCta) * KB * 4 * (split ? 2 : 1);
  // Two sets of epilogue warps (CTA pairs, TMA-only modes): layers whose main loop per tile is short against the drain
  // of its chunks — every 1x1 layer and the 3x3 layers up to 64 channels (K <= 1024) — are paced by the epilogue.
  // The second set costs 32 KB of staging: one pipeline stage of the widest tiles.  TNN_NHWC_EPI_SETS=1/2 forces it.
  const bool convert = split && !aSplit;
  int epiSets = 1;
  // (stride-2 layers are paced by their input traffic, not by the drain: measured 1-6 % slower with two sets)
  if (pairMode && !convert && !outNchw)
    epiSets = ctx->nhwc_epi_sets ? ctx->nhwc_epi_sets : ((K <= 1024 && g.sx == 1 && g.sy == 1) ? 2 : 1);
  // the single-CTA kernel (tiles too narrow for a pair: 64 -> 32 @208^2) alternates TILES between its two sets
  if (!pairMode && !convert && !outNchw && p.BN <= 64 && p.sched.kSplits == 1)
    epiSets = ctx->nhwc_epi_sets ? ctx->nhwc_epi_sets : 2;
  const size_t epiBytes = outNchw ? 0 : (size_t)epiSets * kEpilogueWarps * kStageBufs * kStageTile + (epiSets == 2 ? 1024 : 0);
  int st = (int)((224 * 1024 - epiBytes) / stageBytes);   // stages + staging + 1 KB of alignment slack = at most 225 KB
  p.stages = st > kMaxStages ? kMaxStages : st;
  TNN_REQUIRE(p.stages >= 2, "tile does not fit shared memory");
  p.smemBytes = stageBytes * p.stages + epiBytes + 1024;
  p.sched.kBlocks = p.sched.kBlocksPerSplit = (int)(K / KB);
  float *whi, *wlo;
  bool fresh;
  // bf16 planes take half the floats of the fp32 ones (N*K is a multiple of 32)
  int rc = tnn_weight_buffer(ctx, weights, 1, bf ? (size_t)N * K / 2 : (size_t)N * K, &whi, &wlo, &fresh);
  if (rc) return rc;
  if (fresh) {
    int64_t blocks = cdiv64(N * K, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (bf)
      k_prepare_weights_bf16<<<(int)blocks, 256, 0, ctx->stream>>>(weights, (int)N, g.C, kk, CB, (uint16_t *)whi, (uint16_t *)wlo);
    else
      k_prepare_weights_patch<<<(int)blocks, 256, 0, ctx->stream>>>(weights, (int)N, g.C, kk, CB, 0, whi, wlo, split ? 1 : 0);
    TNN_LAUNCHED(ctx);
  }
  Maps m;
  memset(&m, 0, sizeof(m));
  if (aSplit) {
    rc = encode_im2col_map(&m.a, inSplit.hi, g, batch, inSplit.pitch, KB, true);
    if (rc) return rc;
    rc = encode_im2col_map(&m.a2, inSplit.hi + inSplit.plane, g, batch, inSplit.pitch, KB, true);
  } else {
    rc = encode_im2col_map(&m.a, in, g, batch, inPitch, KB);
  }
  if (rc) return rc;
  if (bf) {
    rc = encode_weights_map_bf16(&m.bhi, (const uint16_t *)whi, N, K, bnCta, KB);
    if (rc) return rc;
    rc = encode_weights_map_bf16(&m.blo, (const uint16_t *)wlo, N, K, bnCta, KB);
  } else {
    rc = KB == 16 ? encode_weights_map16(&m.bhi, whi, N, K, bnCta) : encode_weights_map(&m.bhi, whi, N, K, bnCta);
    if (rc) return rc;
    rc = KB == 16 ? encode_weights_map16(&m.blo, split ? wlo : whi, N, K, bnCta)
                  : encode_weights_map(&m.blo, split ? wlo : whi, N, K, bnCta);
  }
  if (rc) return rc;
  Geo geo{g.OW, g.OH * g.OW, g.sx, g.sy, g.pw, g.ph, g.dx, g.dy, g.kw, CB * 32 / KB, (int)M};
  ctx->last_conv_kernel = pairMode ? 11 : 7;
  auto go = [&](auto epi) -> int {
    if (pairMode) {
      const int grid = (int)(2 * (pairTiles < sms / 2 ? pairTiles : sms / 2));
      const int kBlocks = (int)(K / KB);
#define TNN_PAIR_GO(MODE, KBV, AS) \
  launch_nhwc_pair<MODE, KBV, AS>(ctx, grid, p.smemBytes, p.sched.mTiles, p.sched.nTiles, kBlocks, geo, p.BN, p.stages, epiSets, epi, m)
      if (aSplit) return TNN_PAIR_GO(MODE_BF16X3, 32, true);
      if (bf) return KB == 16 ? TNN_PAIR_GO(MODE_BF16X3, 16, false) : TNN_PAIR_GO(MODE_BF16X3, 32, false);
      if (!split) return TNN_PAIR_GO(MODE_TF32, 32, false);
      return KB == 16 ? TNN_PAIR_GO(MODE_3X, 16, false) : TNN_PAIR_GO(MODE_3X, 32, false);
#undef TNN_PAIR_GO
    }
    if (aSplit) return launch_nhwc<MODE_BF16X3, 32, true>(ctx, p, geo, epiSets, epi, m);
    if (bf) return KB == 16 ? launch_nhwc<MODE_BF16X3, 16, false>(ctx, p, geo, epiSets, epi, m)
                            : launch_nhwc<MODE_BF16X3, 32, false>(ctx, p, geo, epiSets, epi, m);
    if (!split) return launch_nhwc<MODE_TF32, 32, false>(ctx, p, geo, epiSets, epi, m);
    return KB == 16 ? launch_nhwc<MODE_3X, 16, false>(ctx, p, geo, epiSets, epi, m) : launch_nhwc<MODE_3X, 32, false>(ctx, p, geo, epiSets, epi, m);
  };
  if (outNchw) {
    TNN_REQUIRE(out != nullptr && outSplit.hi == nullptr, "planar output is fp32 only");
    NchwEpilogue epi{{out, scale, bias, g.OH * g.OW, g.outImg, act, M, N, false, 0, 0, 0, 0, 0, 0, 0, 0, nullptr},
                     yoloEntries, yoloScaleXY, 0.5f * (1.0f - yoloScaleXY)};
    return go(epi);
  }
  TNN_REQUIRE(out != nullptr || outSplit.hi != nullptr, "no output tensor");
  if (out) {
    TNN_REQUIRE((outPitch % 4) == 0 && outPitch >= N && (((uintptr_t)out) & 15) == 0,
                "NHWC output needs a pixel pitch that is a multiple of 4 floats and a 16-byte aligned base");
    rc = encode_out_map(&m.out, out, M, N, outPitch);
    if (rc) return rc;
  }
  if (outSplit.hi) {
    TNN_REQUIRE((outSplit.pitch % 8) == 0 && outSplit.pitch >= N && (outSplit.plane % 8) == 0 &&
                    (((uintptr_t)outSplit.hi) & 15) == 0,
                "pre-split output needs a pixel pitch and a plane stride that are multiples of 8 and a 16-byte aligned base");
    rc = encode_out_map(&m.outHi, (void *)outSplit.hi, M, N, outSplit.pitch, true);
    if (rc) return rc;
    rc = encode_out_map(&m.outLo, (void *)(outSplit.hi + outSplit.plane), M, N, outSplit.pitch, true);
    if (rc) return rc;
  }
  if (residual)
    TNN_REQUIRE((resPitch % 4) == 0 && resPitch >= N && (N % 4) == 0 && (((uintptr_t)residual) & 15) == 0,
                "the residual needs 16-byte aligned rows and a filter count that is a multiple of 4");
  if (resSplit.hi)
    TNN_REQUIRE(residual == nullptr && (resSplit.pitch % 8) == 0 && resSplit.pitch >= N && (resSplit.plane % 8) == 0 &&
                    (N % 8) == 0 && (((uintptr_t)resSplit.hi) & 15) == 0,
                "a pre-split residual needs pitch, plane stride and filter count multiples of 8 and a 16-byte aligned base");
  NhwcEpilogue epi{scale, bias, act, M, N, residual, resPitch, resSplit.hi, resSplit.pitch, resSplit.plane,
                   out != nullptr, outSplit.hi != nullptr};
  return go(epi);
}

// =====================================================================================================================
// The image layer (planar NCHW input, C*kh*kw <= 32: 3 channels x 3x3 = 27) on the tensor cores.
// Its contraction is ONE 32-deep k-block, k = (c, kr, kc) in the weights' own order.  A tile is a patch of 4 x 32 output
// pixels of one image:
//   TMA     one tiled 4-D load fetches the patch's input HALO from the planes ({W, H, C, N} box, columns from a 16-byte
//           aligned start, padding / image edges zero-filled by the TMA unit) — a few KB of fp32
//   gather  thread = one pixel: 27 shared-memory reads at fixed offsets (no address arithmetic, no bounds checks),
//           split to bf16 hi/lo, four 16-byte stores per plane into the K-major SWIZZLE_64B tile
//   MMA     six 128 x F x 16 MMAs (BF16x3) into one slot of a ring of up to eight TMEM accumulators
//   out     the usual epilogue; stores are 4-D ({F, OW, OH, N} boxes of 32 filters x 32 columns x 1 row), so ragged
//           patches are clipped by the TMA unit
// The work per tile is tiny and every role's per-tile instruction chain is latency bound, so everything is about keeping
// many tiles in flight: the twelve worker warps form three sets of four (one warp per TMEM lane quadrant); a set gathers
// tile i, then drains tile i - 3 (whose MMAs have finished meanwhile), then gathers tile i + 3, ... — all twelve warps
// always have independent work.  Halos / operand tiles / accumulators are rings, and the weights (F x 32 bf16 x 2
// planes) stay in shared memory for the whole kernel.  The CUDA-core direct kernel needs 1.2 ms for YOLOv3's first layer at batch 64
// (864 FMAs per pixel); a first tensor-core version that gathered from global memory needed 1.0 ms (5700 warp
// instructions per tile, most of them address arithmetic and edge tests, ncu).
constexpr int kImgStages = 4;      // operand tiles
constexpr int kImgHalos = 4;       // halo ring
constexpr int kImgHaloBytes = 8192;
#ifndef TNN_IMG_SETS
#define TNN_IMG_SETS 3
#endif
constexpr int kImgSets = TNN_IMG_SETS;   // worker sets
constexpr int kImgSetWarps = 4;    // 4 warps = 128 threads = one pixel each = the four TMEM lane quadrants
constexpr int kImgWorkers = kImgSets * kImgSetWarps;   // then one MMA warp and one TMA warp
constexpr int kImgMmaWarp = kImgWorkers, kImgTmaWarp = kImgWorkers + 1;
constexpr int kImgThreads = (kImgWorkers + 2) * 32;
constexpr int kImgMaxAcc = 8;
constexpr int kImgTW = 32, kImgTH = 4;   // output patch: one warp = one patch row, so its halo reads are 32 consecutive floats

struct ImgCtl {
  uint64_t h_full[kImgHalos], h_empty[kImgHalos];
  uint64_t a_full[kImgStages], a_empty[kImgStages];
  uint64_t tmem_full[kImgMaxAcc], tmem_empty[kImgMaxAcc];
  uint32_t tmem_base;
  alignas(16) int ktab[32];   // k -> byte offset of tap (c, kr, kc) from a pixel's corner in the halo; k >= K: -1
  alignas(16) float epi_sb[kImgWorkers][64];
};
struct ImgGeo {
  int tilesX, tilesY;     // patches per image row / column
  int PC, PR;             // halo box: columns (multiple of 4), rows
  int delta;              // columns between the aligned box start and the first tap column
  int haloBytes;
  int haloSlot;           // bytes between the halos of the ring (haloBytes rounded up to 128)
};

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *tmap, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__global__ void __launch_bounds__(kImgThreads, 1)
k_conv_image(NhwcEpilogue epi, simt::ConvGeom g, ImgGeo ig, const float *__restrict__ w, int BN, int nAcc, uint32_t accCols,
             int numTiles, const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
             const __grid_constant__ CUtensorMap tmOutHi, const __grid_constant__ CUtensorMap tmOutLo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ ImgCtl ctl;
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the shared address space
  // and emits LDS / STS instead of generic LD / ST for every access below)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int aPlane = BM * 64;                 // one bf16 plane of a tile: 128 rows x 64 bytes
  constexpr int aStage = 2 * aPlane;
  const int bPlane = BN * 64;
  uint8_t *sB = smem + kImgStages * aStage;       // B_hi | B_lo
  uint8_t *sStg = sB + 2 * bPlane;                // epilogue staging (1024-aligned: BN is a multiple of 16)
  uint8_t *sHalo = sStg + kImgWorkers * kStageBufs * kStageTile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = g.C * g.kh * g.kw, kk = g.kh * g.kw;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kImgHalos; ++s) {
      mbar_init(&ctl.h_full[s], 1);
      mbar_init(&ctl.h_empty[s], kImgSetWarps);
    }
    for (int s = 0; s < kImgStages; ++s) {
      mbar_init(&ctl.a_full[s], kImgSetWarps);
      mbar_init(&ctl.a_empty[s], 1);
    }
    for (int s = 0; s < kImgMaxAcc; ++s) {
      mbar_init(&ctl.tmem_full[s], 1);
      mbar_init(&ctl.tmem_empty[s], kImgSetWarps);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    const int k = threadIdx.x;
    const int c = k / kk, tap = k - c * kk, kr = tap / g.kw, kc = tap - kr * g.kw;
    ctl.ktab[k] = k < K ? ((c * ig.PR + kr * g.dy) * ig.PC + kc * g.dx) * 4 : -1;
  }
  // weights -> bf16 hi / lo K-major SWIZZLE_64B tiles (row n = filter, 64 bytes = 32 k; rows >= F and k >= K zero)
  for (int u = threadIdx.x; u < BN * 4; u += kImgThreads) {
    const int n = u >> 2, j = u & 3;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = j * 8 + e;
      v[e] = (n < g.F && k < K) ? __ldg(w + (int64_t)n * K + k) : 0.f;
    }
    uint4 hi, lo;
    split_bf16x8(make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]), hi, lo);
    const int o = n * 64 + ((j ^ ((n >> 1) & 3)) << 4);
    *reinterpret_cast<uint4 *>(sB + o) = hi;
    *reinterpret_cast<uint4 *>(sB + bPlane + o) = lo;
  }
  if (warp == kImgMmaWarp) tmem_alloc(&ctl.tmem_base, 512);
  if (warp == kImgTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmIn);
    if (epi.f32Out) tma_prefetch_desc(&tmOut);
    if (epi.splitOut) { tma_prefetch_desc(&tmOutHi); tma_prefetch_desc(&tmOutLo); }
  }
  fence_proxy_async();   // the weight tiles are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl.tmem_base;
  const int tilesPerImg = ig.tilesX * ig.tilesY;

  if (warp < kImgWorkers) {
    // ===================== workers: gather tile i, drain tile i - 3 =====================
    const int set = warp / kImgSetWarps, quad = warp & 3;
    const int p = threadIdx.x - set * kImgSetWarps * 32;   // row of the operand tile = pixel of the patch
    const int pr = p / kImgTW, pc = p - pr * kImgTW;
    const int corner = ((pr * g.sy) * ig.PC + pc * g.sx + ig.delta) * 4;   // this pixel's tap (0, 0, 0) in the halo, bytes
    const int sw = (p >> 1) & 3;
    uint8_t *stg = sStg + (size_t)warp * (kStageBufs * kStageTile);
    float *scratch = ctl.epi_sb[warp];
    int sbuf = 0;

    auto gather = [&](int it) {
      const int stage = it % kImgStages, hs = it % kImgHalos;
      mbar_wait_warp(&ctl.h_full[hs], (uint32_t)((it / kImgHalos) & 1));
      const uint8_t *halo = sHalo + hs * ig.haloSlot + corner;
      uint4 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int4 o = *reinterpret_cast<const int4 *>(&ctl.ktab[8 * j + 4 * q]);
          x[4 * q + 0] = o.x >= 0 ? *reinterpret_cast<const float *>(halo + o.x) : 0.f;
          x[4 * q + 1] = o.y >= 0 ? *reinterpret_cast<const float *>(halo + o.y) : 0.f;
          x[4 * q + 2] = o.z >= 0 ? *reinterpret_cast<const float *>(halo + o.z) : 0.f;
          x[4 * q + 3] = o.w >= 0 ? *reinterpret_cast<const float *>(halo + o.w) : 0.f;
        }
        split_bf16x8(make_float4(x[0], x[1], x[2], x[3]), make_float4(x[4], x[5], x[6], x[7]), hi[j], lo[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.h_empty[hs]);   // the halo is in registers
      mbar_wait_warp(&ctl.a_empty[stage], (uint32_t)(((it / kImgStages) & 1) ^ 1));
      uint8_t *shi = smem + stage * aStage, *slo = shi + aPlane;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = p * 64 + ((j ^ sw) << 4);
        *reinterpret_cast<uint4 *>(shi + o) = hi[j];
        *reinterpret_cast<uint4 *>(slo + o) = lo[j];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.a_full[stage]);
    };
    auto drain = [&](int it) {
      const int t = blockIdx.x + it * gridDim.x;
      const int as = it % nAcc;
      const uint32_t use = (uint32_t)(it / nAcc);
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols + ((uint32_t)(quad * 32) << 16);
      const int b = t / tilesPerImg, rem = t - b * tilesPerImg, ty = rem / ig.tilesX, tx = rem - ty * ig.tilesX;
      const int ow0 = tx * kImgTW, oh0 = ty * kImgTH + quad;   // this warp's patch row
      mbar_wait_warp(&ctl.tmem_full[as], use & 1);
      tc_fence_after();
      drain_tile(epi, tmem_d, BN, 0, 0, scratch, stg, sbuf, &tmOut, &tmOutHi, &tmOutLo,
                 [&] { mbar_arrive(&ctl.tmem_empty[as]); },
                 [&](const CUtensorMap *tm, uint32_t src, int n) { tma_store_4d(tm, src, n, ow0, oh0, b); });
    };
    const int nIt = (numTiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // this CTA's tiles
    int prev = -1;
    for (int it = set; it < nIt; it += kImgSets) {
      gather(it);
      if (prev >= 0) drain(prev);
      prev = it;
    }
    if (prev >= 0) drain(prev);
    if (lane == 0) bulk_wait_all();
  } else if (warp == kImgMmaWarp) {
    // ===================== MMA issuer: six MMAs per tile, accumulator ring =====================
    const uint32_t idesc = make_idesc_bf16(BN);
    const uint32_t sb = smem_u32(sB);
    int it = 0;
    for (int t = blockIdx.x; t < numTiles; t += gridDim.x, ++it) {
      const int stage = it % kImgStages;
      const uint32_t phase = (uint32_t)((it / kImgStages) & 1);
      const int as = it % nAcc;
      const uint32_t use = (uint32_t)(it / nAcc);
      mbar_wait_warp(&ctl.tmem_empty[as], (use & 1) ^ 1);
      mbar_wait_warp(&ctl.a_full[stage], phase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols;
        const uint32_t shi = smem_u32(smem) + (uint32_t)stage * aStage, slo = shi + aPlane;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t ah = make_desc(shi + ks * 32, 16, 512, 4u), al = make_desc(slo + ks * 32, 16, 512, 4u);
          const uint64_t bh = make_desc(sb + ks * 32, 16, 512, 4u), bl = make_desc(sb + bPlane + ks * 32, 16, 512, 4u);
          mma_bf16(tmem_d, al, bh, idesc, ks ? 1u : 0u);
          mma_bf16(tmem_d, ah, bl, idesc, 1u);
          mma_bf16(tmem_d, ah, bh, idesc, 1u);
        }
        mma_commit(&ctl.a_empty[stage]);
        mma_commit(&ctl.tmem_full[as]);
      }
      __syncwarp();
    }
  } else if (warp == kImgTmaWarp) {
    // ===================== TMA: one halo per tile =====================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < numTiles; t += gridDim.x, ++it) {
        const int hs = it % kImgHalos;
        mbar_wait(&ctl.h_empty[hs], (uint32_t)(((it / kImgHalos) & 1) ^ 1));
        const int b = t / tilesPerImg, rem = t - b * tilesPerImg, ty = rem / ig.tilesX, tx = rem - ty * ig.tilesX;
        mbar_arrive_expect_tx(&ctl.h_full[hs], (uint32_t)ig.haloBytes);
        tma_load_4d(smem_u32(sHalo + hs * ig.haloSlot), &tmIn, tx * kImgTW * g.sx - g.pw - ig.delta,
                    ty * kImgTH * g.sy - g.ph, 0, b, &ctl.h_full[hs]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kImgMmaWarp) tmem_dealloc(tmem_base, 512);
}

static bool image_geo(const simt::ConvGeom &g, ImgGeo &ig) {
  // first tap column of a patch = X - pw with X a multiple of 16; the box starts `delta` columns earlier on a multiple
  // of 4 (tiled TMA wants a 16-byte aligned start):  (X - pw - delta) % 4 == 0  <=>  delta = (-pw) mod 4
  ig.delta = ((-g.pw) % 4 + 4) % 4;
  const int cols = (kImgTW - 1) * g.sx + (g.kw - 1) * g.dx + 1 + ig.delta;
  ig.PC = (cols + 3) / 4 * 4;
  ig.PR = (kImgTH - 1) * g.sy + (g.kh - 1) * g.dy + 1;
  ig.haloBytes = ig.PC * ig.PR * g.C * 4;
  ig.haloSlot = (ig.haloBytes + 127) / 128 * 128;
  ig.tilesX = (g.OW + kImgTW - 1) / kImgTW;
  ig.tilesY = (g.OH + kImgTH - 1) / kImgTH;
  return ig.PC <= 256 && ig.PR <= 256 && g.C <= 256 && ig.haloBytes <= kImgHaloBytes;
}

static bool image_supported(const simt::ConvGeom &g, const float *in) {
  ImgGeo ig;
  return g.C * g.kh * g.kw <= 32 && g.F <= 256 && (g.W % 4) == 0 && (((uintptr_t)in) & 15) == 0 && image_geo(g, ig);
}

// 4-D output map {F, OW, OH, N} over rows of channels: boxes of 32 filters x 32 columns x 1 row (= 32 staged pixel rows)
static int encode_out_map4(CUtensorMap *tm, void *base, const simt::ConvGeom &g, int batch, int64_t pitch, bool bf16) {
  EncodeTiledFn enc = get_encode_tiled();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t es = bf16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)g.F, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * g.OW * es, (cuuint64_t)pitch * g.OW * g.OH * es};
  cuuint32_t box[4] = {32, kImgTW, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeTiled (image layer output) failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

static int conv_image(tnn_context *ctx, const simt::ConvGeom &g, int batch, const float *in, const float *weights,
                      const float *scale, const float *bias, int act, float *out, int64_t outPitch, SplitRef outSplit) {
  const int64_t N = g.F;
  TNN_REQUIRE(out != nullptr || outSplit.hi != nullptr, "no output tensor");
  ImgGeo ig;
  TNN_REQUIRE(image_geo(g, ig), "image layer geometry outside the halo limits");
  const int BN = (int)((N + 15) / 16 * 16);
  const uint32_t accCols = pow2_cols(BN);
  int nAcc = (int)(512 / accCols);
  if (nAcc > kImgMaxAcc) nAcc = kImgMaxAcc;
  Maps m;
  memset(&m, 0, sizeof(m));
  int rc;
  {
    EncodeTiledFn enc = get_encode_tiled();
    TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.C, (cuuint64_t)batch};
    cuuint64_t strides[3] = {(cuuint64_t)g.W * 4, (cuuint64_t)g.W * g.H * 4, (cuuint64_t)g.inImg * 4};
    cuuint32_t box[4] = {(cuuint32_t)ig.PC, (cuuint32_t)ig.PR, (cuuint32_t)g.C, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&m.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)in, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      tnn_set_error("cuTensorMapEncodeTiled (image planes) failed with CUresult %d", (int)r);
      return 6;
    }
  }
  if (out) {
    TNN_REQUIRE((outPitch % 4) == 0 && outPitch >= N && (((uintptr_t)out) & 15) == 0,
                "NHWC output needs a pixel pitch that is a multiple of 4 floats and a 16-byte aligned base");
    rc = encode_out_map4(&m.out, out, g, batch, outPitch, false);
    if (rc) return rc;
  }
  if (outSplit.hi) {
    TNN_REQUIRE((outSplit.pitch % 8) == 0 && outSplit.pitch >= N && (outSplit.plane % 8) == 0 &&
                    (((uintptr_t)outSplit.hi) & 15) == 0,
                "pre-split output needs a pixel pitch and a plane stride that are multiples of 8 and a 16-byte aligned base");
    rc = encode_out_map4(&m.outHi, (void *)outSplit.hi, g, batch, outSplit.pitch, true);
    if (rc) return rc;
    rc = encode_out_map4(&m.outLo, (void *)(outSplit.hi + outSplit.plane), g, batch, outSplit.pitch, true);
    if (rc) return rc;
  }
  // rows = 1 / row0 = 0: the 4-D stores clip ragged patches themselves
  NhwcEpilogue epi{scale, bias, act, 1, N, nullptr, 0, nullptr, 0, 0, out != nullptr, outSplit.hi != nullptr};
  const int64_t tiles = (int64_t)batch * ig.tilesX * ig.tilesY;
  TNN_REQUIRE(tiles < (1ll << 30), "too many patches");
  const int numTiles = (int)tiles;
  const int sms = ctx->prop.multiProcessorCount;
  const int grid = numTiles < sms ? numTiles : sms;
  const size_t smemBytes = (size_t)kImgStages * 2 * BM * 64 + (size_t)2 * BN * 64 +
                           (size_t)kImgWorkers * kStageBufs * kStageTile + (size_t)kImgHalos * ig.haloSlot + 1024;
  TNN_CUDA(cudaFuncSetAttribute(k_conv_image, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  k_conv_image<<<grid, kImgThreads, smemBytes, ctx->stream>>>(epi, g, ig, weights, BN, nAcc, accCols, numTiles, m.a, m.out,
                                                           m.outHi, m.outLo);
  TNN_LAUNCHED(ctx);
  return 0;
}

// ------------------------------------------------------------------ layout companions (HBM-bound)
// [B][C][HW] <-> [B][HW][pitch]: 32 x 32 tiles through shared memory, both sides coalesced
__global__ void __launch_bounds__(256) k_nchw_to_nhwc(const float *__restrict__ in, float *__restrict__ out, int C, int HW,
                                                      int64_t pitch) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float *src = in + (int64_t)b * C * HW;
  float *dst = out + (int64_t)b * HW * pitch;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + tx;
    t[j][tx] = (c < C && p < HW) ? __ldg(src + (int64_t)c * HW + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    if (p < HW && c < C) dst[(int64_t)p * pitch + c] = t[tx][j];
  }
}
__global__ void __launch_bounds__(256) k_nhwc_to_nchw(const float *__restrict__ in, float *__restrict__ out, int C, int HW,
                                                      int64_t pitch) {
  __shared__ float t[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float *src = in + (int64_t)b * HW * pitch;
  float *dst = out + (int64_t)b * C * HW;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + tx;
    t[j][tx] = (p < HW && c < C) ? __ldg(src + (int64_t)p * pitch + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + tx;
    if (c < C && p < HW) dst[(int64_t)c * HW + p] = t[tx][j];
  }
}

// rows of C floats with independent pitches: out = a (+ b).  float4 when everything is 16-byte aligned.
template <bool ADD>
__global__ void __launch_bounds__(256) k_rows(int64_t rows, int C4, const float4 *__restrict__ a, int64_t lda4,
                                              const float4 *__restrict__ b, int64_t ldb4, float4 *__restrict__ out,
                                              int64_t ldo4) {
  const int64_t total = rows * C4;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int64_t r = i / C4;
    const int c = (int)(i - r * C4);
    float4 v = __ldg(a + r * lda4 + c);
    if (ADD) {
      const float4 w = __ldg(b + r * ldb4 + c);
      v = make_float4(__fadd_rn(v.x, w.x), __fadd_rn(v.y, w.y), __fadd_rn(v.z, w.z), __fadd_rn(v.w, w.w));
    }
    out[r * ldo4 + c] = v;
  }
}

// nearest-neighbour upsample of channels-innermost rows (TUpsampleLayer.forwardGPU, nupsamplelayer.pas:190-206):
// out[b][h*s+i][w*s+j][c] = scale * in[b][h][w][c]
__global__ void __launch_bounds__(256) k_upsample_nhwc(int64_t total, int C4, int W, int H, int s, float scale,
                                                       const float4 *__restrict__ in, int64_t ldi4,
                                                       float4 *__restrict__ out, int64_t ldo4) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int OW = W * s, OH = H * s;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int64_t px = i / C4;
    const int c = (int)(i - px * C4);
    const int ow = (int)(px % OW);
    const int64_t t = px / OW;
    const int oh = (int)(t % OH);
    const int64_t b = t / OH;
    float4 v = __ldg(in + ((b * H + oh / s) * W + ow / s) * ldi4 + c);
    if (scale != 1.0f) v = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
    out[px * ldo4 + c] = v;
  }
}

// fp32 rows -> the pre-split planes (hi = bf16_rn(x), lo = bf16_rn(x - hi)) and back (hi + lo, exact in fp32);
// 8 channels per thread: two 16-byte loads, one 16-byte store per plane
__global__ void __launch_bounds__(256) k_rows_to_split(int64_t rows, int C8, const float4 *__restrict__ in, int64_t ldi4,
                                                       uint4 *__restrict__ hi, uint4 *__restrict__ lo, int64_t ldo8) {
  const int64_t total = rows * C8;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int64_t r = i / C8;
    const int c = (int)(i - r * C8);
    const float4 a = __ldg(in + r * ldi4 + 2 * c), b = __ldg(in + r * ldi4 + 2 * c + 1);
    uint4 h, l;
    split_bf16x8(a, b, h, l);
    hi[r * ldo8 + c] = h;
    lo[r * ldo8 + c] = l;
  }
}
__global__ void __launch_bounds__(256) k_split_to_rows(int64_t rows, int C8, const uint4 *__restrict__ hi,
                                                       const uint4 *__restrict__ lo, int64_t ldi8, float4 *__restrict__ out,
                                                       int64_t ldo4) {
  const int64_t total = rows * C8;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int64_t r = i / C8;
    const int c = (int)(i - r * C8);
    const uint4 h = __ldg(hi + r * ldi8 + c), l = __ldg(lo + r * ldi8 + c);
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = __fadd_rn(__uint_as_float(hh[j] << 16), __uint_as_float(ll[j] << 16));
      v[2 * j + 1] = __fadd_rn(__uint_as_float(hh[j] & 0xFFFF0000u), __uint_as_float(ll[j] & 0xFFFF0000u));
    }
    out[r * ldo4 + 2 * c] = make_float4(v[0], v[1], v[2], v[3]);
    out[r * ldo4 + 2 * c + 1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

static int grid_for(int64_t total) {
  int64_t blocks = cdiv64(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  return (int)(blocks < 1 ? 1 : blocks);
}

// ------------------------------------------------------------------ SGEMM on the convolution kernel
// C[M][N] = alpha * op(A) op(B) + beta * C is a 1x1 convolution: pixels = the M rows of op(A) (K "channels" each, K
// contiguous), filters = the N rows of op(B)^T, fused epilogue = *alpha (+ C).  Large products therefore run on the
// CTA-pair kernel above — TMA im2col rows, half a weight tile per CTA, TMA store — in all three tensor modes: the
// TMA-fed single-CTA SGEMM of umma_gemm.cuh tops out at 332 TFLOP/s in TF32 mode (shared-memory bandwidth, §4.3 of
// DESIGN.md), the pair kernel reaches the cuBLAS rate.  Operands that are not already K-contiguous are transposed into a
// scratch copy first (one pass over the operand, < 5 % of the product for the sizes taken); in BF16x3 mode the rows
// of op(A) are pre-split into the bf16 planes the kernel multiplies.
__global__ void __launch_bounds__(256) k_transpose(const float *__restrict__ src, int64_t lds, int64_t rows, int64_t cols,
                                                   float *__restrict__ dst, int64_t ldd) {   // dst[c][r] = src[r][c]
  __shared__ float t[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j, c = c0 + tx;
    t[j][tx] = (r < rows && c < cols) ? __ldg(src + r * lds + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int64_t c = c0 + j, r = r0 + tx;
    if (c < cols && r < rows) dst[c * ldd + r] = t[tx][j];
  }
}
__global__ void __launch_bounds__(256) k_fill1(float *p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

constexpr int kGemmNotTaken = umma::kNotTaken;

static int gemm_on_conv(tnn_context *ctx, bool tA, bool tB, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                        int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc) {
  if (!ctx->gemm_nhwc || ctx->math_mode == TNN_MATH_FP32) return kGemmNotTaken;
  if (!(beta == 0.0f || beta == 1.0f)) return kGemmNotTaken;
  const int sms = ctx->prop.multiProcessorCount;
  // enough 256 x 256 pair tiles for every TPC and a contraction worth the operand passes
  if (((M + 255) / 256) * ((N + 255) / 256) < sms / 2 || K < 256 || (K % 8) != 0 || M >= (1ll << 31) - 256 || N > 65535 * 16)
    return kGemmNotTaken;
  if ((ldc % 4) != 0 || (((uintptr_t)C) & 15) != 0 || (beta == 1.0f && (N % 4) != 0)) return kGemmNotTaken;
  const bool bf = ctx->math_mode == TNN_MATH_BF16X3;
  const bool aAligned = (lda % 4) == 0 && (((uintptr_t)A) & 15) == 0;
  const bool bDirect = tB && ldb == K;
  // rows that are read in place (fp32 rows by TMA, or by the 128-bit split / copy kernels) must be 16-byte aligned
  if (!tA && !aAligned) return kGemmNotTaken;
  if (tB && !bDirect && ((ldb % 4) != 0 || (((uintptr_t)B) & 15) != 0)) return kGemmNotTaken;
  // scratch: [A rows M x K fp32 (transposed A)] [A planes M x K x 4 bytes (bf16x3)] [B rows N x K] [alpha vector N]
  const size_t aRows = (tA ? (size_t)M * K : 0), aPlanes = bf ? (size_t)M * K : 0, bRows = bDirect ? 0 : (size_t)N * K;
  const size_t floats = aRows + aPlanes + bRows + (size_t)N + 64;
  if (ctx->gemm_scratch_bytes < floats * sizeof(float)) {
    if (ctx->capturing) return kGemmNotTaken;   // no allocation inside a capture: the eager warm-up step sized it, or not
    TNN_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->gemm_scratch) {
      int rc = tnn_release_scratch(ctx, ctx->gemm_scratch);
      if (rc) return rc;
    }
    ctx->gemm_scratch = nullptr;
    ctx->gemm_scratch_bytes = 0;
    TNN_CUDA(cudaMalloc((void **)&ctx->gemm_scratch, floats * sizeof(float)));
    ctx->gemm_scratch_bytes = floats * sizeof(float);
  }
  float *sA = ctx->gemm_scratch, *sP = sA + ((aRows + 3) & ~(size_t)3), *sB = sP + ((aPlanes + 3) & ~(size_t)3),
        *sAlpha = sB + ((bRows + 3) & ~(size_t)3);
  const float *rowsA = A;
  int64_t pitchA = lda;
  if (tA) {   // A is [K][M]: op(A) rows [M][K]
    dim3 grid((unsigned)cdiv64(M, 32), (unsigned)cdiv64(K, 32));
    k_transpose<<<grid, 256, 0, ctx->stream>>>(A, lda, K, M, sA, K);
    TNN_LAUNCHED(ctx);
    rowsA = sA;
    pitchA = K;
  }
  SplitRef in{nullptr, 0, 0};
  if (bf) {
    int rc = tnn_nhwc_to_split(ctx, M, K, rowsA, pitchA, sP, K, M * K);
    if (rc) return rc;
    in = SplitRef{reinterpret_cast<const uint16_t *>(sP), K, M * K};
  }
  const float *w = B;
  if (!bDirect) {
    if (tB) {   // [N][K] with padded rows: dense copy
      k_rows<false><<<grid_for(N * (K / 4)), 256, 0, ctx->stream>>>(N, (int)(K / 4), reinterpret_cast<const float4 *>(B),
                                                                 ldb / 4, nullptr, 0, reinterpret_cast<float4 *>(sB), K / 4);
    } else {    // B is [K][N]: op(B)^T rows [N][K]
      dim3 grid((unsigned)cdiv64(N, 32), (unsigned)cdiv64(K, 32));
      k_transpose<<<grid, 256, 0, ctx->stream>>>(B, ldb, K, N, sB, K);
    }
    TNN_LAUNCHED(ctx);
    w = sB;
  }
  const float *scale = nullptr;
  if (alpha != 1.0f) {
    k_fill1<<<(unsigned)cdiv64(N, 256), 256, 0, ctx->stream>>>(sAlpha, N, alpha);
    TNN_LAUNCHED(ctx);
    scale = sAlpha;
  }
  simt::ConvGeom g;
  g.C = (int)K; g.H = 1; g.W = (int)M; g.F = (int)N;
  g.kh = g.kw = 1; g.ph = g.pw = 0; g.sy = g.sx = 1; g.dy = g.dx = 1;
  g.OH = 1; g.OW = (int)M;
  g.inImg = K * M;
  g.outImg = N * M;
  ctx->last_conv_kernel = 7;
  return conv_forward(ctx, g, 1, bf ? nullptr : rowsA, pitchA, in, w, scale, nullptr, TNN_ACT_LINEAR, C, ldc, false,
                      SplitRef{nullptr, 0, 0}, beta == 1.0f ? C : nullptr, ldc);
}

}  // namespace nhwc

// (called by tnn_gemm's dispatch in contract.cu; umma::kNotTaken when the product stays on the SGEMM kernels)
int tnn_gemm_on_conv(tnn_context *ctx, bool tA, bool tB, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                     int64_t lda, const float *B, int64_t ldb, float beta, float *C, int64_t ldc) {
  return nhwc::gemm_on_conv(ctx, tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}

// ------------------------------------------------------------------ C ABI
namespace {
int fill_geom_nhwc(const tnn_conv_desc *d, simt::ConvGeom &g) {
  TNN_REQUIRE(d != nullptr, "null conv descriptor");
  TNN_REQUIRE(d->batch > 0 && d->c > 0 && d->h > 0 && d->w > 0 && d->filters > 0 && d->kh > 0 && d->kw > 0,
              "non-positive conv dimension");
  TNN_REQUIRE(d->stride_y > 0 && d->stride_x > 0 && d->dilation_y > 0 && d->dilation_x > 0 && d->pad_h >= 0 &&
                  d->pad_w >= 0, "bad stride/dilation/padding");
  const int64_t OH = (d->h + 2 * d->pad_h - (d->dilation_y * (d->kh - 1) + 1)) / d->stride_y + 1;
  const int64_t OW = (d->w + 2 * d->pad_w - (d->dilation_x * (d->kw - 1) + 1)) / d->stride_x + 1;
  TNN_REQUIRE(OH > 0 && OW > 0, "empty convolution output");
  TNN_REQUIRE(d->batch * OH * OW < (1ll << 31) && d->batch * d->h * d->w < (1ll << 31) &&
                  d->c * d->kh * d->kw < (1ll << 31), "convolution too large for 32-bit pixel indices");
  g.C = (int)d->c; g.H = (int)d->h; g.W = (int)d->w; g.F = (int)d->filters;
  g.kh = (int)d->kh; g.kw = (int)d->kw; g.ph = (int)d->pad_h; g.pw = (int)d->pad_w;
  g.sy = (int)d->stride_y; g.sx = (int)d->stride_x; g.dy = (int)d->dilation_y; g.dx = (int)d->dilation_x;
  g.OH = (int)OH; g.OW = (int)OW;
  g.inImg = d->c * d->h * d->w;
  g.outImg = d->filters * OH * OW;
  return 0;
}
}  // namespace

extern "C" {

int32_t tnn_conv2d_forward_nhwc(tnn_context *ctx, const tnn_conv_desc *d, const float *in, int64_t inOffset,
                                int64_t inPitch, int32_t inLayout, const float *weights, const float *scale,
                                const float *bias, int32_t activation, float *out, int64_t outOffset, int64_t outPitch,
                                int32_t outLayout) {
  return tnn_conv2d_forward_nhwc_residual(ctx, d, in, inOffset, inPitch, inLayout, weights, scale, bias, activation,
                                          nullptr, 0, 0, out, outOffset, outPitch, outLayout);
}

int32_t tnn_conv2d_forward_nhwc_residual(tnn_context *ctx, const tnn_conv_desc *d, const float *in, int64_t inOffset,
                                         int64_t inPitch, int32_t inLayout, const float *weights, const float *scale,
                                         const float *bias, int32_t activation, const float *residual,
                                         int64_t residualOffset, int64_t residualPitch, float *out, int64_t outOffset,
                                         int64_t outPitch, int32_t outLayout) {
  return tnn_conv2d_forward_split(ctx, d, in, inOffset, inPitch, inLayout, 0, weights, scale, bias, activation, residual,
                                  residualOffset, residualPitch, out, outOffset, outPitch, outLayout, nullptr, 0, 0, 0);
}

int32_t tnn_conv2d_forward_split(tnn_context *ctx, const tnn_conv_desc *d, const void *in, int64_t inOffset, int64_t inPitch,
                                 int32_t inLayout, int64_t inPlane, const float *weights, const float *scale,
                                 const float *bias, int32_t activation, const float *residual, int64_t residualOffset,
                                 int64_t residualPitch, float *out, int64_t outOffset, int64_t outPitch, int32_t outLayout,
                                 void *outSplit, int64_t outSplitOffset, int64_t outSplitPitch, int64_t outSplitPlane) {
  return tnn_conv2d_forward_split_residual(ctx, d, in, inOffset, inPitch, inLayout, inPlane, weights, scale, bias, activation,
                                           residual, residualOffset, residualPitch, TNN_LAYOUT_NHWC, 0, out, outOffset,
                                           outPitch, outLayout, outSplit, outSplitOffset, outSplitPitch, outSplitPlane);
}

int32_t tnn_conv2d_forward_split_residual(tnn_context *ctx, const tnn_conv_desc *d, const void *in, int64_t inOffset,
                                          int64_t inPitch, int32_t inLayout, int64_t inPlane, const float *weights,
                                          const float *scale, const float *bias, int32_t activation, const void *residualAny,
                                          int64_t residualOffset, int64_t residualPitch, int32_t residualLayout,
                                          int64_t residualPlane, float *out, int64_t outOffset, int64_t outPitch,
                                          int32_t outLayout, void *outSplit, int64_t outSplitOffset, int64_t outSplitPitch,
                                          int64_t outSplitPlane) {
  TNN_CHECK_CTX(ctx);
  TNN_REQUIRE(residualLayout == TNN_LAYOUT_NHWC || residualLayout == TNN_LAYOUT_NHWC_SPLIT,
              "the residual is rows of channels (fp32) or pre-split planes");
  const float *residual = residualLayout == TNN_LAYOUT_NHWC ? (const float *)residualAny : nullptr;
  nhwc::SplitRef rs{nullptr, 0, 0};
  if (residualAny && residualLayout == TNN_LAYOUT_NHWC_SPLIT) {
    rs = nhwc::SplitRef{(const uint16_t *)residualAny + residualOffset, residualPitch, residualPlane};
    TNN_REQUIRE(inLayout != TNN_LAYOUT_NCHW && outLayout == TNN_LAYOUT_NHWC,
                "the residual add is fused for channels-innermost tensors");
  }
  simt::ConvGeom g;
  int rc = fill_geom_nhwc(d, g);
  if (rc) return rc;
  TNN_REQUIRE(activation == TNN_ACT_LINEAR || tnn_act_is_elementwise(activation), "activation not elementwise");
  TNN_REQUIRE((inLayout == TNN_LAYOUT_NCHW || inLayout == TNN_LAYOUT_NHWC || inLayout == TNN_LAYOUT_NHWC_SPLIT) &&
                  (outLayout == TNN_LAYOUT_NCHW || outLayout == TNN_LAYOUT_NHWC), "unknown layout");
  TNN_REQUIRE(ctx->math_mode != TNN_MATH_FP32, "the channels-innermost path runs on the tensor cores (3xTF32 / TF32 mode)");
  TNN_REQUIRE(in != nullptr, "null input");
  if (out) out += outOffset;
  nhwc::SplitRef os{outSplit ? (const uint16_t *)outSplit + outSplitOffset : nullptr, outSplitPitch, outSplitPlane};
  if (residual) {
    residual += residualOffset;
    TNN_REQUIRE(inLayout != TNN_LAYOUT_NCHW && outLayout == TNN_LAYOUT_NHWC,
                "the residual add is fused for channels-innermost tensors");
  }
  if (inLayout == TNN_LAYOUT_NCHW) {
    // planar input (the image fed to the first layer): direct fp32 convolution writing channels-innermost rows
    const float *inf = (const float *)in + inOffset;
    TNN_REQUIRE(outLayout == TNN_LAYOUT_NHWC && direct::supported(g),
                "NCHW input is only taken by the small-contraction kernels (C*kh*kw <= 64) writing NHWC");
    if ((ctx->math_mode == TNN_MATH_BF16X3 || ctx->math_mode == TNN_MATH_TF32) && nhwc::image_supported(g, inf) && ctx->nhwc_image) {
      // tensor cores, BF16x3 arithmetic (in TF32 mode too: it is the tighter of the two)
      ctx->last_conv_kernel = 9;
      return nhwc::conv_image(ctx, g, (int)d->batch, inf, weights, scale, bias, activation, out, outPitch, os);
    }
    ctx->last_conv_kernel = 8;
    rc = direct::conv_forward(ctx, g, (int)d->batch, inf, weights, scale, bias, activation, out, outPitch);
    if (rc || !os.hi) return rc;
    return tnn_nhwc_to_split(ctx, d->batch * g.OH * g.OW, g.F, out, outPitch, (void *)os.hi, os.pitch, os.plane);
  }
  ctx->last_conv_kernel = 7;
  nhwc::SplitRef is{nullptr, 0, 0};
  const float *inf = nullptr;
  if (inLayout == TNN_LAYOUT_NHWC_SPLIT) is = nhwc::SplitRef{(const uint16_t *)in + inOffset, inPitch, inPlane};
  else inf = (const float *)in + inOffset;
  return nhwc::conv_forward(ctx, g, (int)d->batch, inf, inPitch, is, weights, scale, bias, activation, out, outPitch,
                            outLayout == TNN_LAYOUT_NCHW, os, residual, residualPitch, rs);
}

int32_t tnn_conv2d_forward_yolo_head(tnn_context *ctx, const tnn_conv_desc *d, const void *in, int64_t inOffset,
                                     int64_t inPitch, int32_t inLayout, int64_t inPlane, const float *weights,
                                     const float *scale, const float *bias, float *out, int64_t outOffset, int64_t classes,
                                     float scale_xy) {
  TNN_CHECK_CTX(ctx);
  simt::ConvGeom g;
  int rc = fill_geom_nhwc(d, g);
  if (rc) return rc;
  TNN_REQUIRE(inLayout == TNN_LAYOUT_NHWC || inLayout == TNN_LAYOUT_NHWC_SPLIT, "the head reads channels-innermost activations");
  TNN_REQUIRE(ctx->math_mode != TNN_MATH_FP32, "the channels-innermost path runs on the tensor cores");
  TNN_REQUIRE(in != nullptr && out != nullptr && classes > 0 && (g.F % (classes + 5)) == 0,
              "a yolo head has anchors * (classes + 5) filters");
  ctx->last_conv_kernel = 7;
  nhwc::SplitRef is{nullptr, 0, 0};
  const float *inf = nullptr;
  if (inLayout == TNN_LAYOUT_NHWC_SPLIT) is = nhwc::SplitRef{(const uint16_t *)in + inOffset, inPitch, inPlane};
  else inf = (const float *)in + inOffset;
  return nhwc::conv_forward(ctx, g, (int)d->batch, inf, inPitch, is, weights, scale, bias, TNN_ACT_LINEAR, out + outOffset, 0,
                            true, nhwc::SplitRef{nullptr, 0, 0}, nullptr, 0, nhwc::SplitRef{nullptr, 0, 0},
                            (int)(classes + 5), scale_xy);
}

int32_t tnn_nhwc_to_split(tnn_context *ctx, int64_t rows, int64_t channels, const float *in, int64_t inPitch, void *outSplit,
                          int64_t outPitch, int64_t outPlane) {
  TNN_CHECK_CTX(ctx);
  if (rows <= 0 || channels <= 0) return 0;
  TNN_REQUIRE((channels % 8) == 0 && (inPitch % 4) == 0 && (outPitch % 8) == 0 && (outPlane % 8) == 0 &&
                  ((((uintptr_t)in) | ((uintptr_t)outSplit)) & 15) == 0,
              "pre-split rows: channels, pitches and the plane stride multiples of 8, 16-byte aligned bases");
  uint16_t *hi = (uint16_t *)outSplit;
  nhwc::k_rows_to_split<<<nhwc::grid_for(rows * channels / 8), 256, 0, ctx->stream>>>(
      rows, (int)(channels / 8), reinterpret_cast<const float4 *>(in), inPitch / 4, reinterpret_cast<uint4 *>(hi),
      reinterpret_cast<uint4 *>(hi + outPlane), outPitch / 8);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_split_to_nhwc(tnn_context *ctx, int64_t rows, int64_t channels, const void *inSplit, int64_t inPitch,
                          int64_t inPlane, float *out, int64_t outPitch) {
  TNN_CHECK_CTX(ctx);
  if (rows <= 0 || channels <= 0) return 0;
  TNN_REQUIRE((channels % 8) == 0 && (inPitch % 8) == 0 && (inPlane % 8) == 0 && (outPitch % 4) == 0 &&
                  ((((uintptr_t)inSplit) | ((uintptr_t)out)) & 15) == 0,
              "pre-split rows: channels, pitches and the plane stride multiples of 8, 16-byte aligned bases");
  const uint16_t *hi = (const uint16_t *)inSplit;
  nhwc::k_split_to_rows<<<nhwc::grid_for(rows * channels / 8), 256, 0, ctx->stream>>>(
      rows, (int)(channels / 8), reinterpret_cast<const uint4 *>(hi), reinterpret_cast<const uint4 *>(hi + inPlane),
      inPitch / 8, reinterpret_cast<float4 *>(out), outPitch / 4);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_nchw_to_nhwc(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize, const float *in,
                         float *out, int64_t outPitch) {
  TNN_CHECK_CTX(ctx);
  if (batch <= 0 || channels <= 0 || blockSize <= 0) return 0;
  TNN_REQUIRE(outPitch >= channels && batch <= 65535 && cdiv64(channels, 32) <= 65535, "bad shape");
  dim3 grid((unsigned)cdiv64(blockSize, 32), (unsigned)cdiv64(channels, 32), (unsigned)batch);
  nhwc::k_nchw_to_nhwc<<<grid, 256, 0, ctx->stream>>>(in, out, (int)channels, (int)blockSize, outPitch);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_nhwc_to_nchw(tnn_context *ctx, int64_t batch, int64_t channels, int64_t blockSize, const float *in,
                         int64_t inPitch, float *out) {
  TNN_CHECK_CTX(ctx);
  if (batch <= 0 || channels <= 0 || blockSize <= 0) return 0;
  TNN_REQUIRE(inPitch >= channels && batch <= 65535 && cdiv64(channels, 32) <= 65535, "bad shape");
  dim3 grid((unsigned)cdiv64(blockSize, 32), (unsigned)cdiv64(channels, 32), (unsigned)batch);
  nhwc::k_nhwc_to_nchw<<<grid, 256, 0, ctx->stream>>>(in, out, (int)channels, (int)blockSize, inPitch);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_rows_copy(tnn_context *ctx, int64_t rows, int64_t channels, const float *in, int64_t inPitch, float *out,
                      int64_t outPitch) {
  TNN_CHECK_CTX(ctx);
  if (rows <= 0 || channels <= 0) return 0;
  TNN_REQUIRE(((channels | inPitch | outPitch) % 4) == 0 && ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0,
              "rows must be 16-byte aligned multiples of 4 floats");
  nhwc::k_rows<false><<<nhwc::grid_for(rows * channels / 4), 256, 0, ctx->stream>>>(
      rows, (int)(channels / 4), reinterpret_cast<const float4 *>(in), inPitch / 4, nullptr, 0,
      reinterpret_cast<float4 *>(out), outPitch / 4);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_rows_add(tnn_context *ctx, int64_t rows, int64_t channels, const float *a, int64_t aPitch, const float *b,
                     int64_t bPitch, float *out, int64_t outPitch) {
  TNN_CHECK_CTX(ctx);
  if (rows <= 0 || channels <= 0) return 0;
  TNN_REQUIRE(((channels | aPitch | bPitch | outPitch) % 4) == 0 &&
                  ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)out)) & 15) == 0,
              "rows must be 16-byte aligned multiples of 4 floats");
  nhwc::k_rows<true><<<nhwc::grid_for(rows * channels / 4), 256, 0, ctx->stream>>>(
      rows, (int)(channels / 4), reinterpret_cast<const float4 *>(a), aPitch / 4, reinterpret_cast<const float4 *>(b),
      bPitch / 4, reinterpret_cast<float4 *>(out), outPitch / 4);
  TNN_LAUNCHED(ctx);
  return 0;
}

int32_t tnn_upsample_nhwc(tnn_context *ctx, int64_t batch, int64_t channels, int64_t inHeight, int64_t inWidth,
                          int64_t stride, float scale, const float *in, int64_t inPitch, float *out, int64_t outPitch) {
  TNN_CHECK_CTX(ctx);
  if (batch <= 0 || channels <= 0) return 0;
  TNN_REQUIRE(stride >= 1 && ((channels | inPitch | outPitch) % 4) == 0 &&
                  ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0, "rows must be 16-byte aligned multiples of 4 floats");
  const int64_t total = batch * inHeight * stride * inWidth * stride * (channels / 4);
  nhwc::k_upsample_nhwc<<<nhwc::grid_for(total), 256, 0, ctx->stream>>>(
      total, (int)(channels / 4), (int)inWidth, (int)inHeight, (int)stride, scale, reinterpret_cast<const float4 *>(in),
      inPitch / 4, reinterpret_cast<float4 *>(out), outPitch / 4);
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // extern "C"
