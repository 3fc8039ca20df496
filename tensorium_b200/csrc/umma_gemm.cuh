// SGEMM with both operands fed by TMA (replaces the cublasSgemm behind TNNCuda.gemm, source/nncuda.pas:857-880,
// for 16-byte aligned operands; anything else stays on the register-staged loaders of umma_conv.cuh).
//
// One elected thread of the TMA warp lands the raw fp32 tiles of a k-block straight in the layout the tensor core
// reads: a k-contiguous operand as SWIZZLE_128B rows of 32 k, an m/n-contiguous operand as 4 KB atoms of
// 32 k x 32 m/n (SWIZZLE_128B_ATOM_32B, the layout kind::tf32 needs for MN-major).  The eight producer warps
// never touch global memory: they wait for the tiles and write the 3xTF32 low halves (lo = x - trunc(x); the raw
// tile is the high half because kind::tf32 reads only the top 19 bits) or round the tile in place (TF32 mode).
// OOB rows and the k tail are zero-filled by the TMA unit, so M, N, K are arbitrary.
#pragma once
#include "umma_conv.cuh"

namespace umma {

struct GemmVecEpilogue {  // C[m][n] = alpha*acc + beta*C, float4 stores on full aligned chunks
  float *C;
  int64_t ldc;
  float alpha, beta;
  int64_t rows, cols;
  int vec;  // C and ldc are 16-byte aligned
  struct Row { float *q; };
  __device__ __forceinline__ Row row(int64_t m, bool &valid) const {
    valid = m < rows;
    return Row{C + (valid ? m : 0) * ldc};
  }
  __device__ __forceinline__ void store_chunk(const Row &r, int64_t n, const uint32_t (&acc)[32], int cnt,
                                              bool rowvalid, float *) const {
    if (!rowvalid) return;
    float *q = r.q + n;
    if (vec && cnt == 32) {
      float4 *q4 = reinterpret_cast<float4 *>(q);
      if (beta == 0.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          q4[j] = make_float4(alpha * __uint_as_float(acc[4 * j]), alpha * __uint_as_float(acc[4 * j + 1]),
                              alpha * __uint_as_float(acc[4 * j + 2]), alpha * __uint_as_float(acc[4 * j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 c = q4[j];
          q4[j] = make_float4(fmaf(alpha, __uint_as_float(acc[4 * j]), beta * c.x),
                              fmaf(alpha, __uint_as_float(acc[4 * j + 1]), beta * c.y),
                              fmaf(alpha, __uint_as_float(acc[4 * j + 2]), beta * c.z),
                              fmaf(alpha, __uint_as_float(acc[4 * j + 3]), beta * c.w));
        }
      }
      return;
    }
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

// AMN / BMN: the operand is contiguous along m / n (transA / !transB); otherwise along k.
// KB = k extent of one pipeline stage: 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B; twice the
// stages in the same shared memory, which is what hides the landing + low-half latency in 3xTF32 mode).
template <bool SPLIT, bool AMN, bool BMN, int KB, typename EPI>
__global__ void __launch_bounds__(kThreads, 1)
k_umma_gemm(EPI epi, TileSched sched, int BN, int stages, uint32_t accCols, const __grid_constant__ CUtensorMap tmA,
            const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ SharedCtl ctl;
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the shared address space
  // and emits LDS / STS instead of generic LD / ST for every access below)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int aBytes = BM * KB * 4;
  const int bBytes = BN * KB * 4;
  const int stageBytes = (aBytes + bBytes) * (SPLIT ? 2 : 1);
  const int bOff = (SPLIT ? 2 : 1) * aBytes;  // stage layout: A | A_lo | B | B_lo   (TF32: A | B)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&ctl.full[s], kProducerWarps);
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
    tma_prefetch_desc(&tmB);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl.tmem_base;

  if (warp >= kFirstProducerWarp && warp < kMmaWarp) {
    // ===================== producers: raw tiles -> low halves (or rounded in place) =====================
    const int tid = threadIdx.x - kFirstProducerWarp * 32;
    int64_t remaining = 0;
    for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x) {
      int mt, nt, k0, k1;
      sched.decode(t, mt, nt, k0, k1);
      remaining += k1 - k0;
    }
    int stage = 0;
    uint32_t phase = 0;
    const int nB = bBytes / 16;
    for (; remaining > 0; --remaining) {
      mbar_wait_warp(&ctl.araw[stage], phase);
      uint8_t *sa = smem + (size_t)stage * stageBytes;
      uint8_t *sb = sa + bOff;
#pragma unroll
      for (int i = 0; i < aBytes / (16 * kProducerThreads); ++i) {
        const int off = (i * kProducerThreads + tid) * 16;
        const float4 v = *reinterpret_cast<const float4 *>(sa + off);
        if (SPLIT) {
          *reinterpret_cast<float4 *>(sa + aBytes + off) =
              make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
        } else {
          *reinterpret_cast<float4 *>(sa + off) = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        }
      }
#pragma unroll 4
      for (int i = tid; i < nB; i += kProducerThreads) {
        const int off = i * 16;
        const float4 v = *reinterpret_cast<const float4 *>(sb + off);
        if (SPLIT) {
          *reinterpret_cast<float4 *>(sb + bBytes + off) =
              make_float4(v.x - tf32_hi(v.x), v.y - tf32_hi(v.y), v.z - tf32_hi(v.z), v.w - tf32_hi(v.w));
        } else {
          *reinterpret_cast<float4 *>(sb + off) = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.full[stage]);
      if (++stage == stages) { stage = 0; phase ^= 1; }
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
      const uint32_t idesc = make_idesc(sched.half(t) ? BN >> 1 : BN, AMN, BMN);
      const int as = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait_warp(&ctl.tmem_empty[as], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_warp(&ctl.full[stage], phase);
        tc_fence_after();
        if (TNN_MMA_ELECT) {
          const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
          const uint32_t sb = sa + bOff;
          // k-contiguous: SWIZZLE_128B rows, 8-row groups 1024 B apart, one UMMA_K = 8 step = +32 B.
          // m/n-contiguous: 4 KB atoms (32 k x 32 m/n) side by side; inside an atom 8-k groups are 512 B... the
          // descriptor walks LBO = 4096 between atoms, SBO = 512 between k groups, one UMMA_K step = +1024 B.
          constexpr uint32_t kAtom = KB * 128, kSboK = KB == 32 ? 1024 : 512, kLtK = KB == 32 ? 2u : 4u;
          constexpr uint32_t a_lbo = AMN ? kAtom : 16, a_sbo = AMN ? 512 : kSboK, a_lt = AMN ? 1u : kLtK;
          constexpr uint32_t b_lbo = BMN ? kAtom : 16, b_sbo = BMN ? 512 : kSboK, b_lt = BMN ? 1u : kLtK;
#pragma unroll
          for (int ks = 0; ks < KB / 8; ++ks) {
            const uint32_t aoff = AMN ? ks * 1024 : ks * 32;
            const uint32_t boff = BMN ? ks * 1024 : ks * 32;
            const uint64_t ah = make_desc(sa + aoff, a_lbo, a_sbo, a_lt);
            const uint64_t bh = make_desc(sb + boff, b_lbo, b_sbo, b_lt);
            const uint32_t acc = (kb != kb0 || ks != 0) ? 1u : 0u;
            if (SPLIT) {
              const uint64_t al = make_desc(sa + aBytes + aoff, a_lbo, a_sbo, a_lt);
              const uint64_t bl = make_desc(sb + bBytes + boff, b_lbo, b_sbo, b_lt);
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
        const int m0 = mt * BM, n0 = nt * BN + (sched.half(t) == 2 ? BN >> 1 : 0);  // half tile: full box from its first row
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&ctl.empty[stage], phase ^ 1);
          const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
          const uint32_t sb = sa + bOff;
          mbar_arrive_expect_tx(&ctl.araw[stage], (uint32_t)(aBytes + bBytes));
          if (AMN) {
#pragma unroll
            for (int a = 0; a < BM / 32; ++a) tma_load_2d(sa + a * (KB * 128), &tmA, m0 + a * 32, kb * KB, &ctl.araw[stage]);
          } else {
            tma_load_2d(sa, &tmA, kb * KB, m0, &ctl.araw[stage]);
          }
          if (BMN) {
            for (int a = 0; a < (BN >> 5); ++a) tma_load_2d(sb + a * (KB * 128), &tmB, n0 + a * 32, kb * KB, &ctl.araw[stage]);
          } else {
            tma_load_2d(sb, &tmB, kb * KB, n0, &ctl.araw[stage]);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps 0-3: TMEM -> registers -> global =====================
    const int quad = warp;
    int it = 0;
    for (int t = blockIdx.x; t < sched.numTiles; t += gridDim.x, ++it) {
      int mt, nt, kb0, kb1;
      sched.decode(t, mt, nt, kb0, kb1);
      const int as = (int)(it & 1);
      const uint32_t use = (uint32_t)(it >> 1);
      const int64_t m = (int64_t)mt * BM + quad * 32 + lane;
      const int hf = sched.half(t);
      const int bn = hf ? BN >> 1 : BN;
      const int64_t n0 = (int64_t)nt * BN + (hf == 2 ? BN >> 1 : 0);
      bool mvalid;
      typename EPI::Row rc = epi.row(m, mvalid);
      mbar_wait_warp_relaxed(&ctl.tmem_full[as], use & 1, 64);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols + ((uint32_t)(quad * 32) << 16);
      const int nChunks = (bn + 31) >> 5;
      for (int ci = 0; ci < nChunks; ++ci) {
        const int c0 = ci * 32;
        uint32_t r[32];
        tmem_ld32(tmem_d + (uint32_t)c0, r);
        if (ci == nChunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl.tmem_empty[as]);
        }
        int64_t cnt = min((int64_t)(bn - c0), epi.cols - (n0 + c0));
        if (cnt > 32) cnt = 32;
        if (cnt > 0) epi.store_chunk(rc, n0 + c0, r, (int)cnt, mvalid, nullptr);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, accCols * 2);
}

// rows x K operand, k-contiguous (ld floats between rows): boxes of 32 k x boxRows rows, SWIZZLE_128B.
// m/n-contiguous (ld floats between k): boxes of 32 m/n x 32 k, SWIZZLE_128B_ATOM_32B.
inline int encode_gemm_operand(CUtensorMap *tm, const float *base, bool mnContig, int64_t rows, int64_t K, int64_t ld,
                               int boxRows, int kb) {
  EncodeTiledFn enc = get_encode_tiled();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)(mnContig ? rows : K), (cuuint64_t)(mnContig ? K : rows)};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)(mnContig ? 32 : kb), (cuuint32_t)(mnContig ? kb : boxRows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mnContig ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                            : (kb == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeTiled (gemm operand) failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

template <bool SPLIT, bool AMN, bool BMN, int KB, typename EPI>
int launch_gemm_tma(tnn_context *ctx, const Plan &p, EPI epi, const CUtensorMap &tmA, const CUtensorMap &tmB) {
  auto kern = k_umma_gemm<SPLIT, AMN, BMN, KB, EPI>;
  TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smemBytes));
  kern<<<p.grid, kThreads, p.smemBytes, ctx->stream>>>(epi, p.sched, p.BN, p.stages, p.accCols, tmA, tmB);
  TNN_LAUNCHED(ctx);
  return 0;
}

// make_plan with a stage of KB k (the tile scheduler counts k-blocks of KB)
inline Plan make_gemm_plan(tnn_context *ctx, int64_t M, int64_t N, int64_t K, bool split, bool bMnMajor, bool allowSplitK,
                           int kb) {
  Plan p = make_plan(ctx, M, N, K, split, bMnMajor, false);
  const size_t stageBytes = (size_t)(BM + p.BN) * kb * 4 * (split ? 2 : 1);
  int st = (int)((200 * 1024) / stageBytes);
  p.stages = st > kMaxStages ? kMaxStages : st < 2 ? 2 : st;
  p.smemBytes = stageBytes * p.stages + 1024;
  TileSched &s = p.sched;
  s.kBlocks = (int)cdiv64(K, kb);
  if (s.kBlocks < 1) s.kBlocks = 1;
  const int64_t sms = ctx->prop.multiProcessorCount;
  const int64_t tiles = (int64_t)s.mTiles * s.nTiles;
  int64_t ks = 1;
  const int perSplitMin = 8 * (BK / kb);  // same k extent per split as the register-staged plan
  if (allowSplitK && tiles < sms && s.kBlocks >= 2 * perSplitMin) {
    ks = sms / tiles;
    const int64_t maxs = s.kBlocks / perSplitMin;
    if (ks > maxs) ks = maxs;
    if (ks < 1) ks = 1;
  }
  s.kBlocksPerSplit = (int)cdiv64(s.kBlocks, ks);
  s.kSplits = (int)cdiv64(s.kBlocks, s.kBlocksPerSplit);
  s.numTiles = s.mTiles * s.nTiles * s.kSplits;
  s.wideTiles = s.numTiles;
  split_tail_wave(ctx, p);
  p.grid = (int)(s.numTiles < sms ? s.numTiles : sms);
  return p;
}

template <bool AMN, bool BMN, int KB>
int gemm_tma_one(tnn_context *ctx, int64_t M, int64_t N, int64_t K, float alpha, const float *A, int64_t lda,
                 const float *B, int64_t ldb, float beta, float *C, int64_t ldc) {
  const bool split = tnn_math_is_split(ctx);
  const bool zeroFirst = beta == 0.0f && ldc == N;
  Plan p = make_gemm_plan(ctx, M, N, K, split, BMN, beta == 1.0f || zeroFirst, KB);
  alignas(64) CUtensorMap tmA, tmB;
  int rc = encode_gemm_operand(&tmA, A, AMN, M, K, lda, BM, KB);
  if (rc) return rc;
  rc = encode_gemm_operand(&tmB, B, BMN, N, K, ldb, p.BN, KB);
  if (rc) return rc;
  if (p.sched.kSplits > 1) {
    if (beta == 0.0f) TNN_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), ctx->stream));
    GemmAtomicEpilogue epi{C, ldc, alpha, beta, M, N};
    return split ? launch_gemm_tma<true, AMN, BMN, KB>(ctx, p, epi, tmA, tmB)
                 : launch_gemm_tma<false, AMN, BMN, KB>(ctx, p, epi, tmA, tmB);
  }
  GemmVecEpilogue epi{C, ldc, alpha, beta, M, N, ((((uintptr_t)C) & 15) == 0 && (ldc % 4) == 0) ? 1 : 0};
  return split ? launch_gemm_tma<true, AMN, BMN, KB>(ctx, p, epi, tmA, tmB)
               : launch_gemm_tma<false, AMN, BMN, KB>(ctx, p, epi, tmA, tmB);
}

// kNotTaken unless every operand meets the TMA rules (16-byte aligned base and row pitch, 32-bit coordinates).
inline int gemm_tma(tnn_context *ctx, bool tA, bool tB, int64_t M, int64_t N, int64_t K, float alpha, const float *A,
                    int64_t lda, int64_t sA, const float *B, int64_t ldb, int64_t sB, float beta, float *C, int64_t ldc,
                    int64_t sC, int batch) {
  auto ok = [](const float *p, int64_t ld, int64_t zs) {
    return (((uintptr_t)p) & 15) == 0 && (ld % 4) == 0 && (zs % 4) == 0 && ld < (1ll << 31);
  };
  if (!ok(A, lda, sA) || !ok(B, ldb, sB) || K < BK || M >= (1ll << 31) || N >= (1ll << 31) || K >= (1ll << 31))
    return kNotTaken;
  // 3xTF32 stages are twice as large: half-depth stages keep four of them in flight (ctx->tma_gemm == 2 forces 32)
  const bool kb16 = ctx->tma_gemm == 3 || (ctx->tma_gemm == 1 && tnn_math_is_split(ctx));
  for (int z = 0; z < batch; ++z) {
    const float *Az = A + z * sA, *Bz = B + z * sB;
    float *Cz = C + z * sC;
    int rc;
#define TNN_GEMM_CASE(AMN, BMN)                                                                             \
  (kb16 ? gemm_tma_one<AMN, BMN, 16>(ctx, M, N, K, alpha, Az, lda, Bz, ldb, beta, Cz, ldc)                   \
        : gemm_tma_one<AMN, BMN, 32>(ctx, M, N, K, alpha, Az, lda, Bz, ldb, beta, Cz, ldc))
    if (!tA && tB) rc = TNN_GEMM_CASE(false, false);
    else if (!tA && !tB) rc = TNN_GEMM_CASE(false, true);
    else if (tA && !tB) rc = TNN_GEMM_CASE(true, true);
    else rc = TNN_GEMM_CASE(true, false);
#undef TNN_GEMM_CASE
    if (rc) return rc;
  }
  return 0;
}

}  // namespace umma
