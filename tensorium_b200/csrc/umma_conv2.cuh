// umma_conv2.cuh — the gathering implicit-GEMM convolution on CTA PAIRS (tcgen05.mma.cta_group::2).
//
// The single-CTA kernel is bound by shared-memory bandwidth on the wide layers: per 32-deep k-block the tensor core
// reads 3 x (16 KB A + 32 KB B), the producers write 32 KB and the TMA writes 64 KB — 240 KB against ~205 KB the SM
// can move in the 1600 cycles the twelve MMAs take.  A CTA pair (two SMs of one TPC, cluster 2x1x1) computes a
// 256 x 256 tile: each CTA gathers the A rows of ITS 128 pixels and loads only HALF of the weight tile (128 of the
// 256 filters); the pair's tensor cores share the two B halves.  Per CTA and k-block: 3 x (16 + 16) KB read,
// 32 + 32 KB written — and a stage is 64 KB instead of 96, so the ring is 3 deep instead of 2.
//
// Protocol (r = rank in the pair, leader = rank 0):
//   full[s]   leader's barrier: 8 local + 8 remote producer-warp arrivals + the leader's expect_tx for BOTH weight
//             halves (the peer's TMA signals the leader's barrier: cp.async.bulk.tensor ... cta_group::2)
//   empty[s]  one barrier per CTA, released together by the leader's multicast tcgen05.commit
//   tmem_full[a] per CTA (multicast commit);  tmem_empty[a] leader's: 4 local + 4 remote epilogue-warp arrivals
//   The leader's elected thread issues every MMA; each CTA's epilogue drains its own 128 TMEM lanes.
#pragma once
#include "umma_conv.cuh"

#ifndef TNN_PAIR_ELECT
#define TNN_PAIR_ELECT (lane == 0)
#endif
namespace umma {

constexpr int kPairStages = 3;
constexpr int kPairBN = 256;       // filters per pair tile (128 per CTA)

struct PairCtl {
  uint64_t full[kPairStages], empty[kPairStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  int lut[64];
  alignas(16) float epi_sb[kEpilogueWarps][64];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t rank) {
  asm volatile(
      "{\n.reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}\n" ::"r"(smem_u32(bar)),
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
__device__ __forceinline__ void mma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  const uint32_t z = 0;
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
// weight half -> local smem, completion bytes on the LEADER's barrier (peer bit of the barrier address cleared)
__device__ __forceinline__ void tma2_load_2d(uint32_t dst_smem, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], "
      "[%4];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar) & 0xFEFFFFFFu)
      : "memory");
}

template <bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
k_umma_pair(Im2colPixelLoader la, ConvEpilogueT<false> epi, int mTiles, int nTiles, int kBlocks,
            const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ PairCtl ctl;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int aBytes = BM * BK * 4;               // this CTA's 128 pixel rows
  constexpr int bBytes = (kPairBN / 2) * BK * 4;    // this CTA's half of the weight tile
  constexpr int stageBytes = (aBytes + bBytes) * (SPLIT ? 2 : 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, numPairs = gridDim.x >> 1;
  const int mPairs = (mTiles + 1) >> 1;
  const int numTiles = mPairs * nTiles;             // pair tiles: 256 pixels x 256 filters

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&ctl.full[s], 2 * kProducerWarps + 1);
      mbar_init(&ctl.empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl.tmem_full[s], 1);
      mbar_init(&ctl.tmem_empty[s], 2 * kEpilogueWarps);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc2(&ctl.tmem_base, 512);
  la.fill_lut(ctl.lut, threadIdx.x);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmBhi);
    if (SPLIT) tma_prefetch_desc(&tmBlo);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = ctl.tmem_base;

  if (warp >= kFirstProducerWarp && warp < kMmaWarp) {
    // ===================== producers: gather this CTA's 128 pixel rows (as in k_umma) =====================
    constexpr int D = Im2colPixelLoader::kPrefetch;
    const int tid = threadIdx.x - kFirstProducerWarp * 32;
    Im2colPixelLoader::Regs ra[D];
    int ltile = pair, lkb = 0;
    Im2colPixelLoader::TileCtx ca;
    bool lvalid = ltile < numTiles;
    if (lvalid) ca = la.tile_ctx(BM, (int64_t)((ltile % mPairs) * 2 + (int)rank) * BM, tid, 0);
    auto issue_load = [&](Im2colPixelLoader::Regs &xa) {
      if (!lvalid) return;
      la.load(ca, xa, lkb, tid, ctl.lut);
      if (++lkb == kBlocks) {
        lkb = 0;
        ltile += numPairs;
        lvalid = ltile < numTiles;
        if (lvalid) ca = la.tile_ctx(BM, (int64_t)((ltile % mPairs) * 2 + (int)rank) * BM, tid, 0);
      }
    };
    int64_t remaining = 0;
    for (int t = pair; t < numTiles; t += numPairs) remaining += kBlocks;
    int stage = 0;
    uint32_t phase = 0;
    int prime = D;
    while (remaining > 0) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (remaining > 0) {
          if (prime == 0) {
            mbar_wait_warp(&ctl.empty[stage], phase ^ 1);
            uint8_t *sa = smem + (size_t)stage * stageBytes;
            la.template store<SPLIT>(ra[d], sa, sa + aBytes, BM, tid);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) mbar_arrive(&ctl.full[stage]);
              else mbar_arrive_cluster(&ctl.full[stage], 0);
            }
            if (++stage == kPairStages) { stage = 0; phase ^= 1; }
            --remaining;
          } else {
            --prime;
          }
          issue_load(ra[d]);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer: the leader's elected thread drives both tensor cores =====================
    if (rank == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(kPairBN >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);  // D=F32, A=B=TF32, A MN-major, N=256, M=256
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = pair; t < numTiles; t += numPairs, ++it) {
        const int as = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait_warp(&ctl.tmem_empty[as], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * kPairBN;
        for (int kb = 0; kb < kBlocks; ++kb) {
          mbar_wait_warp(&ctl.full[stage], phase);
          tc_fence_after();
          if (TNN_PAIR_ELECT) {
            const uint32_t sa = smem_base + (uint32_t)stage * stageBytes;
            const uint32_t sa_lo = sa + aBytes;
            const uint32_t sb = sa + (SPLIT ? 2 : 1) * aBytes;
            const uint32_t sb_lo = sb + bBytes;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
              const uint32_t aoff = ks * 2 * (BM / 32) * 512, boff = ks * 32;
              const uint64_t ah = make_desc(sa + aoff, 512, (BM / 32) * 512, 1);
              const uint64_t bh = make_desc(sb + boff, 16, 1024, 2);
              const uint32_t acc = (kb != 0 || ks != 0) ? 1u : 0u;
              if (SPLIT) {
                const uint64_t al = make_desc(sa_lo + aoff, 512, (BM / 32) * 512, 1);
                const uint64_t bl = make_desc(sb_lo + boff, 16, 1024, 2);
                mma2_tf32(tmem_d, al, bh, idesc, acc);
                mma2_tf32(tmem_d, ah, bl, idesc, 1u);
                mma2_tf32(tmem_d, ah, bh, idesc, 1u);
              } else {
                mma2_tf32(tmem_d, ah, bh, idesc, acc);
              }
            }
            mma2_commit(&ctl.empty[stage]);
            if (kb == kBlocks - 1) mma2_commit(&ctl.tmem_full[as]);
          }
          __syncwarp();
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kTmaWarp) {
    // ===================== TMA: this CTA's half of the weight tile; bytes are counted on the leader's barrier ======
    if (lane == 0) {
      const uint32_t smem_base = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < numTiles; t += numPairs) {
        const int nt = t / mPairs;
        for (int kb = 0; kb < kBlocks; ++kb) {
          mbar_wait(&ctl.empty[stage], phase ^ 1);
          const uint32_t sb = smem_base + (uint32_t)stage * stageBytes + (SPLIT ? 2 : 1) * aBytes;
          if (rank == 0) mbar_arrive_expect_tx(&ctl.full[stage], (uint32_t)bBytes * (SPLIT ? 2 : 1) * 2);
          const int row = nt * kPairBN + (int)rank * (kPairBN / 2);
          tma2_load_2d(sb, &tmBhi, kb * BK, row, &ctl.full[stage]);
          if (SPLIT) tma2_load_2d(sb + bBytes, &tmBlo, kb * BK, row, &ctl.full[stage]);
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue: each CTA drains its own 128 TMEM lanes =====================
    const int quad = warp;
    int it = 0;
    for (int t = pair; t < numTiles; t += numPairs, ++it) {
      const int mt = (t % mPairs) * 2 + (int)rank, nt = t / mPairs;
      const int as = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int64_t m = (int64_t)mt * BM + quad * 32 + lane;
      const int64_t n0 = (int64_t)nt * kPairBN;
      bool mvalid;
      ConvEpilogueT<false>::Row rc = epi.row(m, mvalid);
      mbar_wait_warp_relaxed(&ctl.tmem_full[as], use & 1, 64);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * kPairBN + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
      for (int ci = 0; ci < kPairBN / 32; ++ci) {
        const int c0 = ci * 32;
        uint32_t r[32];
        tmem_ld32(tmem_d + (uint32_t)c0, r);
        if (ci == kPairBN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (rank == 0) mbar_arrive(&ctl.tmem_empty[as]);
            else mbar_arrive_cluster(&ctl.tmem_empty[as], 0);
          }
        }
        int64_t cnt = min((int64_t)(kPairBN - c0), epi.cols - (n0 + c0));
        if (cnt > 32) cnt = 32;
        if (cnt > 0) epi.store_chunk(rc, n0 + c0, r, (int)cnt, mvalid, ctl.epi_sb[quad]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == kMmaWarp) tmem_dealloc2(tmem_base, 512);
}

// plain forward convolutions with >= 256 filters on the gathering path; kNotTaken otherwise
inline int conv_forward_pair(tnn_context *ctx, const simt::ConvGeom &g, int batch, const float *in, const float *weights,
                             const float *scale, const float *bias, int act, float *out) {
  const int64_t M = (int64_t)batch * g.OH * g.OW;
  const int64_t N = g.F;
  const int kk = g.kh * g.kw;
  if (N < kPairBN || kk > 64 || g.inImg >= (1ll << 31)) return kNotTaken;
  const int64_t K = (int64_t)((g.C + 15) / 16) * kk * 16;
  const bool split = ctx->math_mode == TNN_MATH_3XTF32;
  const int kBlocks = (int)cdiv64(K, BK);
  const int64_t Kpad = (int64_t)kBlocks * BK;
  const int mTiles = (int)cdiv64(M, BM), nTiles = (int)cdiv64(N, kPairBN);
  if (mTiles < 2 * 74) return kNotTaken;  // too few tiles to fill 74 pairs
  float *whi, *wlo;
  bool fresh;
  int rc = tnn_weight_buffer(ctx, weights, 3, (size_t)N * Kpad, &whi, &wlo, &fresh);
  if (rc) return rc;
  if (fresh) {
    int64_t blocks = cdiv64(N * Kpad, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_prepare_weights<<<(int)blocks, 256, 0, ctx->stream>>>(weights, N, g.C, kk, Kpad, whi, wlo, split ? 1 : 0);
    TNN_LAUNCHED(ctx);
  }
  alignas(64) CUtensorMap tmHi, tmLo;
  rc = encode_weights_map(&tmHi, whi, N, Kpad, kPairBN / 2);
  if (rc) return rc;
  rc = encode_weights_map(&tmLo, split ? wlo : whi, N, Kpad, kPairBN / 2);
  if (rc) return rc;
  Im2colPixelLoader la{in, g, M, g.H * g.W, kk};
  ConvEpilogueT<false> epi{out, scale, bias, g.OH * g.OW, g.outImg, act, M, N, false, 0, 0, 0, 0, 0, 0, 0, 0, nullptr};
  const size_t stageBytes = (size_t)(BM * BK * 4 + (kPairBN / 2) * BK * 4) * (split ? 2 : 1);
  const size_t smemBytes = stageBytes * kPairStages + 1024;
  const int sms = ctx->prop.multiProcessorCount & ~1;
  const int64_t pairTiles = (int64_t)((mTiles + 1) / 2) * nTiles;
  const int grid = (int)(2 * pairTiles < sms ? 2 * pairTiles : sms);
  auto kern = split ? k_umma_pair<true> : k_umma_pair<false>;
  TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  kern<<<grid, kThreads, smemBytes, ctx->stream>>>(la, epi, mTiles, nTiles, kBlocks, tmHi, tmLo);
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // namespace umma
