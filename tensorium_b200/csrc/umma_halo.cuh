// umma_halo.cuh — stride-1 KxK convolution on tcgen05 with the input tile staged ONCE per channel block.
//
// The general kernel (umma_conv.cuh) re-gathers the activations for every kernel tap: kh*kw im2col tiles per
// channel block, all written by the producer warps.  Here a tile is a patch of 16 x 8 output pixels of one
// image, and the producers stage its HALO — HR x HC = (16+kh-1) x (8+kw-1) input pixels x 32 channels — a single
// time, as a K-major operand in the un-swizzled ("interleaved") canonical layout:
//     byte(pixel R = line*HC + column, channel ci) = (ci/4)*plane + R*16 + (ci%4)*4,     plane = HR*HC*16
// i.e. 8 planes of 4 channels, one 16-byte chunk per pixel.  A core matrix (8 rows x 16 B) is 8 consecutive pixels
// = 128 contiguous bytes, so every kernel tap (kr,kc) is the SAME data seen through a descriptor shifted by whole
// pixels:   start = halo + ((kr*dy)*HC + kc*dx)*16,  SBO = HC*16 (next patch line),  LBO = plane (next 4 channels).
// The gather work and the L2->SM traffic drop by ~kh*kw and the MMA warp issues kh*kw k-blocks per staged halo.
// (A swizzled K-major halo does not work: measured on B200, the hardware derives the 128B-swizzle phase from the
// row index relative to the descriptor start, so a view shifted by one pixel reads permuted 16-byte chunks.)
//
// Weights arrive by TMA (K-major SWIZZLE_128B, pre-split) in the order k = ((cb*kk + tap)*32 + ci).
#pragma once
#include "umma_conv.cuh"

namespace umma {

constexpr int kPatchW = 8, kPatchH = 16;
constexpr int kHaloItems = 6;   // (pixel, 4-channel chunk) work items per producer thread and halo (<= 1536 items)
constexpr int kHaloMaxStages = 3, kHaloMaxBStages = 4;

struct HaloGeom {
  const float *in;
  int C, H, W, ph, pw, kh, kw, dy, dx;
  int HR, HC;      // halo lines / columns that are filled
  int PX, PY;      // patches per image row / column
  int batch, CB, kk, HW;
  int64_t inImg;
};

struct HaloCtl {
  uint64_t halo_full[kHaloMaxStages], halo_empty[kHaloMaxStages];
  uint64_t b_full[kHaloMaxBStages], b_empty[kHaloMaxBStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  alignas(16) float epi_sb[kEpilogueWarps][64];
};


template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
k_umma_halo(HaloGeom g, ConvEpilogue epi, int mTiles, int nTiles, int BN, int HS, int BS, uint32_t accCols,
            const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ HaloCtl ctl;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int plane = g.HR * g.HC * 16;  // 4 channels of every halo pixel
  const int haloBytes = plane * 8;     // one precision copy (32 channels)
  const int haloStage = haloBytes * (SPLIT ? 2 : 1);
  const int bBytes = BN * BK * 4;
  const int bStage = bBytes * (SPLIT ? 2 : 1);
  uint8_t *bRing = smem;  // TMA tiles need the 1024-byte alignment; the halos follow
  uint8_t *haloRing = smem + (size_t)BS * bStage;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int numTiles = mTiles * nTiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HS; ++s) {
      mbar_init(&ctl.halo_full[s], kProducerWarps);
      mbar_init(&ctl.halo_empty[s], 1);
    }
    for (int s = 0; s < BS; ++s) {
      mbar_init(&ctl.b_full[s], 1);
      mbar_init(&ctl.b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl.tmem_full[s], 1);
      mbar_init(&ctl.tmem_empty[s], kEpilogueWarps);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&ctl.tmem_base, accCols * 2);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmBhi);
    if (SPLIT) tma_prefetch_desc(&tmBlo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl.tmem_base;

  if (warp >= kFirstProducerWarp && warp < kMmaWarp) {
    // ===================== producers: stage the halo of (tile, channel block) once =====================
    const int tid = threadIdx.x - kFirstProducerWarp * 32;
    const int HP = g.HR * g.HC, items = HP * 8;
    // per-thread item table (the same for every halo): shared-memory offset, global offset from the halo origin,
    // and (line, column) for the bounds test
    int soff[kHaloItems], goff[kHaloItems], rc[kHaloItems], q4[kHaloItems];
#pragma unroll
    for (int it = 0; it < kHaloItems; ++it) {
      const int id = it * kProducerThreads + tid;
      if (id < items) {
        const int q = id / HP, p = id - q * HP;
        const int r = p / g.HC, c = p - r * g.HC;
        soff[it] = q * plane + (p << 4);
        goff[it] = q * 4 * g.HW + r * g.W + c;
        rc[it] = (r << 16) | c;
        q4[it] = q * 4;
      } else {
        soff[it] = -1;
        goff[it] = 0;
        rc[it] = 0;
        q4[it] = 0;
      }
    }
    const int patchesPerImg = g.PX * g.PY;
    int hs = 0;
    uint32_t hphase = 0;
    for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
      const int mt = t % mTiles;
      const int n = mt / patchesPerImg;
      const int rr = mt - n * patchesPerImg;
      const int py = rr / g.PX, px = rr - py * g.PX;
      const int ih0 = py * kPatchH - g.ph, iw0 = px * kPatchW - g.pw;
      const float *origin = g.in + (int64_t)n * g.inImg + (int64_t)ih0 * g.W + iw0;
      for (int cb = 0; cb < g.CB; ++cb) {
        float4 v[kHaloItems];
        const float *base = origin + (int64_t)cb * 32 * g.HW;
        const int cleft = g.C - cb * 32;  // channels of this block that exist
#pragma unroll
        for (int it = 0; it < kHaloItems; ++it) {
          const int r = rc[it] >> 16, c = rc[it] & 0xffff;
          const bool inb = soff[it] >= 0 && (unsigned)(ih0 + r) < (unsigned)g.H && (unsigned)(iw0 + c) < (unsigned)g.W;
          const float *q = base + goff[it];
          const int nc = inb ? cleft - q4[it] : 0;
          v[it].x = nc > 0 ? __ldg(q) : 0.f;
          v[it].y = nc > 1 ? __ldg(q + g.HW) : 0.f;
          v[it].z = nc > 2 ? __ldg(q + 2 * g.HW) : 0.f;
          v[it].w = nc > 3 ? __ldg(q + 3 * g.HW) : 0.f;
        }
        mbar_wait_warp(&ctl.halo_empty[hs], hphase ^ 1);
        uint8_t *hi = haloRing + (size_t)hs * haloStage;
        uint8_t *lo = hi + haloBytes;
#pragma unroll
        for (int it = 0; it < kHaloItems; ++it) {
          if (soff[it] >= 0) {
            if (SPLIT) {
              const float4 h = make_float4(tf32_hi(v[it].x), tf32_hi(v[it].y), tf32_hi(v[it].z), tf32_hi(v[it].w));
              *reinterpret_cast<float4 *>(hi + soff[it]) = h;
              *reinterpret_cast<float4 *>(lo + soff[it]) =
                  make_float4(v[it].x - h.x, v[it].y - h.y, v[it].z - h.z, v[it].w - h.w);
            } else {
              *reinterpret_cast<float4 *>(hi + soff[it]) =
                  make_float4(tf32_rna(v[it].x), tf32_rna(v[it].y), tf32_rna(v[it].z), tf32_rna(v[it].w));
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl.halo_full[hs]);
        if (++hs == HS) { hs = 0; hphase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer: kh*kw shifted views of each staged halo =====================
    const uint32_t idesc = make_idesc(BN, false, false);
    const uint32_t halo_base = smem_u32(haloRing), b_base = smem_u32(bRing);
    int hs = 0, bs = 0, it = 0;
    uint32_t hphase = 0, bphase = 0;
    for (int t = blockIdx.x; t < numTiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait_warp(&ctl.tmem_empty[as], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols;
      for (int cb = 0; cb < g.CB; ++cb) {
        mbar_wait_warp(&ctl.halo_full[hs], hphase);
        const uint32_t ha = halo_base + (uint32_t)hs * haloStage;
        int kr = 0, kc = 0;
        for (int tap = 0; tap < g.kk; ++tap) {
          mbar_wait_warp(&ctl.b_full[bs], bphase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t sb = b_base + (uint32_t)bs * bStage;
            const uint32_t shift = (uint32_t)((kr * g.dy * g.HC + kc * g.dx) << 4);
            const uint32_t sbo = (uint32_t)g.HC * 16u;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
              // un-swizzled K-major: layout_type 0, LBO = next 16-byte k chunk, SBO = next 8-row group
              const uint64_t ah = make_desc(ha + shift + ks * 2 * plane, plane, sbo, 0);
              const uint64_t bh = make_desc(sb + ks * 32, 16, 1024, 2);
              const uint32_t acc = (cb != 0 || tap != 0 || ks != 0) ? 1u : 0u;
              if (SPLIT) {
                const uint64_t al = make_desc(ha + haloBytes + shift + ks * 2 * plane, plane, sbo, 0);
                const uint64_t bl = make_desc(sb + bBytes + ks * 32, 16, 1024, 2);
                mma_tf32(tmem_d, al, bh, idesc, acc);
                mma_tf32(tmem_d, ah, bl, idesc, 1u);
                mma_tf32(tmem_d, ah, bh, idesc, 1u);
              } else {
                mma_tf32(tmem_d, ah, bh, idesc, acc);
              }
            }
            mma_commit(&ctl.b_empty[bs]);
            if (tap == g.kk - 1) {
              mma_commit(&ctl.halo_empty[hs]);
              if (cb == g.CB - 1) mma_commit(&ctl.tmem_full[as]);
            }
          }
          __syncwarp();
          if (++bs == BS) { bs = 0; bphase ^= 1; }
          if (++kc == g.kw) { kc = 0; ++kr; }
        }
        if (++hs == HS) { hs = 0; hphase ^= 1; }
      }
    }
  } else if (warp == kTmaWarp) {
    // ===================== TMA producer of the weight blocks: one per (channel block, tap) =====================
    if (lane == 0) {
      const uint32_t b_base = smem_u32(bRing);
      int bs = 0;
      uint32_t bphase = 0;
      for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
        const int nt = t / mTiles;
        const int kbs = g.CB * g.kk;
        for (int kb = 0; kb < kbs; ++kb) {
          mbar_wait(&ctl.b_empty[bs], bphase ^ 1);
          const uint32_t sb = b_base + (uint32_t)bs * bStage;
          mbar_arrive_expect_tx(&ctl.b_full[bs], (uint32_t)bStage);
          tma_load_2d(sb, &tmBhi, kb * BK, nt * BN, &ctl.b_full[bs]);
          if (SPLIT) tma_load_2d(sb + bBytes, &tmBlo, kb * BK, nt * BN, &ctl.b_full[bs]);
          if (++bs == BS) { bs = 0; bphase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue warps 0-3 =====================
    const int quad = warp;
    int it = 0;
    for (int t = blockIdx.x; t < numTiles; t += gridDim.x, ++it) {
      const int mt = t % mTiles, nt = t / mTiles;
      const int as = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int64_t m = (int64_t)mt * BM + quad * 32 + lane;
      const int64_t n0 = (int64_t)nt * BN;
      bool mvalid;
      ConvEpilogue::Row rcx = epi.row(m, mvalid);
      mbar_wait_warp_relaxed(&ctl.tmem_full[as], use & 1, 64);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols + ((uint32_t)(quad * 32) << 16);
      const int nChunks = (BN + 31) >> 5;
      for (int ci = 0; ci < nChunks; ++ci) {
        const int c0 = ci * 32;
        uint32_t r[32];
        tmem_ld32(tmem_d + (uint32_t)c0, r);
        if (ci == nChunks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl.tmem_empty[as]);
        }
        int64_t cnt = min((int64_t)(BN - c0), epi.cols - (n0 + c0));
        if (cnt > 32) cnt = 32;
        if (cnt > 0) epi.store_chunk(rcx, n0 + c0, r, (int)cnt, mvalid, ctl.epi_sb[quad]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, accCols * 2);
}

// weights in halo order: out[row][((cb*kk + tap)*32 + ci)] with channel = cb*32 + ci
//   forward: W[row][channel][tap];   dgrad: W[channel][row][kk-1-tap]
__global__ void __launch_bounds__(256) k_prepare_weights_halo(const float *__restrict__ w, int rows, int chans, int kk,
                                                              int CB, int dgrad, float *__restrict__ hi,
                                                              float *__restrict__ lo, int split) {
  const int Kpad = kk * CB * 32;
  const int64_t total = (int64_t)rows * Kpad;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int r = (int)(i / Kpad);
    const int k = (int)(i - (int64_t)r * Kpad);
    const int ci = k & 31, gq = k >> 5;
    const int cb = gq / kk, tap = gq - cb * kk;
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

// v = the forward problem computed (dgrad passes the virtual one); kNotTaken when the geometry does not fit
inline int conv_halo(tnn_context *ctx, const simt::ConvGeom &v, int batch, const float *in, const float *weights,
                     bool dgradWeights, const float *scale, const float *bias, int act, float *out, bool accumulate,
                     const float *residual) {
  if (v.sy != 1 || v.sx != 1) return kNotTaken;
  const int kk = v.kh * v.kw;
  if (kk < 2 || v.C < 16) return kNotTaken;
  const int HR = kPatchH + (v.kh - 1) * v.dy, HC = kPatchW + (v.kw - 1) * v.dx;
  if (HR * HC * 8 > kHaloItems * kProducerThreads) return kNotTaken;
  const int PX = (v.OW + kPatchW - 1) / kPatchW, PY = (v.OH + kPatchH - 1) / kPatchH;
  const double util = (double)v.OW * v.OH / ((double)PX * kPatchW * PY * kPatchH);
  // measured on B200 (3xTF32, batch 64): 2.0x over the gathering kernel at 32->64 @208, 1.75x at 64->128 @104, 2.0x on
  // the CIFAR shapes; at 128->256 @52 (75 % of the patch rows in use) the MMA-bound gathering kernel wins
  if (util < 0.9) return kNotTaken;
  const int64_t mTiles = (int64_t)batch * PX * PY;
  if (mTiles >= (1ll << 24) || v.inImg >= (1ll << 31)) return kNotTaken;
  const bool split = ctx->math_mode == TNN_MATH_3XTF32;
  const int64_t N = v.F;
  int bn = N >= 128 ? 128 : (int)(((N + 15) / 16) * 16);
  const int CB = (v.C + 31) / 32;
  const int64_t K = (int64_t)kk * CB * 32;
  const int haloStage = HR * HC * 128 * (split ? 2 : 1);
  const int bStage = bn * BK * 4 * (split ? 2 : 1);
  int HS = 2, BS = 4;
  const int budget = 220 * 1024;
  while (BS > 2 && HS * haloStage + BS * bStage > budget) --BS;
  if (HS * haloStage + BS * bStage > budget) return kNotTaken;
  if ((HS + 1) * haloStage + BS * bStage <= budget) ++HS;
  const size_t smemBytes = (size_t)HS * haloStage + (size_t)BS * bStage + 1024;
  const int nTiles = (int)cdiv64(N, bn);
  int rc = tnn_reserve_work(ctx, (size_t)N * K * sizeof(float) * 2);
  if (rc) return rc;
  float *whi = ctx->work, *wlo = ctx->work + (size_t)N * K;
  {
    int64_t blocks = cdiv64(N * K, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_prepare_weights_halo<<<(int)blocks, 256, 0, ctx->stream>>>(weights, (int)N, v.C, kk, CB, dgradWeights ? 1 : 0, whi,
                                                                wlo, split ? 1 : 0);
    TNN_LAUNCHED(ctx);
  }
  alignas(64) CUtensorMap tmHi, tmLo;
  rc = encode_weights_map(&tmHi, whi, N, K, bn);
  if (rc) return rc;
  rc = encode_weights_map(&tmLo, split ? wlo : whi, N, K, bn);
  if (rc) return rc;
  HaloGeom g{in, v.C, v.H, v.W, v.ph, v.pw, v.kh, v.kw, v.dy, v.dx, HR, HC, PX, PY, batch, CB, kk, v.H * v.W, v.inImg};
  ConvEpilogue epi{out, scale, bias, v.OH * v.OW, v.outImg, act, mTiles * BM, N, accumulate,
                   kPatchW, kPatchH, PX, PX * PY, v.OH, v.OW, batch, 7, residual};
  const int64_t numTiles = mTiles * nTiles;
  const int64_t sms = ctx->prop.multiProcessorCount;
  const int grid = (int)(numTiles < sms ? numTiles : sms);
  auto kern = split ? k_umma_halo<true> : k_umma_halo<false>;
  TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  kern<<<grid, kThreads, smemBytes, ctx->stream>>>(g, epi, (int)mTiles, nTiles, bn, HS, BS, pow2_cols(bn), tmHi, tmLo);
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // namespace umma
