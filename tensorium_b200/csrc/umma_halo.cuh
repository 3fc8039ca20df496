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
constexpr int kHaloGroup = 6;   // halo pixels a producer thread keeps in flight (loads) before it stores them
constexpr int kHaloMaxPasses = 12;
constexpr int kHaloMaxStages = 2, kHaloMaxBStages = 3;

// Stride s in {1,2}: input pixel (s*a + ra, s*b + rb) lives in parity plane (ra,rb) at (a,b); tap (kr,kc) with
// kr*dy - ph = s*sa + ra reads plane (ra,rb) shifted by (sa,sb) — a stride-1 access inside the plane.  s = 1 has
// one plane.  Every plane holds HR x HC pixels starting at (oh0 + saMin, ow0 + sbMin) in plane coordinates.
struct HaloGeom {
  const float *in;
  int C, H, W, ph, pw, kh, kw, dy, dx, s;
  int HR, HC;        // lines / columns per parity plane
  int saMin, sbMin;  // smallest line / column shift of any tap
  int NP;            // pixels per halo = s*s*HR*HC
  int CH;            // channels per staged halo: 32 or 16
  int PX, PY;        // patches per image row / column
  int batch, CB, kk, HW;
  int64_t inImg;
};

struct HaloCtl {
  uint64_t halo_full[kHaloMaxStages], halo_empty[kHaloMaxStages];
  uint64_t b_full[kHaloMaxBStages], b_empty[kHaloMaxBStages];
  uint64_t tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
  int tapoff[64];  // byte offset of each tap's view inside one precision copy of the halo
  alignas(16) float epi_sb[kEpilogueWarps][64];
};

__host__ __device__ inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <bool SPLIT, bool EXTRA>
__global__ void __launch_bounds__(kThreads, 1)
k_umma_halo(HaloGeom g, ConvEpilogueT<EXTRA> epi, int mTiles, int nTiles, int BN, int HS, int BS, uint32_t accCols,
            const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ HaloCtl ctl;
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the shared address space
  // and emits LDS / STS instead of generic LD / ST for every access below)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int Q = g.CH >> 2;            // 16-byte channel chunks per pixel
  const int plane = g.NP * 16;        // one chunk (4 channels) of every halo pixel
  const int haloBytes = plane * Q;    // one precision copy
  const int haloStage = haloBytes * (SPLIT ? 2 : 1);
  const int bBytes = BN * g.CH * 4;
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
  if ((int)threadIdx.x < g.kk) {
    const int tap = threadIdx.x;
    const int kr = tap / g.kw, kc = tap - kr * g.kw;
    const int sa = floor_div(kr * g.dy - g.ph, g.s), sb = floor_div(kc * g.dx - g.pw, g.s);
    const int ra = kr * g.dy - g.ph - g.s * sa, rb = kc * g.dx - g.pw - g.s * sb;
    ctl.tapoff[tap] = (((ra * g.s + rb) * g.HR + (sa - g.saMin)) * g.HC + (sb - g.sbMin)) * 16;
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
    // thread -> (chunk q, pixel lane): TPC threads share a chunk and walk the halo pixels TPC at a time, so a warp
    // reads runs of one channel (coalesced) and writes consecutive 16-byte slots (conflict free)
    const int tid = threadIdx.x - kFirstProducerWarp * 32;
    const int TPC = kProducerThreads / Q;  // 32 (CH = 32) or 64 (CH = 16)
    const int q = tid / TPC, p0 = tid - q * TPC;
    const int passes = (g.NP + TPC - 1) / TPC;  // <= kHaloMaxPasses (host checks)
    const int HP = g.HR * g.HC;
    // per-thread pixel table (the same for every halo): plane | line | column packed, -1 = no pixel
    int pix[kHaloMaxPasses];
#pragma unroll
    for (int it = 0; it < kHaloMaxPasses; ++it) {
      const int pid = p0 + it * TPC;
      if (it < passes && pid < g.NP) {
        const int pl = pid / HP, rem = pid - pl * HP;
        const int r = rem / g.HC, c = rem - r * g.HC;
        pix[it] = (pl << 16) | (r << 8) | c;
      } else {
        pix[it] = -1;
      }
    }
    const int patchesPerImg = g.PX * g.PY;
    const int sbase = q * plane + p0 * 16;
    int hs = 0;
    uint32_t hphase = 0;
    for (int t = blockIdx.x; t < numTiles; t += gridDim.x) {
      const int mt = t % mTiles;
      const int n = mt / patchesPerImg;
      const int rr = mt - n * patchesPerImg;
      const int py = rr / g.PX, px = rr - py * g.PX;
      const int a0 = py * kPatchH + g.saMin, b0 = px * kPatchW + g.sbMin;  // plane coordinates of the halo origin
      const float *img = g.in + (int64_t)n * g.inImg;
      for (int cb = 0; cb < g.CB; ++cb) {
        const int c0 = cb * g.CH + q * 4;
        const float *base = img + (int64_t)c0 * g.HW;
        const int nc = g.C - c0;  // channels of this chunk that exist (<= 0: none)
        uint8_t *hi = haloRing + (size_t)hs * haloStage + sbase;
        bool waited = false;
#pragma unroll
        for (int g0 = 0; g0 < kHaloMaxPasses; g0 += kHaloGroup) {
          if (g0 < passes) {
            float4 v[kHaloGroup];
#pragma unroll
            for (int u = 0; u < kHaloGroup; ++u) {
              const int e = pix[g0 + u];
              const int pl = e >> 16, r = (e >> 8) & 0xff, c = e & 0xff;
              const int ra = (g.s == 2) ? (pl >> 1) : 0, rb = (g.s == 2) ? (pl & 1) : 0;
              const int ih = g.s * (a0 + r) + ra, iw = g.s * (b0 + c) + rb;
              const bool inb = e >= 0 && (unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W;
              const float *src = base + (int64_t)ih * g.W + iw;
              const int m = inb ? nc : 0;
              v[u].x = m > 0 ? __ldg(src) : 0.f;
              v[u].y = m > 1 ? __ldg(src + g.HW) : 0.f;
              v[u].z = m > 2 ? __ldg(src + 2 * g.HW) : 0.f;
              v[u].w = m > 3 ? __ldg(src + 3 * g.HW) : 0.f;
            }
            if (!waited) {
              mbar_wait_warp(&ctl.halo_empty[hs], hphase ^ 1);
              waited = true;
            }
#pragma unroll
            for (int u = 0; u < kHaloGroup; ++u) {
              if (pix[g0 + u] >= 0) {
                uint8_t *d = hi + (g0 + u) * TPC * 16;
                if (SPLIT) {
                  const float4 h = make_float4(tf32_hi(v[u].x), tf32_hi(v[u].y), tf32_hi(v[u].z), tf32_hi(v[u].w));
                  *reinterpret_cast<float4 *>(d) = h;
                  *reinterpret_cast<float4 *>(d + haloBytes) =
                      make_float4(v[u].x - h.x, v[u].y - h.y, v[u].z - h.z, v[u].w - h.w);
                } else {
                  *reinterpret_cast<float4 *>(d) =
                      make_float4(tf32_rna(v[u].x), tf32_rna(v[u].y), tf32_rna(v[u].z), tf32_rna(v[u].w));
                }
              }
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
    const int ksteps = g.CH >> 3;  // UMMA_K = 8 channels = two 16-byte chunks
    // A: un-swizzled K-major (layout_type 0), LBO = next 16-byte k chunk (one plane), SBO = next 8-row group (a line)
    const uint64_t a_base = make_desc(halo_base, (uint32_t)plane, (uint32_t)g.HC * 16u, 0);
    const uint32_t a_kstep = (uint32_t)(2 * plane) >> 4;
    // B: weights, SWIZZLE_128B (128-byte rows) or SWIZZLE_64B (64-byte rows)
    const uint64_t b_desc = make_desc(b_base, 16, g.CH == 32 ? 1024u : 512u, g.CH == 32 ? 2u : 4u);
    // ONE elected thread runs the whole issue loop (waits included): no warp-collective operation per tap, and the
    // compiler keeps the descriptors in uniform registers
    if (elect_one_sync()) {
      int hs = 0, bs = 0, it = 0;
      uint32_t hphase = 0, bphase = 0;
      const int kk = g.kk, CB = g.CB;
      for (int t = blockIdx.x; t < numTiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&ctl.tmem_empty[as], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)as * accCols;
        for (int cb = 0; cb < CB; ++cb) {
          mbar_wait(&ctl.halo_full[hs], hphase);
          const uint32_t a_stage = (uint32_t)a_base + (uint32_t)hs * (uint32_t)(haloStage >> 4);
          for (int tap = 0; tap < kk; ++tap) {
            mbar_wait(&ctl.b_full[bs], bphase);
            tc_fence_after();
            // descriptors differ from the per-stage bases only in the 14-bit start-address field (16-byte units)
            uint32_t a_lo32 = a_stage + ((uint32_t)ctl.tapoff[tap] >> 4);
            uint32_t b_lo32 = (uint32_t)b_desc + (uint32_t)bs * (uint32_t)(bStage >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks >= ksteps) break;
              const uint64_t ah = (a_base & 0xFFFFFFFF00000000ull) | a_lo32;
              const uint64_t bh = (b_desc & 0xFFFFFFFF00000000ull) | b_lo32;
              const uint32_t acc = (cb != 0 || tap != 0 || ks != 0) ? 1u : 0u;
              if (SPLIT) {
                const uint64_t al = (a_base & 0xFFFFFFFF00000000ull) | (a_lo32 + (uint32_t)(haloBytes >> 4));
                const uint64_t bl = (b_desc & 0xFFFFFFFF00000000ull) | (b_lo32 + (uint32_t)(bBytes >> 4));
                mma_tf32(tmem_d, al, bh, idesc, acc);
                mma_tf32(tmem_d, ah, bl, idesc, 1u);
                mma_tf32(tmem_d, ah, bh, idesc, 1u);
              } else {
                mma_tf32(tmem_d, ah, bh, idesc, acc);
              }
              a_lo32 += a_kstep;
              b_lo32 += 2;  // 32 bytes = one UMMA_K step inside the swizzled row
            }
            mma_commit(&ctl.b_empty[bs]);
            if (tap == kk - 1) {
              mma_commit(&ctl.halo_empty[hs]);
              if (cb == CB - 1) mma_commit(&ctl.tmem_full[as]);
            }
            if (++bs == BS) { bs = 0; bphase ^= 1; }
          }
          if (++hs == HS) { hs = 0; hphase ^= 1; }
        }
      }
    }
    __syncwarp();
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
          tma_load_2d(sb, &tmBhi, kb * g.CH, nt * BN, &ctl.b_full[bs]);
          if (SPLIT) tma_load_2d(sb + bBytes, &tmBlo, kb * g.CH, nt * BN, &ctl.b_full[bs]);
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
      typename ConvEpilogueT<EXTRA>::Row rcx = epi.row(m, mvalid);
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

// weights in halo order: out[row][((cb*kk + tap)*CH + ci)] with channel = cb*CH + ci
//   forward: W[row][channel][tap];   dgrad: W[channel][row][kk-1-tap]
__global__ void __launch_bounds__(256) k_prepare_weights_halo(const float *__restrict__ w, int rows, int chans, int kk,
                                                              int CB, int CH, int dgrad, float *__restrict__ hi,
                                                              float *__restrict__ lo, int split) {
  const int Kpad = kk * CB * CH;
  const int64_t total = (int64_t)rows * Kpad;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthreads) {
    const int r = (int)(i / Kpad);
    const int k = (int)(i - (int64_t)r * Kpad);
    const int ci = k % CH, gq = k / CH;
    const int cb = gq / kk, tap = gq - cb * kk;
    const int ch = cb * CH + ci;
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

// K-major fp32 matrix [rows][Kpad] -> tiles of CH k x boxRows rows: SWIZZLE_128B for 128-byte rows, 64B for 64-byte
inline int encode_weights_map_ch(CUtensorMap *tm, const float *base, int64_t rows, int64_t Kpad, int boxRows, int CH) {
  if (CH == 32) return encode_weights_map(tm, base, rows, Kpad, boxRows);
  EncodeTiledFn enc = get_encode_tiled();
  TNN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * 4};
  cuuint32_t box[2] = {(cuuint32_t)CH, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tnn_set_error("cuTensorMapEncodeTiled (64B) failed with CUresult %d", (int)r);
    return 6;
  }
  return 0;
}

// v = the forward problem computed (dgrad passes the virtual one); kNotTaken when the geometry does not fit
inline int conv_halo(tnn_context *ctx, const simt::ConvGeom &v, int batch, const float *in, const float *weights,
                     bool dgradWeights, const float *scale, const float *bias, int act, float *out, bool accumulate,
                     const float *residual) {
  if (v.sy != v.sx || (v.sy != 1 && v.sy != 2)) return kNotTaken;
  const int s = v.sy;
  const int kk = v.kh * v.kw;
  // (a 3-channel first layer was tried here too: 1.6 ms against 1.0 ms for the direct fp32 kernel)
  if (kk < 2 || kk > 64 || v.C < 16) return kNotTaken;
  // shift ranges of the taps inside a parity plane
  int saMin = 1 << 30, saMax = -(1 << 30), sbMin = 1 << 30, sbMax = -(1 << 30);
  for (int kr = 0; kr < v.kh; ++kr) {
    const int sa = floor_div(kr * v.dy - v.ph, s);
    saMin = sa < saMin ? sa : saMin;
    saMax = sa > saMax ? sa : saMax;
  }
  for (int kc = 0; kc < v.kw; ++kc) {
    const int sb = floor_div(kc * v.dx - v.pw, s);
    sbMin = sb < sbMin ? sb : sbMin;
    sbMax = sb > sbMax ? sb : sbMax;
  }
  const int HR = kPatchH + saMax - saMin, HC = kPatchW + sbMax - sbMin;
  const int NP = s * s * HR * HC;
  if (HR > 255 || HC > 255) return kNotTaken;
  const bool split = tnn_math_is_split(ctx);
  // channels per staged halo: 32 unless two stages of it do not fit next to the weight ring
  int CH = 32;
  if ((size_t)NP * 32 * 4 * (split ? 2 : 1) * 2 > 150 * 1024) CH = 16;
  const int TPC = kProducerThreads / (CH / 4);
  if ((NP + TPC - 1) / TPC > kHaloMaxPasses) return kNotTaken;
  const int PX = (v.OW + kPatchW - 1) / kPatchW, PY = (v.OH + kPatchH - 1) / kPatchH;
  const double util = (double)v.OW * v.OH / ((double)PX * kPatchW * PY * kPatchH);
  // measured on B200 (3xTF32, batch 64): 2.0x over the gathering kernel at 32->64 @208, 1.75x at 64->128 @104, 2.0x on
  // the CIFAR shapes; at 128->256 @52 (75 % of the patch rows in use) the MMA-bound gathering kernel wins
  if (util < 0.9) return kNotTaken;
  const int64_t mTiles = (int64_t)batch * PX * PY;
  if (mTiles >= (1ll << 24) || v.inImg >= (1ll << 31)) return kNotTaken;
  const int64_t N = v.F;
  const int bn = N >= 128 ? 128 : (int)(((N + 15) / 16) * 16);
  const int CB = (v.C + CH - 1) / CH;
  const int64_t K = (int64_t)kk * CB * CH;
  const int haloStage = NP * CH * 4 * (split ? 2 : 1);
  const int bStage = bn * CH * 4 * (split ? 2 : 1);
  if ((bStage % 1024) != 0) return kNotTaken;  // TMA tiles of the weight ring stay 1024-byte aligned
  // Two halo stages and three weight stages.  More of either is SLOWER (measured: 0.67 -> 0.78 ms at 32->64 @208
  // with 3+4 stages, 1.02 -> 1.48 ms on the stride-2 layer with 2+8): shared memory comes out of the same 256 KB as
  // the L1 that serves the producers' 32-byte-sector halo reads.
  const int HS = 2, BS = 3;
  if (HS * haloStage + BS * bStage > 220 * 1024) return kNotTaken;
  const size_t smemBytes = (size_t)HS * haloStage + (size_t)BS * bStage + 1024;
  const int nTiles = (int)cdiv64(N, bn);
  float *whi, *wlo;
  bool fresh;
  int rc = tnn_weight_buffer(ctx, weights, (dgradWeights ? 6 : 5) + (CH == 16 ? 2 : 0), (size_t)N * K, &whi, &wlo, &fresh);
  if (rc) return rc;
  if (fresh) {
    int64_t blocks = cdiv64(N * K, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_prepare_weights_halo<<<(int)blocks, 256, 0, ctx->stream>>>(weights, (int)N, v.C, kk, CB, CH, dgradWeights ? 1 : 0,
                                                                whi, wlo, split ? 1 : 0);
    TNN_LAUNCHED(ctx);
  }
  alignas(64) CUtensorMap tmHi, tmLo;
  rc = encode_weights_map_ch(&tmHi, whi, N, K, bn, CH);
  if (rc) return rc;
  rc = encode_weights_map_ch(&tmLo, split ? wlo : whi, N, K, bn, CH);
  if (rc) return rc;
  HaloGeom g{in, v.C, v.H, v.W, v.ph, v.pw, v.kh, v.kw, v.dy, v.dx, s, HR, HC, saMin, sbMin, NP, CH, PX, PY, batch, CB, kk,
             v.H * v.W, v.inImg};
  const int64_t numTiles = mTiles * nTiles;
  const int64_t sms = ctx->prop.multiProcessorCount;
  const int grid = (int)(numTiles < sms ? numTiles : sms);
  if (accumulate || residual) {
    ConvEpilogueT<true> epi{out, scale, bias, v.OH * v.OW, v.outImg, act, mTiles * BM, N, accumulate,
                            kPatchW, kPatchH, PX, PX * PY, v.OH, v.OW, batch, 7, residual};
    auto kern = split ? k_umma_halo<true, true> : k_umma_halo<false, true>;
    TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    kern<<<grid, kThreads, smemBytes, ctx->stream>>>(g, epi, (int)mTiles, nTiles, bn, HS, BS, pow2_cols(bn), tmHi, tmLo);
  } else {
    ConvEpilogueT<false> epi{out, scale, bias, v.OH * v.OW, v.outImg, act, mTiles * BM, N, false,
                             kPatchW, kPatchH, PX, PX * PY, v.OH, v.OW, batch, 7, nullptr};
    auto kern = split ? k_umma_halo<true, false> : k_umma_halo<false, false>;
    TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    kern<<<grid, kThreads, smemBytes, ctx->stream>>>(g, epi, (int)mTiles, nTiles, bn, HS, BS, pow2_cols(bn), tmHi, tmLo);
  }
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // namespace umma
