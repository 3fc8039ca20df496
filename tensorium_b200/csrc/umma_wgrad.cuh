// umma_wgrad.cuh — convolution weight gradient on tcgen05 for stride-1 layers (nConvolutionLayer.pas:1203-1232:
// per-image im2col + NT gemm with beta = 1, accumulated over the batch):
//     weight_updates[f][(c,kr,kc)] += sum_{b,oh,ow} delta[b][f][oh][ow] * in[b][c][oh - ph + kr*dy][ow - pw + kc*dx]
//
// The contraction runs over PIXELS, so a kernel tap is a shift ALONG k.  Both operands are staged K-major in the
// un-swizzled canonical layout — 16-byte chunks of 4 consecutive pixels, [chunk][row][16 B]:
//   A = delta   rows f (128, zero beyond F)          one copy per pixel tile
//   B = input   rows c (32 channels per CTA)         kw copies, copy kc pre-shifted by (kc*dx - pw) pixels, each holding
//                                                    the TH + (kh-1)*dy lines the taps of that column need
// A row shift (kr) is a whole number of chunks, i.e. a descriptor start offset; the column shift is baked into the kw
// copies (3 gathers per input element instead of kh*kw = 9 in the im2col formulation).  Every tap owns 32 TMEM
// columns (kh*kw*32 <= 512), the accumulators live across ALL pixel tiles a CTA walks, and one epilogue per CTA adds
// them to weight_updates with fp32 atomics (the pixel range is split across CTAs).
//
// A pixel tile is 32 output pixels = TH lines x TW columns (TW in {32,16,8}); chunk strides are padded by 16 bytes
// (LBO = rows*16 + 16) so that the producers' 16-byte stores — lanes along the pixel chunks, which keeps the global
// reads coalesced — spread over all banks.
#pragma once
#include "umma_conv.cuh"

namespace umma {

constexpr int kWgKT = 32;                       // pixels per tile (contraction block)
constexpr int kWgARows = 128, kWgBRows = 32;    // delta rows per CTA (M), input channels per CTA (N per tap)
constexpr int kWgLboA = kWgARows * 16 + 16;     // 2064
constexpr int kWgLboB = kWgBRows * 16 + 16;     // 528
constexpr int kWgStages = 2;
constexpr int kWgMaxA = 4, kWgMaxB = 9;         // work items (16-byte chunks) per producer thread and tile

struct WgradGeom {
  const float *in, *delta;
  float *wu;
  int C, H, W, F, OH, OW, ph, pw, kh, kw, dy, dx;
  int TW, TH, CPL;    // tile shape; 16-byte chunks per tile line = TW/4
  int HL;             // input lines per copy = TH + (kh-1)*dy
  int tilesX, tilesY; // tiles per image row / column
  int numTiles;       // batch * tilesX * tilesY
  int CBn, FTn;       // channel blocks (32) and filter tiles (128): units = CBn*FTn
  int G;              // CTAs per unit (the pixel tiles of a unit are split over them)
  int kk, HW, P;
  int64_t inImg, outImg;
  int ldw;            // C*kk
};

struct WgradCtl {
  uint64_t full[kWgStages], empty[kWgStages], done;
  uint32_t tmem_base;
};

template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) k_umma_wgrad(WgradGeom g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ WgradCtl ctl;
  uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // keeps the shared address space (LDS / STS)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int aBytes = (kWgKT / 4) * kWgLboA;             // one precision copy of the delta tile
  const int bCopy = g.HL * g.CPL * kWgLboB;             // one pre-shifted input copy, one precision
  const int bBytes = g.kw * bCopy;
  const int stageBytes = (aBytes + bBytes) * (SPLIT ? 2 : 1);
  const int unit = blockIdx.x % (g.CBn * g.FTn), slice = blockIdx.x / (g.CBn * g.FTn);
  const int cb = unit % g.CBn, ft = unit / g.CBn;
  const int c0 = cb * kWgBRows, f0 = ft * kWgARows;
  const int Fv = min(kWgARows, g.F - f0);
  const int myTiles = (g.numTiles - slice + g.G - 1) / g.G;   // tiles slice, slice+G, ...

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&ctl.full[s], kProducerWarps);
      mbar_init(&ctl.empty[s], 1);
    }
    mbar_init(&ctl.done, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(&ctl.tmem_base, 512);
  // rows of A beyond F stay zero for the whole kernel: clear every stage's delta region once
  for (int i = threadIdx.x; i < kWgStages * stageBytes / 16; i += blockDim.x)
    reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl.tmem_base;

  if (warp >= kFirstProducerWarp && warp < kMmaWarp) {
    // ===================== producers: delta tile + kw pre-shifted input copies =====================
    const int tid = threadIdx.x - kFirstProducerWarp * 32;
    const int tilesPerImg = g.tilesX * g.tilesY;
    const int aItems = Fv * (kWgKT / 4);                 // (f, chunk)
    const int bItems = g.kw * g.HL * g.CPL * kWgBRows;   // (kc, line, chunk j, c)
    int stage = 0;
    uint32_t phase = 0;
    // per-thread item tables (the same for every tile): shared-memory offset and source offset from the tile origin
    int aSm[kWgMaxA], aSrc[kWgMaxA];
#pragma unroll
    for (int u = 0; u < kWgMaxA; ++u) {
      const int it = tid + u * kProducerThreads;
      if (it < aItems) {
        const int f = it / (kWgKT / 4), q = it - f * (kWgKT / 4);
        const int dh = q / g.CPL, j = q - dh * g.CPL;
        aSm[u] = q * kWgLboA + f * 16;
        aSrc[u] = f * g.P + dh * g.OW + 4 * j;
      } else {
        aSm[u] = -1;
        aSrc[u] = 0;
      }
    }
    int bSm[kWgMaxB], bSrc[kWgMaxB], bLC[kWgMaxB];  // bLC: line | column offset (+ 64 so it is non-negative) packed
#pragma unroll
    for (int u = 0; u < kWgMaxB; ++u) {
      const int it = tid + u * kProducerThreads;
      if (it < bItems) {
        const int j = it % g.CPL;
        int rest = it / g.CPL;
        const int c = rest % kWgBRows;
        rest /= kWgBRows;
        const int l = rest % g.HL, kc = rest / g.HL;
        const int dw = kc * g.dx - g.pw + 4 * j;
        bSm[u] = (c0 + c < g.C) ? kc * bCopy + (l * g.CPL + j) * kWgLboB + c * 16 : -2 - (kc * bCopy + (l * g.CPL + j) * kWgLboB + c * 16);
        bSrc[u] = c * g.HW + (l - g.ph) * g.W + dw;
        bLC[u] = (l << 16) | (dw + 64);
      } else {
        bSm[u] = -1;
        bSrc[u] = 0;
        bLC[u] = 0;
      }
    }
    for (int i = 0; i < myTiles; ++i) {
      const int t = slice + i * g.G;
      const int b = t / tilesPerImg;
      const int r = t - b * tilesPerImg;
      const int ty = r / g.tilesX, tx = r - ty * g.tilesX;
      const int oh0 = ty * g.TH, ow0 = tx * g.TW;
      const float *dsrc = g.delta + (int64_t)b * g.outImg + (int64_t)f0 * g.P + oh0 * g.OW + ow0;
      const float *isrc = g.in + (int64_t)b * g.inImg + (int64_t)c0 * g.HW + (int64_t)oh0 * g.W + ow0;
      // ---- all loads of the tile first (in flight together, and before the wait: they overlap the MMAs still
      //      reading this stage), then the stores
      float4 va[kWgMaxA], vb[kWgMaxB];
#pragma unroll
      for (int u = 0; u < kWgMaxA; ++u)
        va[u] = aSm[u] >= 0 ? __ldg(reinterpret_cast<const float4 *>(dsrc + aSrc[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kWgMaxB; ++u) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int ih = oh0 - g.ph + (bLC[u] >> 16);
        const int iw = ow0 + (bLC[u] & 0xffff) - 64;
        if (bSm[u] >= 0 && (unsigned)ih < (unsigned)g.H) {
          const float *src = isrc + bSrc[u];
          if ((unsigned)(iw + 0) < (unsigned)g.W) v.x = __ldg(src + 0);
          if ((unsigned)(iw + 1) < (unsigned)g.W) v.y = __ldg(src + 1);
          if ((unsigned)(iw + 2) < (unsigned)g.W) v.z = __ldg(src + 2);
          if ((unsigned)(iw + 3) < (unsigned)g.W) v.w = __ldg(src + 3);
        }
        vb[u] = v;
      }
      mbar_wait_warp(&ctl.empty[stage], phase ^ 1);
      uint8_t *sa = smem + (size_t)stage * stageBytes;
      uint8_t *sb = sa + aBytes * (SPLIT ? 2 : 1);
#pragma unroll
      for (int u = 0; u < kWgMaxA; ++u) {
        if (aSm[u] >= 0) {
          uint8_t *d = sa + aSm[u];
          const float4 v = va[u];
          if (SPLIT) {
            const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            *reinterpret_cast<float4 *>(d) = h;
            *reinterpret_cast<float4 *>(d + aBytes) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
          } else {
            *reinterpret_cast<float4 *>(d) = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kWgMaxB; ++u) {
        if (bSm[u] != -1) {
          uint8_t *d = sb + (bSm[u] >= 0 ? bSm[u] : -2 - bSm[u]);  // channels past C: the zero loaded above is stored
          const float4 v = vb[u];
          if (SPLIT) {
            const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
            *reinterpret_cast<float4 *>(d) = h;
            *reinterpret_cast<float4 *>(d + bBytes) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
          } else {
            *reinterpret_cast<float4 *>(d) = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl.full[stage]);
      if (++stage == kWgStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer: kh*kw taps x 4 k-steps per tile, accumulators persist =====================
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(kWgBRows, false, false);
      const uint32_t sbase = smem_u32(smem);
      // un-swizzled K-major: LBO = next 16-byte k chunk, SBO = next 8-row group (128 B)
      const uint64_t a_desc = make_desc(sbase, kWgLboA, 128, 0);
      const uint64_t b_desc = make_desc(sbase, kWgLboB, 128, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < myTiles; ++i) {
        mbar_wait(&ctl.full[stage], phase);
        tc_fence_after();
        const uint32_t sa = (uint32_t)stage * (uint32_t)stageBytes;
        const uint32_t sb = sa + (uint32_t)aBytes * (SPLIT ? 2 : 1);
        int kr = 0, kc = 0;
        for (int tap = 0; tap < g.kk; ++tap) {
          const uint32_t tmem_d = tmem_base + (uint32_t)tap * kWgBRows;
#pragma unroll
          for (int s = 0; s < kWgKT / 8; ++s) {
            const int q0 = 2 * s;
            const int dh = q0 / g.CPL, j0 = q0 - dh * g.CPL;
            const uint32_t aoff = sa + (uint32_t)q0 * kWgLboA;
            const uint32_t boff = sb + (uint32_t)kc * bCopy + (uint32_t)((dh + kr * g.dy) * g.CPL + j0) * kWgLboB;
            const uint64_t ah = (a_desc & 0xFFFFFFFF00000000ull) | ((uint32_t)a_desc + (aoff >> 4));
            const uint64_t bh = (b_desc & 0xFFFFFFFF00000000ull) | ((uint32_t)b_desc + (boff >> 4));
            const uint32_t acc = (i != 0 || s != 0) ? 1u : 0u;
            if (SPLIT) {
              const uint64_t al = (a_desc & 0xFFFFFFFF00000000ull) | ((uint32_t)a_desc + ((aoff + aBytes) >> 4));
              const uint64_t bl = (b_desc & 0xFFFFFFFF00000000ull) | ((uint32_t)b_desc + ((boff + bBytes) >> 4));
              mma_tf32(tmem_d, al, bh, idesc, acc);
              mma_tf32(tmem_d, ah, bl, idesc, 1u);
              mma_tf32(tmem_d, ah, bh, idesc, 1u);
            } else {
              mma_tf32(tmem_d, ah, bh, idesc, acc);
            }
          }
          if (++kc == g.kw) { kc = 0; ++kr; }
        }
        mma_commit(&ctl.empty[stage]);
        if (i == myTiles - 1) mma_commit(&ctl.done);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp < kEpilogueWarps) {
    // ===================== epilogue (once per CTA): TMEM -> atomic adds into weight_updates =====================
    if (myTiles > 0) {
      mbar_wait_warp(&ctl.done, 0);
      tc_fence_after();
      const int f = warp * 32 + lane;  // TMEM lane = delta row
      const bool fvalid = f < Fv;
      float *row = g.wu + (int64_t)(f0 + (fvalid ? f : 0)) * g.ldw;
      const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16);
      if (g.kk == 9 && (g.C & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0 && (g.ldw & 3) == 0) {
        // 3x3: the 4 channels x 9 taps a thread holds for one channel group are 36 CONSECUTIVE floats of its weight
        // row (index c*9 + tap): nine 16-byte vector reductions instead of 36 scalar ones, each of which would be
        // its own L2 operation (every lane owns a different row)
        for (int cg = 0; cg < kWgBRows / 4; ++cg) {  // warp-uniform trip count
          uint32_t r[9][4];
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) tmem_ld4_nowait(tbase + (uint32_t)(tap * kWgBRows + cg * 4), r[tap]);
          tmem_ld_wait();
          const int c = c0 + cg * 4;
          if (fvalid && c < g.C) {
            float *q = row + (int64_t)c * 9;
#pragma unroll
            for (int v = 0; v < 9; ++v) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q + 4 * v),
                           "f"(__uint_as_float(r[(4 * v) % 9][(4 * v) / 9])),
                           "f"(__uint_as_float(r[(4 * v + 1) % 9][(4 * v + 1) / 9])),
                           "f"(__uint_as_float(r[(4 * v + 2) % 9][(4 * v + 2) / 9])),
                           "f"(__uint_as_float(r[(4 * v + 3) % 9][(4 * v + 3) / 9]))
                           : "memory");
            }
          }
        }
      } else {
        for (int tap = 0; tap < g.kk; ++tap) {
          uint32_t r[32];
          tmem_ld32(tbase + (uint32_t)tap * kWgBRows, r);
          if (fvalid) {
#pragma unroll
            for (int c = 0; c < kWgBRows; ++c)
              if (c0 + c < g.C) atomicAdd(row + (int64_t)(c0 + c) * g.kk + tap, __uint_as_float(r[c]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

// kNotTaken when the geometry does not fit (caller falls back to the split-K GEMM with the gathering loader)
inline int conv_wgrad_tiles(tnn_context *ctx, const simt::ConvGeom &g, int batch, const float *in, const float *delta,
                            float *weight_updates) {
  if (g.sy != 1 || g.sx != 1) return kNotTaken;
  const int kk = g.kh * g.kw;
  if (kk * kWgBRows > 512) return kNotTaken;                                  // one 32-column accumulator per tap
  if ((g.OW % 4) != 0 || (((uintptr_t)delta) & 15) != 0) return kNotTaken;    // 16-byte delta chunks
  int TW = 0;
  for (int tw = 32; tw >= 8; tw >>= 1)
    if (g.OW % tw == 0 && g.OH % (kWgKT / tw) == 0) { TW = tw; break; }
  if (!TW) return kNotTaken;
  const int TH = kWgKT / TW, CPL = TW / 4;
  const int HL = TH + (g.kh - 1) * g.dy;
  if (g.kw * HL * CPL * kWgBRows > kWgMaxB * kProducerThreads || HL > 255 || g.pw > 60) return kNotTaken;
  const bool split = tnn_math_is_split(ctx);
  const int aBytes = (kWgKT / 4) * kWgLboA, bBytes = g.kw * HL * CPL * kWgLboB;
  const size_t stageBytes = (size_t)(aBytes + bBytes) * (split ? 2 : 1);
  const size_t smemBytes = kWgStages * stageBytes + 128;
  if (smemBytes > 220 * 1024) return kNotTaken;
  const int tilesX = g.OW / TW, tilesY = g.OH / TH;
  const int64_t numTiles = (int64_t)batch * tilesX * tilesY;
  if (numTiles >= (1ll << 30) || g.inImg >= (1ll << 31) || g.outImg >= (1ll << 31)) return kNotTaken;
  const int CBn = (g.C + kWgBRows - 1) / kWgBRows, FTn = (g.F + kWgARows - 1) / kWgARows;
  const int units = CBn * FTn;
  const int sms = ctx->prop.multiProcessorCount;
  int G = sms / units;
  if (G < 1) G = 1;
  if (G > numTiles) G = (int)numTiles;
  WgradGeom w{in, delta, weight_updates, g.C, g.H, g.W, g.F, g.OH, g.OW, g.ph, g.pw, g.kh, g.kw, g.dy, g.dx,
              TW, TH, CPL, HL, tilesX, tilesY, (int)numTiles, CBn, FTn, G, kk, g.H * g.W, g.OH * g.OW, g.inImg, g.outImg,
              g.C * kk};
  auto kern = split ? k_umma_wgrad<true> : k_umma_wgrad<false>;
  TNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  kern<<<units * G, kThreads, smemBytes, ctx->stream>>>(w);
  TNN_LAUNCHED(ctx);
  return 0;
}

}  // namespace umma
