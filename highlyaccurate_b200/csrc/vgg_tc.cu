// tcgen05 / TMA implicit-GEMM 3x3 convolution for the VGG16 U-Net (VGG.py:121-158), sm_100a.
//
// GEMM view per layer: M = B*H*W output pixels (tile = 8 rows x 16 cols = 128 pixels = UMMA M),
// N = Cout (tile BLOCK_N <= 128), K = 9 taps x Cin (k-block = one tap x 64 channels = one
// 128-byte swizzle atom).  Activations are NHWC fp16 with two planes per pixel,
// [B][H][W][2][C]: plane 0 = hi = fp16(x), plane 1 = lo = fp16((x - hi) * 2^11).  A shifted
// (tap) view of the pixel tile is ONE TMA box of the 5-D tensor (C, plane, W, H, B) at signed
// coordinates — TMA zero-fills the padding ring, so there is no im2col buffer and no halo code.
// Weights are [9][Cout][Cin] fp16 (K-major B operand), also hi/lo.
//
// Precision (HA_CONV_F16X3): x*w ~= hi_x*hi_w + 2^-11 (hi_x*lo_w + lo_x*hi_w): three
// kind::f16 MMAs per k-step into two fp32 TMEM accumulators, error ~2^-22 per product, i.e.
// fp32-grade features from the fp16 tensor pipe.  HA_CONV_F16 issues only the first MMA.
//
// Warp roles (256 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer, warp 1 =
// MMA issuer (one elected lane), warp 2 = TMEM allocator, warps 4-7 = epilogue, one per TMEM lane
// quarter (TMEM -> registers -> bias / ReLU / 2x2 max-pool by warp shuffles -> swizzled smem staging
// row per pixel -> coalesced 64/128-byte segments to global, incl. the x2 nearest upsample).
// smem ring of STAGES k-blocks (full/empty mbarriers), two accumulator stages in TMEM
// (tmem_full/tmem_empty mbarriers) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <unordered_map>

#include "tc_ptx.cuh"
#include "vgg_common.cuh"

namespace ha {

// ------------------------------------------------------------------------------ the kernel
constexpr int kTileW = 16, kTileH = 8, kBlockM = kTileW * kTileH;   // 128 pixels = UMMA M
constexpr int kBlockK = 64;                                           // fp16 elements = 128 B
constexpr int kTcThreads = 256;                                       // 4 control warps + 4 epilogue warps
constexpr int kABytes = kBlockM * kBlockK * 2;                        // 16 KB

struct TcConvArgs {
  const float* bias;                          // [cout] or null
  __half* act_full; int af_pitch, af_coff;    // relu(v)       -> [B][H][W][2][pitch] + coff
  __half* act_pool; int ap_pitch, ap_coff;    // relu(pool(v)) -> [B][H/2][W/2][2][pitch] + coff
  __half* act_up; int au_pitch, au_coff;      // relu(x2 nearest upsample of feat) -> [B][2h][2w][2][pitch] + coff
  float* feat; int feat_pooled;               // v or pool(v), raw fp32 -> [B][h][w][cout]
  int B, H, W, cout;
  int n_kchunks;                              // ceil(cin / 64)
  int n_taps;                                 // 9 (3x3, pad 1) or 1 (1x1: conv0 over im2col channels)
  int tiles_x, tiles_y, tiles_n, n_tiles;
};

template <int BLOCK_N, bool SPLIT>
struct TcCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = (kABytes + kBBytes) * (SPLIT ? 2 : 1);
  static constexpr int kStages = (192 * 1024 / kStageBytes) > 8 ? 8 : (192 * 1024 / kStageBytes);
  static constexpr int kStagingBytes = 4 * 32 * 128;     // epilogue: one 32-row x 128-byte staging tile per epilogue warp
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kStagingBytes;
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
// two fp32 values -> packed fp16 hi pair and packed fp16 (v - hi) * 2^11 pair
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(__fmaf_rn(v0, kLoScale, -hf.x * kLoScale), __fmaf_rn(v1, kLoScale, -hf.y * kLoScale));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int NV>
__device__ __forceinline__ void store_split(__half* dst_hi, int pitch, const float (&v)[NV], int n) {
  // v[0..n) -> fp16 hi at dst_hi[0..n), fp16 lo*2^11 at dst_hi[pitch + 0..n); n in {16, 32}.
  // Two elements per conversion: cvt.rn.f16x2.f32 for hi, one FFMA each for (v - hi) * 2^11, cvt again for lo.
#pragma unroll
  for (int j = 0; j < NV; j += 8) {
    if (j >= n) break;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v0 = v[j + 2 * e], v1 = v[j + 2 * e + 1];
      const __half2 h = __floats2half2_rn(v0, v1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(__fmaf_rn(v0, kLoScale, -hf.x * kLoScale), __fmaf_rn(v1, kLoScale, -hf.y * kLoScale));
      hi[e] = *reinterpret_cast<const uint32_t*>(&h);
      lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(dst_hi + j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(dst_hi + pitch + j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <int BLOCK_N, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_bh,
                  const __grid_constant__ CUtensorMap tmap_bl, const TcConvArgs a) {
  using Cfg = TcCfg<BLOCK_N, SPLIT>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int CH = BLOCK_N >= 32 ? 32 : 16;            // accumulator columns per TMEM load
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * STAGES;  // [2]       MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;     // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* staging = smem + STAGES * Cfg::kStageBytes + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_bh);
    if (SPLIT) tma_prefetch_desc(&tmap_bl);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_kb = a.n_taps * a.n_kchunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int n_idx = tile % a.tiles_n; int pt = tile / a.tiles_n;
        const int x0 = (pt % a.tiles_x) * kTileW; pt /= a.tiles_x;
        const int y0 = (pt % a.tiles_y) * kTileH; const int b = pt / a.tiles_y;
        // nested tap / channel-chunk loops: no integer division on the producer's critical path
        int tap = 0, kc = 0, ky = a.n_taps == 1 ? 1 : 0, kx = ky;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* st = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(full + stage, Cfg::kStageBytes);
          // stage layout: [A_hi 16 KB][B_hi][B_lo][A_lo 16 KB] — B_hi and B_lo adjacent so that ONE MMA with
          // N = 2*BLOCK_N multiplies A_hi by both (see the MMA issuer)
          tma_load_5d(st, &tmap_a, full + stage, kc * kBlockK, 0, x0 + kx - 1, y0 + ky - 1, b);
          tma_load_3d(st + kABytes, &tmap_bh, full + stage, kc * kBlockK, n_idx * BLOCK_N, tap);
          if (SPLIT) {
            tma_load_3d(st + kABytes + Cfg::kBBytes, &tmap_bl, full + stage, kc * kBlockK, n_idx * BLOCK_N, tap);
            tma_load_5d(st + kABytes + 2 * Cfg::kBBytes, &tmap_a, full + stage, kc * kBlockK, 1, x0 + kx - 1, y0 + ky - 1, b);
          }
          if (++kc == a.n_kchunks) { kc = 0; ++tap; if (++kx == 3) { kx = 0; ++ky; } }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_f16(BLOCK_N);             // A x B_hi
    constexpr uint32_t idesc2 = umma_idesc_f16(2 * BLOCK_N);        // A x [B_hi ; B_lo]
    // stage layout [A_hi][B_hi][B_lo][A_lo]: descriptors of stage 0, K advance = +32 B = +2 in the address field
    const uint64_t desc_a_hi = umma_desc_sw128(smem_u32(smem));
    const uint64_t desc_b = umma_desc_sw128(smem_u32(smem) + kABytes);
    const uint64_t desc_a_lo = umma_desc_sw128(smem_u32(smem) + kABytes + 2 * Cfg::kBBytes);
    int stage = 0; uint32_t phase = 0; int t = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++t) {
      const int as = t & 1; const uint32_t aphase = (t >> 1) & 1;
      mbar_wait(tmem_empty + as, aphase ^ 1);
      tc_fence_after();
      // accumulator columns of this stage: [0, BLOCK_N) = hi*hi, [BLOCK_N, 2*BLOCK_N) = hi*lo + lo*hi
      const uint32_t acc0 = tmem_base + as * 256, acc1 = acc0 + BLOCK_N;
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait(full + stage, phase);
        tc_fence_after();
        if (lane == 0) {
          // The single issuing thread is the critical path of short-K (N = 64) layers: descriptors are one
          // 64-bit add away from precomputed bases (the 14-bit address field cannot carry: smem < 256 KB).
          const uint64_t so = (uint64_t)((uint32_t)(stage * Cfg::kStageBytes) >> 4);
          const uint64_t da0 = desc_a_hi + so, db0 = desc_b + so, dl0 = desc_a_lo + so;
          if (SPLIT) {
            // one N = 2*BLOCK_N MMA reads A_hi once for both hi_x*hi_w and hi_x*lo_w, then lo_x*hi_w is added
            umma_f16(acc0, da0, db0, idesc2, kb != 0);
            umma_f16(acc1, dl0, db0, idesc, 1);
#pragma unroll
            for (int k = 1; k < kBlockK / 16; ++k) {
              umma_f16(acc0, da0 + 2 * k, db0 + 2 * k, idesc2, 1);
              umma_f16(acc1, dl0 + 2 * k, db0 + 2 * k, idesc, 1);
            }
          } else {
            umma_f16(acc0, da0, db0, idesc, kb != 0);
#pragma unroll
            for (int k = 1; k < kBlockK / 16; ++k) umma_f16(acc0, da0 + 2 * k, db0 + 2 * k, idesc, 1);
          }
          umma_commit(empty + stage);                       // smem slot reusable once these MMAs retire
          if (kb == total_kb - 1) umma_commit(tmem_full + as);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const uint32_t stg = smem_u32(staging) + (warp - 4) * (32 * 128);
    const bool any_pool = a.act_pool != nullptr || a.feat_pooled;
    int t = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++t) {
      const int as = t & 1; const uint32_t aphase = (t >> 1) & 1;
      const int n_idx = tile % a.tiles_n; int pt = tile / a.tiles_n;
      const int x0 = (pt % a.tiles_x) * kTileW; pt /= a.tiles_x;
      const int y0 = (pt % a.tiles_y) * kTileH; const int b = pt / a.tiles_y;
      mbar_wait(tmem_full + as, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += CH) {
        uint32_t r0[32], r1[32];
        if (CH == 32) { tmem_ld_x32(taddr + c0, r0); if (SPLIT) tmem_ld_x32(taddr + BLOCK_N + c0, r1); }
        else { tmem_ld_x16(taddr + c0, r0); if (SPLIT) tmem_ld_x16(taddr + BLOCK_N + c0, r1); }
        tmem_ld_wait();
        const int n0 = n_idx * BLOCK_N + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          const float4 bz = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float bb[4] = {bz.x, bz.y, bz.z, bz.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float acc = __uint_as_float(r0[j + e]);
            if (SPLIT) acc = fmaf(__uint_as_float(r1[j + e]), kLoInvScale, acc);
            v[j + e] = acc + bb[e];
          }
        }
        float pv[32];
        if (any_pool) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            pv[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
          }
        }
        // Every lane owns one pixel, and pixels are >= 2*CH bytes apart in every destination, so direct stores
        // touch 32 lines per instruction.  Instead each lane parks its 128-byte row in a warp-private,
        // XOR-swizzled staging tile and the warp writes it out with N16 lanes per row (whole 64 / 128-byte segments).
        constexpr int N16 = CH / 4;                    // 16-byte pieces per staged row (fp32: CH floats; fp16: hi | lo)
        constexpr int RPI = 32 / N16;                  // rows written per store instruction
        const int rsub = lane / N16, piece = lane % N16;
        auto stage_f32 = [&](const float (&x)[32]) {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < N16; ++j)
            st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), __float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]),
                         __float_as_uint(x[4 * j + 2]), __float_as_uint(x[4 * j + 3]));
          __syncwarp();
        };
        auto stage_f16 = [&](const float (&x)[32]) {   // relu'd values -> [hi (N16/2 pieces) | lo (N16/2 pieces)]
          __syncwarp();
#pragma unroll
          for (int j = 0; j < N16 / 2; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split2(x[8 * j + 2 * e], x[8 * j + 2 * e + 1], hi[e], lo[e]);
            st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(stg + lane * 128 + (((j + N16 / 2) ^ (lane & 7)) << 4), lo[0], lo[1], lo[2], lo[3]);
          }
          __syncwarp();
        };
        // rows: 0 = all 32 pixels of the warp (2 tile rows x 16), 1 = the 8 pool-window owners (rows 0,2,..,14)
        auto row_of = [&](int i, bool owners) { const int r = i * RPI + rsub; return owners ? 2 * r : r; };
        auto write_f32 = [&](float* base, int hh, int ww, bool owners) {     // base: [B][hh][ww][cout]
          const int n_rows = owners ? 8 : 32;
#pragma unroll 1
          for (int i = 0; i < (n_rows + RPI - 1) / RPI; ++i) {
            const int r = row_of(i, owners);
            if (i * RPI + rsub >= n_rows) continue;
            const int yy = y0 + q * 2 + (r >> 4), xx = x0 + (r & 15);
            const size_t px = owners ? ((size_t)b * hh + yy / 2) * ww + xx / 2 : ((size_t)b * hh + yy) * ww + xx;
            const uint4 d = ld_shared_v4(stg + r * 128 + ((piece ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(base + px * a.cout + n0 + piece * 4) = d;
          }
        };
        // mode 0: same pixel; 1: pooled pixel (owners only); 2: x2 upsample position (dy, dx)
        auto write_f16 = [&](__half* base, int pitch, int coff, int hh, int ww, int mode, int dy, int dx) {
          const bool owners = mode == 1;
          const int n_rows = owners ? 8 : 32;
          const int plane = piece / (N16 / 2), pc = piece % (N16 / 2);
#pragma unroll 2
          for (int i = 0; i < (n_rows + RPI - 1) / RPI; ++i) {
            const int r = row_of(i, owners);
            if (i * RPI + rsub >= n_rows) continue;
            const int yy = y0 + q * 2 + (r >> 4), xx = x0 + (r & 15);
            size_t px;
            if (mode == 0) px = ((size_t)b * hh + yy) * ww + xx;
            else if (mode == 1) px = ((size_t)b * hh + yy / 2) * ww + xx / 2;
            else px = ((size_t)b * hh + 2 * yy + dy) * ww + 2 * xx + dx;
            const uint4 d = ld_shared_v4(stg + r * 128 + ((piece ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(base + (px * 2 + plane) * pitch + coff + n0 + pc * 8) = d;
          }
        };
        if (a.feat) {
          if (a.feat_pooled) { stage_f32(pv); write_f32(a.feat, a.H / 2, a.W / 2, true); }
          else { stage_f32(v); write_f32(a.feat, a.H, a.W, false); }
        }
        // everything below stores relu'd fp16 hi/lo activations
#pragma unroll
        for (int j = 0; j < CH; ++j) { v[j] = fmaxf(v[j], 0.f); if (any_pool) pv[j] = fmaxf(pv[j], 0.f); }
        if (a.act_full || (a.act_up && !a.feat_pooled)) {
          stage_f16(v);
          if (a.act_full) write_f16(a.act_full, a.af_pitch, a.af_coff, a.H, a.W, 0, 0, 0);
          if (a.act_up && !a.feat_pooled) {
#pragma unroll 1
            for (int dd = 0; dd < 4; ++dd) write_f16(a.act_up, a.au_pitch, a.au_coff, 2 * a.H, 2 * a.W, 2, dd >> 1, dd & 1);
          }
        }
        if ((a.act_pool) || (a.act_up && a.feat_pooled)) {
          stage_f16(pv);
          if (a.act_pool) write_f16(a.act_pool, a.ap_pitch, a.ap_coff, a.H / 2, a.W / 2, 1, 0, 0);
          // pooled then upsampled x2: lands on this very pixel of the conv-resolution buffer
          if (a.act_up && a.feat_pooled) write_f16(a.act_up, a.au_pitch, a.au_coff, a.H, a.W, 0, 0, 0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_512(tmem_base); }
}

// ------------------------------------------------------------------------------ halo-tile variant (Cout = 64 layers)
// ncu on B200: the N = 64 layers (conv2, dec2.1, dec2.3) sat at 52-56 % tensor-pipe activity while moving the same
// 15.5 TB/s of TMA fill as the N = 128 layers at 91 %: with nine shifted copies of the activation tile per tile they
// are bound by L2 -> shared-memory bandwidth, not by the tensor pipe.  This variant loads the (16 + 2) x (8 + 2) pixel
// halo of a 16 x 8 pixel tile ONCE per 64-channel chunk (one TMA box of 18 rows x 10 pixels per plane) and the nine
// taps are nine shared-memory descriptors into it: start address + (ky * 10 + kx) * 128 B, stride between 8-row
// groups = one halo row = 1280 B.  The start is no longer 1024-byte aligned; the 128-byte swizzle is a function of
// the absolute shared-memory address on both the TMA and the UMMA side, so the descriptor's base_offset stays 0
// (measured: base_offset = kx gives wrong results, 0 passes every parity test).  A-operand fill per tile drops from
// 288 KB to 74 KB; the weights (16 KB per tap and chunk) stream through their own ring.
constexpr int kHaloTW = 8, kHaloTH = 16;                 // pixel tile: 16 rows x 8 columns = 128 = UMMA M
#ifndef HA_HALO_BOX_W
#define HA_HALO_BOX_W 10
#endif
constexpr int kHaloBoxW = HA_HALO_BOX_W, kHaloBoxH = kHaloTH + 2;   // TMA box: 18 rows x 10 pixels
constexpr int kHaloBoxBytes = kHaloBoxH * kHaloBoxW * 128;                          // what one TMA box delivers (22.5 KB)
constexpr int kHaloABytes = (kHaloBoxBytes + 1023) / 1024 * 1024;                   // plane slot: 1024-byte aligned for the swizzle
constexpr int kHaloAStages = 2, kHaloBStages = kHaloBoxW == 10 ? 6 : 3;
constexpr int kHaloBBytes = 2 * 64 * kBlockK * 2;        // [B_hi | B_lo] of one tap and chunk: 16 KB
constexpr int kHaloSmemBytes = kHaloAStages * 2 * kHaloABytes + kHaloBStages * kHaloBBytes + 1024 + 256 + 4 * 32 * 128;

__device__ __forceinline__ uint64_t umma_desc_sw128_halo(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((kHaloBoxW * 128) >> 4) << 32; // stride between 8-row groups: one halo row
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(kTcThreads, 1)
conv3x3_tc_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_bh,
                       const __grid_constant__ CUtensorMap tmap_bl, const TcConvArgs a) {
  constexpr int BLOCK_N = 64, CH = 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + kHaloAStages * 2 * kHaloABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kHaloBStages * kHaloBBytes);
  uint64_t* a_full = bars;                       // [2]  TMA -> MMA (halo tiles)
  uint64_t* a_empty = a_full + kHaloAStages;     // [2]
  uint64_t* b_full = a_empty + kHaloAStages;     // [3]  TMA -> MMA (weights of one tap)
  uint64_t* b_empty = b_full + kHaloBStages;     // [3]
  uint64_t* tmem_full = b_empty + kHaloBStages;  // [2]  MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* staging = smem_b + kHaloBStages * kHaloBBytes + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_bh); tma_prefetch_desc(&tmap_bl); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kHaloAStages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
    for (int s = 0; s < kHaloBStages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        int pt = tile;
        const int x0 = (pt % a.tiles_x) * kHaloTW; pt /= a.tiles_x;
        const int y0 = (pt % a.tiles_y) * kHaloTH; const int b = pt / a.tiles_y;
        for (int kc = 0; kc < a.n_kchunks; ++kc) {
          mbar_wait(a_empty + sa, pa ^ 1);
          uint8_t* st = smem + sa * 2 * kHaloABytes;
          mbar_expect_tx(a_full + sa, 2 * kHaloBoxBytes);
          tma_load_5d(st, &tmap_a, a_full + sa, kc * kBlockK, 0, x0 - 1, y0 - 1, b);
          tma_load_5d(st + kHaloABytes, &tmap_a, a_full + sa, kc * kBlockK, 1, x0 - 1, y0 - 1, b);
          if (++sa == kHaloAStages) { sa = 0; pa ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(b_empty + sb, pb ^ 1);
            uint8_t* sw = smem_b + sb * kHaloBBytes;
            mbar_expect_tx(b_full + sb, kHaloBBytes);
            tma_load_3d(sw, &tmap_bh, b_full + sb, kc * kBlockK, 0, tap);
            tma_load_3d(sw + kHaloBBytes / 2, &tmap_bl, b_full + sb, kc * kBlockK, 0, tap);
            if (++sb == kHaloBStages) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_f16(BLOCK_N);             // A_lo x B_hi
    constexpr uint32_t idesc2 = umma_idesc_f16(2 * BLOCK_N);        // A_hi x [B_hi ; B_lo]
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0; int t = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++t) {
      const int as = t & 1; const uint32_t aphase = (t >> 1) & 1;
      mbar_wait(tmem_empty + as, aphase ^ 1);
      tc_fence_after();
      const uint32_t acc0 = tmem_base + as * 256, acc1 = acc0 + BLOCK_N;
      for (int kc = 0; kc < a.n_kchunks; ++kc) {
        mbar_wait(a_full + sa, pa);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(smem) + sa * 2 * kHaloABytes, a_lo = a_hi + kHaloABytes;
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(b_full + sb, pb);
          tc_fence_after();
          if (lane == 0) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const uint32_t off = (ky * kHaloBoxW + kx) * 128;
            const uint64_t da = umma_desc_sw128_halo(a_hi + off), dl = umma_desc_sw128_halo(a_lo + off);
            const uint64_t db = umma_desc_sw128(smem_u32(smem_b) + sb * kHaloBBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              umma_f16(acc0, da + 2 * k, db + 2 * k, idesc2, (kc | tap | k) != 0);
              umma_f16(acc1, dl + 2 * k, db + 2 * k, idesc, 1);
            }
            umma_commit(b_empty + sb);
            if (tap == 8) {
              umma_commit(a_empty + sa);
              if (kc == a.n_kchunks - 1) umma_commit(tmem_full + as);
            }
          }
          __syncwarp();
          if (++sb == kHaloBStages) { sb = 0; pb ^= 1; }
        }
        if (++sa == kHaloAStages) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (tile = 16 rows x 8 columns: TMEM lane m = pixel (m / 8, m % 8)) =====================
    const int q = warp & 3;                               // TMEM lane quarter = tile rows [4 q, 4 q + 4)
    const uint32_t stg = smem_u32(staging) + (warp - 4) * (32 * 128);
    const bool any_pool = a.act_pool != nullptr || a.feat_pooled;
    int t = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++t) {
      const int as = t & 1; const uint32_t aphase = (t >> 1) & 1;
      int pt = tile;
      const int x0 = (pt % a.tiles_x) * kHaloTW; pt /= a.tiles_x;
      const int y0 = (pt % a.tiles_y) * kHaloTH; const int b = pt / a.tiles_y;
      mbar_wait(tmem_full + as, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += CH) {
        uint32_t r0[32], r1[32];
        tmem_ld_x32(taddr + c0, r0); tmem_ld_x32(taddr + BLOCK_N + c0, r1);
        tmem_ld_wait();
        const int n0 = c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          const float4 bz = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float bb[4] = {bz.x, bz.y, bz.z, bz.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) v[j + e] = fmaf(__uint_as_float(r1[j + e]), kLoInvScale, __uint_as_float(r0[j + e])) + bb[e];
        }
        float pv[32];
        if (any_pool) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));          // x ^ 1
            pv[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));                   // y ^ 1
          }
        }
        constexpr int N16 = CH / 4, RPI = 32 / N16;
        const int rsub = lane / N16, piece = lane % N16;
        auto stage_f32 = [&](const float (&x)[32]) {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < N16; ++j)
            st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), __float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]),
                         __float_as_uint(x[4 * j + 2]), __float_as_uint(x[4 * j + 3]));
          __syncwarp();
        };
        auto stage_f16 = [&](const float (&x)[32]) {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < N16 / 2; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split2(x[8 * j + 2 * e], x[8 * j + 2 * e + 1], hi[e], lo[e]);
            st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(stg + lane * 128 + (((j + N16 / 2) ^ (lane & 7)) << 4), lo[0], lo[1], lo[2], lo[3]);
          }
          __syncwarp();
        };
        // staged row r = lane r of this warp = pixel (4 q + r / 8, r % 8); the 8 pool-window owners are the lanes
        // with even x and even y: r = 2 (k % 4) + 16 (k / 4), k = 0..7
        auto row_of = [&](int i, bool owners) { const int k = i * RPI + rsub; return owners ? 2 * (k & 3) + 16 * (k >> 2) : k; };
        auto write_f32 = [&](float* base, int hh, int ww, bool owners) {
          const int n_rows = owners ? 8 : 32;
#pragma unroll 1
          for (int i = 0; i < (n_rows + RPI - 1) / RPI; ++i) {
            if (i * RPI + rsub >= n_rows) continue;
            const int r = row_of(i, owners);
            const int yy = y0 + q * 4 + (r >> 3), xx = x0 + (r & 7);
            const size_t px = owners ? ((size_t)b * hh + yy / 2) * ww + xx / 2 : ((size_t)b * hh + yy) * ww + xx;
            const uint4 d = ld_shared_v4(stg + r * 128 + ((piece ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(base + px * a.cout + n0 + piece * 4) = d;
          }
        };
        auto write_f16 = [&](__half* base, int pitch, int coff, int hh, int ww, int mode, int dy, int dx) {
          const bool owners = mode == 1;
          const int n_rows = owners ? 8 : 32;
          const int plane = piece / (N16 / 2), pc = piece % (N16 / 2);
#pragma unroll 2
          for (int i = 0; i < (n_rows + RPI - 1) / RPI; ++i) {
            if (i * RPI + rsub >= n_rows) continue;
            const int r = row_of(i, owners);
            const int yy = y0 + q * 4 + (r >> 3), xx = x0 + (r & 7);
            size_t px;
            if (mode == 0) px = ((size_t)b * hh + yy) * ww + xx;
            else if (mode == 1) px = ((size_t)b * hh + yy / 2) * ww + xx / 2;
            else px = ((size_t)b * hh + 2 * yy + dy) * ww + 2 * xx + dx;
            const uint4 d = ld_shared_v4(stg + r * 128 + ((piece ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(base + (px * 2 + plane) * pitch + coff + n0 + pc * 8) = d;
          }
        };
        if (a.feat) {
          if (a.feat_pooled) { stage_f32(pv); write_f32(a.feat, a.H / 2, a.W / 2, true); }
          else { stage_f32(v); write_f32(a.feat, a.H, a.W, false); }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) { v[j] = fmaxf(v[j], 0.f); if (any_pool) pv[j] = fmaxf(pv[j], 0.f); }
        if (a.act_full || (a.act_up && !a.feat_pooled)) {
          stage_f16(v);
          if (a.act_full) write_f16(a.act_full, a.af_pitch, a.af_coff, a.H, a.W, 0, 0, 0);
          if (a.act_up && !a.feat_pooled) {
#pragma unroll 1
            for (int dd = 0; dd < 4; ++dd) write_f16(a.act_up, a.au_pitch, a.au_coff, 2 * a.H, 2 * a.W, 2, dd >> 1, dd & 1);
          }
        }
        if ((a.act_pool) || (a.act_up && a.feat_pooled)) {
          stage_f16(pv);
          if (a.act_pool) write_f16(a.act_pool, a.ap_pitch, a.ap_coff, a.H / 2, a.W / 2, 1, 0, 0);
          if (a.act_up && a.feat_pooled) write_f16(a.act_up, a.au_pitch, a.au_coff, a.H, a.W, 0, 0, 0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_512(tmem_base); }
}

// ------------------------------------------------------------------------------ CTA-pair halo-tile variant (Cout >= 64)
// Round-1 ncu (profiles/r01c_conv_full.csv) showed every f16x3 layer bound by UMMA operand reads from shared memory
// (~69 B/clk per SM whatever the layer): 20 KB per K=16 step on the N = 128 layers (290 cycles against a 192-cycle MMA
// floor), and the nine-box kernels additionally close to the L2 -> shared-memory fill limit.  This kernel attacks both:
//  * `tcgen05.mma.cta_group::2`: a pair of CTAs (one thread-block cluster, the two SMs of a TPC) issues ONE M = 256
//    MMA per step; each CTA supplies its own 128-pixel A tile and only HALF of the B operand, so the weight-side
//    operand bytes per SM halve: 14 KB per K=16 step instead of 20 (N = 128), 11 instead of 14 (N = 64);
//  * the halo tile of the kernel above (one 18 x 10-pixel box per 64-channel chunk, nine descriptors into it), so the
//    activation fill drops 9x and the weights dominate the fill.
// f16x3 in pair form.  MMA1 = [A_hi(P); A_hi(Q)] x [B_hi | B_lo] (N = 2*BLOCK_N): the pair splits B by columns, so CTA 0
// holds B_hi and CTA 1 holds B_lo (region `main`, BLOCK_N rows each); accumulator columns [0, BLOCK_N) = hi*hi and
// [BLOCK_N, 2*BLOCK_N) = hi*lo in BOTH CTAs' TMEM (each for its own pixel tile).  MMA2 = [A_lo(P); A_lo(Q)] x B_hi
// (N = BLOCK_N) accumulates lo*hi onto the second half; its B operand is split by columns too, so CTA r holds rows
// [r*BLOCK_N/2, (r+1)*BLOCK_N/2) of B_hi at the same shared-memory offset (region `extra`; CTA 0 loads its half of B_hi a
// second time, from L2, to keep the two CTAs' layouts identical as the single descriptor requires).
// Synchronisation: the leader (cluster rank 0) issues every MMA.  TMA loads of the peer complete on the LEADER's full
// barriers (`cp.async.bulk.tensor...cta_group::2` with the barrier address mapped to rank 0); `tcgen05.commit` with a
// multicast mask releases shared-memory slots / publishes accumulators in both CTAs; the peer's epilogue warps arrive
// on the leader's tmem_empty barrier through the cluster address space.
template <int BLOCK_N>
struct Halo2Cfg {
  static constexpr int kBMain = BLOCK_N * kBlockK * 2;            // BLOCK_N rows x 128 B
  static constexpr int kBExtra = kBMain / 2;
  static constexpr int kBStage = kBMain + kBExtra;                // 24 KB (N = 128) / 12 KB (N = 64)
  static constexpr int kAStages = 2;
  static constexpr int kBStages = BLOCK_N == 128 ? 4 : 9;     // N <= 64: nine stages = every tap of one 64-channel chunk
  static constexpr int kSmemBytes = kAStages * 2 * kHaloABytes + kBStages * kBStage + 1024 + 256 + 4 * 32 * 128;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma2_load_5d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                             int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tmem_alloc2_512(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2_512(uint32_t addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}
// kind::f16 instruction descriptor for the pair: fp16 x fp16 -> fp32, A and B K-major, M = 256 (128 rows per CTA)
__host__ __device__ constexpr uint32_t umma2_idesc_f16(int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// RESIDENT (Cin = 64, N = 64: conv2 and dec2.3): the layer's whole weight set — nine taps x 12 KB per CTA — is loaded
// into the nine B stages ONCE per CTA and stays there; only the halo tiles stream.  The single-CTA kernels re-read
// 144 KB of weights from L2 for every 128-pixel tile (9.4 of the 12.9 GB of L2 -> SM traffic of conv2 at B = 32), which
// held these layers at 51 % tensor-pipe activity; a micro-benchmark (tools/umma_bench.cu) shows the pipe itself
// sustains the (N = 128, N = 64) instruction pair in 112 cycles against the 188 the streamed kernels achieve.
template <int BLOCK_N, bool RESIDENT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
conv3x3_tc_halo2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_bh,
                        const __grid_constant__ CUtensorMap tmap_bl, const __grid_constant__ CUtensorMap tmap_bh_half,
                        const TcConvArgs a) {
  using Cfg = Halo2Cfg<BLOCK_N>;
  constexpr int CH = 32, AST = Cfg::kAStages, BST = Cfg::kBStages;
  constexpr bool TWO = BLOCK_N <= 64;            // two MMA-issuing warps, three accumulators per stage (see the issuers)
  constexpr uint32_t kIssuers = TWO ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + AST * 2 * kHaloABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + BST * Cfg::kBStage);
  uint64_t* a_full = bars;                       // [AST]  TMA (both CTAs) -> leader's MMA warp
  uint64_t* a_empty = a_full + AST;              // [AST]  MMA commit (multicast) -> both producers
  uint64_t* b_full = a_empty + AST;              // [BST]
  uint64_t* b_empty = b_full + BST;              // [BST]
  uint64_t* tmem_full = b_empty + BST;           // [2]    MMA commit (multicast) -> both epilogues
  uint64_t* tmem_empty = tmem_full + 2;          // [2]    both epilogues (8 warps) -> leader's MMA warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* staging = smem_b + BST * Cfg::kBStage + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_bh); tma_prefetch_desc(&tmap_bl); tma_prefetch_desc(&tmap_bh_half);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < AST; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, kIssuers); }
    for (int s = 0; s < BST; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, kIssuers); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, kIssuers); mbar_init(tmem_empty + s, 8); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2_512(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                           // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pair-tile u -> (n tile, pixel-tile pair); CTA r of the pair owns pixel tile 2 * pp + r
  const int n_pairs = a.n_tiles / 2;            // a.n_tiles = pixel tiles * tiles_n, pixel tiles even (host checks)
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;
  auto decode = [&](int u, int& n_idx, int& x0, int& y0, int& b) {
    n_idx = u % a.tiles_n;
    int pt = (u / a.tiles_n) * 2 + (int)rank;
    x0 = (pt % a.tiles_x) * kHaloTW; pt /= a.tiles_x;
    y0 = (pt % a.tiles_y) * kHaloTH; b = pt / a.tiles_y;
  };

  if (warp == 0) {
    // ===================== TMA producer (one per CTA; completes on the leader's barriers) =====================
    if (lane == 0) {
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      const CUtensorMap* map_main = leader ? &tmap_bh : &tmap_bl;
      if (RESIDENT) {                           // one chunk, one N tile: tap t lives in stage t for the whole kernel
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t sw = smem_u32(smem_b) + tap * Cfg::kBStage;
          const uint32_t bar_b = mapa_rank(smem_u32(b_full + tap), 0);
          if (leader) mbar_expect_tx(b_full + tap, 2 * Cfg::kBStage);
          tma2_load_3d(sw, map_main, bar_b, 0, 0, tap);
          tma2_load_3d(sw + Cfg::kBMain, &tmap_bh_half, bar_b, 0, (int)rank * (BLOCK_N / 2), tap);
        }
      }
      for (int u = pair0; u < n_pairs; u += pair_stride) {
        int n_idx, x0, y0, b;
        decode(u, n_idx, x0, y0, b);
        for (int kc = 0; kc < a.n_kchunks; ++kc) {
          mbar_wait(a_empty + sa, pa ^ 1);
          const uint32_t st = smem_u32(smem) + sa * 2 * kHaloABytes;
          const uint32_t bar_a = mapa_rank(smem_u32(a_full + sa), 0);
          if (leader) mbar_expect_tx(a_full + sa, 2 * 2 * kHaloBoxBytes);        // both CTAs' halo tiles
          tma2_load_5d(st, &tmap_a, bar_a, kc * kBlockK, 0, x0 - 1, y0 - 1, b);
          tma2_load_5d(st + kHaloABytes, &tmap_a, bar_a, kc * kBlockK, 1, x0 - 1, y0 - 1, b);
          if (++sa == AST) { sa = 0; pa ^= 1; }
          if (RESIDENT) continue;
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(b_empty + sb, pb ^ 1);
            const uint32_t sw = smem_u32(smem_b) + sb * Cfg::kBStage;
            const uint32_t bar_b = mapa_rank(smem_u32(b_full + sb), 0);
            if (leader) mbar_expect_tx(b_full + sb, 2 * Cfg::kBStage);
            tma2_load_3d(sw, map_main, bar_b, kc * kBlockK, n_idx * BLOCK_N, tap);
            tma2_load_3d(sw + Cfg::kBMain, &tmap_bh_half, bar_b, kc * kBlockK, n_idx * BLOCK_N + (int)rank * (BLOCK_N / 2), tap);
            if (++sb == BST) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if ((warp == 1 || (TWO && warp == 3)) && leader) {
    // ===================== MMA issuers (leader CTA only) =====================
    // ncu (profiles/r02c): with descriptors built per MMA inside the single-lane region every UTCHMMA cost ~25
    // instructions (R2UR moves, address arithmetic, the compiler's ELECT loop): the issuing warp needed ~800 cycles per
    // tap and WAS the critical path — the Cout = 64 layers sat at 49 % tensor-pipe activity with no barrier ever
    // blocking the issuer, and the N = 128 layers (768 cycles of math per tap) barely hid it.  Now: descriptor fields
    // are warp-uniform values prepared outside the single-lane region, only the low 32 descriptor bits change per MMA,
    // the MMAs of a tap go out in one asm block, and for N = 64 (96 cycles of math per K = 16 step) the two MMA streams
    // are issued by TWO warps: warp 1 issues A_hi x [B_hi | B_lo] into columns [0, 2N), warp 3 issues A_lo x B_hi into
    // its own columns [2N, 3N) (the epilogue adds the two correction accumulators), so neither stream orders the other.
    constexpr uint32_t idesc1 = umma2_idesc_f16(2 * BLOCK_N);       // [A_hi; A_hi'] x [B_hi | B_lo]
    constexpr uint32_t idesc2 = umma2_idesc_f16(BLOCK_N);           // [A_lo; A_lo'] x B_hi
    constexpr uint32_t kDescHiA = (uint32_t)((kHaloBoxW * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO = one halo row | version 1 | SWIZZLE_128B
    constexpr uint32_t kDescHiB = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const bool do1 = warp == 1, do2 = TWO ? warp == 3 : warp == 1;  // which MMA stream(s) this warp issues
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a_field0 = __shfl_sync(0xffffffffu, ((smem_u32(smem) >> 4) & 0x3FFF) | (1u << 16), 0);     // start address | LBO
    const uint32_t b_field0 = __shfl_sync(0xffffffffu, ((smem_u32(smem_b) >> 4) & 0x3FFF) | (1u << 16), 0);
    const uint32_t bar_b_empty = __shfl_sync(0xffffffffu, smem_u32(b_empty), 0);
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0; int t = 0;
    for (int u = pair0; u < n_pairs; u += pair_stride, ++t) {
      const int as = t & 1; const uint32_t aphase = (t >> 1) & 1;
      mbar_wait(tmem_empty + as, aphase ^ 1);
      tc_fence_after();
      const uint32_t acc0 = tmem_u + as * 256, acc1 = acc0 + (TWO ? 2 * BLOCK_N : BLOCK_N);
      for (int kc = 0; kc < a.n_kchunks; ++kc) {
        mbar_wait(a_full + sa, pa);
        tc_fence_after();
        const uint32_t a_hi = a_field0 + (uint32_t)(sa * 2 * kHaloABytes >> 4), a_lo = a_hi + (kHaloABytes >> 4);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if (!RESIDENT || t == 0) {             // resident weights: waited for once, on the first tile
            mbar_wait(b_full + sb, RESIDENT ? 0u : pb);
            tc_fence_after();
          }
          constexpr int kTapOff[9] = {0, 1, 2, kHaloBoxW, kHaloBoxW + 1, kHaloBoxW + 2, 2 * kHaloBoxW, 2 * kHaloBoxW + 1, 2 * kHaloBoxW + 2};
          const uint32_t da = a_hi + kTapOff[tap] * (128 >> 4), dl = a_lo + kTapOff[tap] * (128 >> 4);
          const uint32_t db = b_field0 + (uint32_t)(sb * Cfg::kBStage >> 4), dx = db + (Cfg::kBMain >> 4);
          const uint32_t first = (kc | tap) != 0 ? 1u : 0u;
          if (lane == 0) {
            // the four K = 16 steps of this tap per stream, one asm block each: descriptor high words are constants
#define HA_MMA4(ACC, DA, DB, HA_, HB_, IDESC, FIRST)                                                                     \
            asm volatile(                                                                                                \
                "{\n"                                                                                                    \
                ".reg .pred p0;\n"                                                                                       \
                ".reg .b64 da, db;\n"                                                                                    \
                ".reg .b32 t;\n"                                                                                         \
                "setp.ne.b32 p0, %5, 0;\n"                                                                               \
                "mov.b64 da, {%1, %3}; mov.b64 db, {%2, %4};\n"                                                          \
                "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, p0;\n"                                             \
                "add.u32 t, %1, 2; mov.b64 da, {t, %3}; add.u32 t, %2, 2; mov.b64 db, {t, %4};\n"                        \
                "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, 1;\n"                                              \
                "add.u32 t, %1, 4; mov.b64 da, {t, %3}; add.u32 t, %2, 4; mov.b64 db, {t, %4};\n"                        \
                "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, 1;\n"                                              \
                "add.u32 t, %1, 6; mov.b64 da, {t, %3}; add.u32 t, %2, 6; mov.b64 db, {t, %4};\n"                        \
                "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %6, 1;\n"                                              \
                "}\n" ::"r"(ACC), "r"(DA), "r"(DB), "n"(HA_), "n"(HB_), "r"(FIRST), "n"(IDESC) : "memory")
            if (do1) HA_MMA4(acc0, da, db, kDescHiA, kDescHiB, idesc1, first);
            // one issuer: MMA1's first instruction has zeroed both accumulators; two issuers: the streams are independent
            if (do2) HA_MMA4(acc1, dl, dx, kDescHiA, kDescHiB, idesc2, TWO ? first : 1u);
#undef HA_MMA4
            if (!RESIDENT)
              asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                           ::"r"(bar_b_empty + sb * 8), "h"((uint16_t)3) : "memory");
            if (tap == 8) {
              umma2_commit_mc(a_empty + sa, 3);
              if (kc == a.n_kchunks - 1) umma2_commit_mc(tmem_full + as, 3);
            }
          }
          __syncwarp();
          if (++sb == BST) { sb = 0; pb ^= 1; }
        }
        if (++sa == AST) { sa = 0; pa ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (tile = 16 rows x 8 columns: TMEM lane m = pixel (m / 8, m % 8)) =====================
    const int q = warp & 3;                               // TMEM lane quarter = tile rows [4 q, 4 q + 4)
    const uint32_t stg = smem_u32(staging) + (warp - 4) * (32 * 128);
    const bool any_pool = a.act_pool != nullptr || a.feat_pooled;
    const uint32_t tmem_empty_leader = mapa_rank(smem_u32(tmem_empty), 0);
    int t = 0;
    for (int u = pair0; u < n_pairs; u += pair_stride, ++t) {
      const int as = t & 1; const uint32_t aphase = (t >> 1) & 1;
      int n_idx, x0, y0, b;
      decode(u, n_idx, x0, y0, b);
      mbar_wait(tmem_full + as, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += CH) {
        uint32_t r0[32], r1[32];
        tmem_ld_x32(taddr + c0, r0); tmem_ld_x32(taddr + BLOCK_N + c0, r1);
        tmem_ld_wait();
        if (TWO) {                               // hi*lo and lo*hi live in separate accumulators: add them first
          uint32_t r2[32];
          tmem_ld_x32(taddr + 2 * BLOCK_N + c0, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __uint_as_float(r2[j]));
        }
        const int n0 = n_idx * BLOCK_N + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          const float4 bz = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float bb[4] = {bz.x, bz.y, bz.z, bz.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) v[j + e] = fmaf(__uint_as_float(r1[j + e]), kLoInvScale, __uint_as_float(r0[j + e])) + bb[e];
        }
        float pv[32];
        if (any_pool) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));          // x ^ 1
            pv[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));                   // y ^ 1
          }
        }
        constexpr int N16 = CH / 4, RPI = 32 / N16;
        const int rsub = lane / N16, piece = lane % N16;
        auto stage_f32 = [&](const float (&x)[32]) {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < N16; ++j)
            st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), __float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]),
                         __float_as_uint(x[4 * j + 2]), __float_as_uint(x[4 * j + 3]));
          __syncwarp();
        };
        auto stage_f16 = [&](const float (&x)[32]) {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < N16 / 2; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split2(x[8 * j + 2 * e], x[8 * j + 2 * e + 1], hi[e], lo[e]);
            st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), hi[0], hi[1], hi[2], hi[3]);
            st_shared_v4(stg + lane * 128 + (((j + N16 / 2) ^ (lane & 7)) << 4), lo[0], lo[1], lo[2], lo[3]);
          }
          __syncwarp();
        };
        auto row_of = [&](int i, bool owners) { const int k = i * RPI + rsub; return owners ? 2 * (k & 3) + 16 * (k >> 2) : k; };
        auto write_f32 = [&](float* base, int hh, int ww, bool owners) {
          const int n_rows = owners ? 8 : 32;
#pragma unroll 1
          for (int i = 0; i < (n_rows + RPI - 1) / RPI; ++i) {
            if (i * RPI + rsub >= n_rows) continue;
            const int r = row_of(i, owners);
            const int yy = y0 + q * 4 + (r >> 3), xx = x0 + (r & 7);
            const size_t px = owners ? ((size_t)b * hh + yy / 2) * ww + xx / 2 : ((size_t)b * hh + yy) * ww + xx;
            const uint4 d = ld_shared_v4(stg + r * 128 + ((piece ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(base + px * a.cout + n0 + piece * 4) = d;
          }
        };
        auto write_f16 = [&](__half* base, int pitch, int coff, int hh, int ww, int mode, int dy, int dx) {
          const bool owners = mode == 1;
          const int n_rows = owners ? 8 : 32;
          const int plane = piece / (N16 / 2), pc = piece % (N16 / 2);
#pragma unroll 2
          for (int i = 0; i < (n_rows + RPI - 1) / RPI; ++i) {
            if (i * RPI + rsub >= n_rows) continue;
            const int r = row_of(i, owners);
            const int yy = y0 + q * 4 + (r >> 3), xx = x0 + (r & 7);
            size_t px;
            if (mode == 0) px = ((size_t)b * hh + yy) * ww + xx;
            else if (mode == 1) px = ((size_t)b * hh + yy / 2) * ww + xx / 2;
            else px = ((size_t)b * hh + 2 * yy + dy) * ww + 2 * xx + dx;
            const uint4 d = ld_shared_v4(stg + r * 128 + ((piece ^ (r & 7)) << 4));
            *reinterpret_cast<uint4*>(base + (px * 2 + plane) * pitch + coff + n0 + pc * 8) = d;
          }
        };
        if (a.feat) {
          if (a.feat_pooled) { stage_f32(pv); write_f32(a.feat, a.H / 2, a.W / 2, true); }
          else { stage_f32(v); write_f32(a.feat, a.H, a.W, false); }
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) { v[j] = fmaxf(v[j], 0.f); if (any_pool) pv[j] = fmaxf(pv[j], 0.f); }
        if (a.act_full || (a.act_up && !a.feat_pooled)) {
          stage_f16(v);
          if (a.act_full) write_f16(a.act_full, a.af_pitch, a.af_coff, a.H, a.W, 0, 0, 0);
          if (a.act_up && !a.feat_pooled) {
#pragma unroll 1
            for (int dd = 0; dd < 4; ++dd) write_f16(a.act_up, a.au_pitch, a.au_coff, 2 * a.H, 2 * a.W, 2, dd >> 1, dd & 1);
          }
        }
        if ((a.act_pool) || (a.act_up && a.feat_pooled)) {
          stage_f16(pv);
          if (a.act_pool) write_f16(a.act_pool, a.ap_pitch, a.ap_coff, a.H / 2, a.W / 2, 1, 0, 0);
          if (a.act_up && a.feat_pooled) write_f16(a.act_up, a.au_pitch, a.au_coff, a.H, a.W, 0, 0, 0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                     // this accumulator stage is drained in this CTA
        if (leader) mbar_arrive(tmem_empty + as);
        else mbar_arrive_cluster(tmem_empty_leader + as * 8);
      }
    }
  }
  __syncwarp();                                 // lanes of the single-thread roles rejoin their warps: the cluster barrier is .aligned
  tc_fence_before();
  cluster_sync_all();                           // no CTA of the pair leaves (or frees TMEM) while the other still computes
  if (warp == 2) { tc_fence_after(); tmem_dealloc2_512(tmem_base); }
}

// ------------------------------------------------------------------------------ conv0 + helpers
// conv0 (3 -> 64, K = 27) is 0.7 % of the FLOPs but writes the largest activation of the network (2.1 GB of hi/lo
// planes per branch at B = 32), so it is built around its two real limits:
//  * FP32 pipe: a lane owns 8 horizontally adjacent pixels x 8 output channels and accumulates with packed FFMA2
//    (one 128-bit weight broadcast from shared memory feeds 32 FMAs; the scalar-FFMA version ran at half the rate);
//  * stores: a pixel's output is one contiguous 256-byte record [hi 64 | lo 64]; the eight warps of a CTA (= the eight
//    channel blocks of the same 256 pixels) park their 16-byte pieces in an XOR-swizzled staging tile and the CTA
//    writes whole records, 512 contiguous bytes per store instruction (the direct version touched 32 lines per
//    store and ran at 1.6 TB/s).
// (A tensor-core variant — im2col channels + 1x1 tcgen05 conv — measured 1.39 ms against 1.30-1.40 ms for the first
// CUDA-core kernel at B = 32: with one k-block per tile it is bound by the same epilogue.)
constexpr int kC0TW = 64, kC0TH = 16;     // CTA tile: 64 x 16 pixels, 256 threads, four passes of 64 x 4 pixels
constexpr int kC0Pitch = kC0TW + 4;       // input row pitch in floats: 272 B keeps every 8-pixel window 16-byte aligned
constexpr int kC0Smem = 64 * 4 * 256 /*staging*/ + 27 * 64 * 4 + 64 * 4 + 3 * (kC0TH + 2) * kC0Pitch * 4;

typedef unsigned long long c0_f32x2;
__device__ __forceinline__ c0_f32x2 c0_pack(float lo, float hi) { c0_f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void c0_unpack(c0_f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void c0_fma(c0_f32x2& acc, c0_f32x2 a, c0_f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

__global__ void __launch_bounds__(256, 2) conv0_kernel(const float* __restrict__ img, const float* __restrict__ w /*[9][3][64]*/,
                                                    const float* __restrict__ bias, __half* __restrict__ out, int H, int W) {
  extern __shared__ __align__(128) uint8_t c0_smem[];
  uint8_t* stage = c0_smem;                                                      // [256 pixels][256 B], 16-byte pieces swizzled
  float* w_s = reinterpret_cast<float*>(c0_smem + 64 * 4 * 256);                 // [27][64]
  float* b_s = w_s + 27 * 64;                                                    // [64]
  float* in_s = b_s + 64;                                                        // [3][18][kC0Pitch] halo tile
  const int b = blockIdx.z, tx0 = blockIdx.x * kC0TW, ty0 = blockIdx.y * kC0TH;
  for (int i = threadIdx.x; i < 27 * 64; i += 256) w_s[i] = w[i];
  if (threadIdx.x < 64) b_s[threadIdx.x] = bias[threadIdx.x];
  // halo tile: in_s[c][yy][xx] = img(c, ty0 + yy - 1, tx0 + xx - 1), zero outside the image (padding = 1).  All of a
  // thread's loads are issued before the first store, so their latencies overlap (14 loads per thread).
  {
    constexpr int kHalo = 3 * (kC0TH + 2) * (kC0TW + 2), kPer = (kHalo + 255) / 256;
    float v[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = threadIdx.x + k * 256;
      const int xx = i % (kC0TW + 2), yy = (i / (kC0TW + 2)) % (kC0TH + 2), c = i / ((kC0TW + 2) * (kC0TH + 2));
      const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
      v[k] = (i < kHalo && gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + (((size_t)b * 3 + c) * H + gy) * W + gx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int i = threadIdx.x + k * 256;
      const int xx = i % (kC0TW + 2), yy = (i / (kC0TW + 2)) % (kC0TH + 2), c = i / ((kC0TW + 2) * (kC0TH + 2));
      if (i < kHalo) in_s[(c * (kC0TH + 2) + yy) * kC0Pitch + xx] = v[k];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, cb = threadIdx.x >> 5;      // channel block: channels [8 cb, 8 cb + 8)
  const int r = lane >> 3, x0 = (lane & 7) * 8;                  // this lane's 8 pixels: row r of the pass, columns [x0, x0 + 8)
  const uint32_t stage_u = smem_u32(stage);
#pragma unroll 1
  for (int pass = 0; pass < kC0TH / 4; ++pass) {
    const int ly = pass * 4 + r;
    c0_f32x2 acc[8][4];                                          // [pixel][channel pair]
    {
      const float4 b0 = *reinterpret_cast<const float4*>(b_s + cb * 8), b1 = *reinterpret_cast<const float4*>(b_s + cb * 8 + 4);
      const c0_f32x2 bp[4] = {c0_pack(b0.x, b0.y), c0_pack(b0.z, b0.w), c0_pack(b1.x, b1.y), c0_pack(b1.z, b1.w)};
#pragma unroll
      for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[p][j] = bp[j];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float win[3][12];                                          // rows ly-1..ly+1, columns x0-1..x0+8 (+2 unused)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* row = in_s + (c * (kC0TH + 2) + ly + ky) * kC0Pitch + x0;
        const float4 v0 = *reinterpret_cast<const float4*>(row), v1 = *reinterpret_cast<const float4*>(row + 4);
        const float2 v2 = *reinterpret_cast<const float2*>(row + 8);
        win[ky][0] = v0.x; win[ky][1] = v0.y; win[ky][2] = v0.z; win[ky][3] = v0.w;
        win[ky][4] = v1.x; win[ky][5] = v1.y; win[ky][6] = v1.z; win[ky][7] = v1.w;
        win[ky][8] = v2.x; win[ky][9] = v2.y;
      }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float* wp = w_s + (tap * 3 + c) * 64 + cb * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
        const c0_f32x2 wq[4] = {c0_pack(w0.x, w0.y), c0_pack(w0.z, w0.w), c0_pack(w1.x, w1.y), c0_pack(w1.z, w1.w)};
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const float x = win[tap / 3][p + tap % 3];
          const c0_f32x2 xx = c0_pack(x, x);
#pragma unroll
          for (int j = 0; j < 4; ++j) c0_fma(acc[p][j], wq[j], xx);
        }
      }
    }
    // ReLU, hi/lo split, 16-byte pieces into the staging tile: pixel record = [hi 8 pieces | lo 8 pieces], piece index
    // XOR-ed with (pixel / 8) & 7 so that the 8 lanes of a row hit 8 different bank groups
    if (pass) __syncthreads();                                   // the previous pass has been written out
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v0, v1;
        c0_unpack(acc[p][j], v0, v1);
        split2(fmaxf(v0, 0.f), fmaxf(v1, 0.f), hi[j], lo[j]);
      }
      const int px = r * 64 + x0 + p;
      const uint32_t rec = stage_u + px * 256 + ((cb ^ ((px >> 3) & 7)) << 4);
      st_shared_v4(rec, hi[0], hi[1], hi[2], hi[3]);
      st_shared_v4(rec + 128, lo[0], lo[1], lo[2], lo[3]);
    }
    __syncthreads();
    // write-out: 256 records of 256 B; a warp instruction stores two whole records (512 contiguous bytes).  Warp cb
    // owns pixels [32 cb, 32 cb + 32) of the pass = half a tile row, so its 16 stores are 512 bytes apart.
    {
      const int half_px = lane >> 4, piece = lane & 15;          // piece 0-7: hi, 8-15: lo
      __half* gp = out + (((size_t)b * H + ty0 + pass * 4 + (cb >> 1)) * W + tx0 + (cb & 1) * 32 + half_px) * 128 + piece * 8;
      const uint32_t sp = stage_u + (cb * 32 + half_px) * 256 + (piece >> 3) * 128;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int key = (cb * 4 + (i >> 2)) & 7;                 // ((32 cb + 2 i + half_px) >> 3) & 7
        const uint4 d = ld_shared_v4(sp + i * 512 + (((piece & 7) ^ key) << 4));
        *reinterpret_cast<uint4*>(gp + i * 256) = d;
      }
    }
  }
}

// ------------------------------------------------------------------------------ conv0 on the tensor pipe
// conv0_kernel above is FP32-pipe bound (27 x 64 FMAs per pixel: 0.88 ms per 512 x 512 x 32 branch at 53 % pipe
// utilisation) although its real floor is the 2.1 GB it writes (0.35 ms).  Here the contraction runs on tcgen05 like
// every other layer: K = 27 im2col values per pixel (padded to 32 = two K = 16 steps), f16x3 split of image and
// weights, M = 128 consecutive pixels of one image row per tile.  Nothing is staged in global memory:
//   warps 0-3 (128 threads, one pixel each, one warp per SM sub-partition) read the 3 x 3 x 3 neighbourhood straight
//     from the fp32 image, split it into fp16 hi / lo and write the pixel's K-major row into the SWIZZLE_128B operand
//     tiles (generic stores + fence.proxy.async, then an arrive on the tile's `full` barrier).  The image loads of the
//     tiles one and two ahead are in flight meanwhile (three register buffers): under the kernel's own 3 TB/s of
//     writes a load takes longer than a tile (ncu: with one tile of lead the epilogue warps starved, 695-870 us).
//   warp 12 allocates TMEM and issues A_hi x [B_hi | B_lo] (N = 128) and A_lo x B_hi (N = 64) for the two K steps; the
//     16 KB weight tile is built once per CTA;
//   warps 4-11 are the epilogue, two per TMEM lane quarter (channels 0-31 and 32-63: one warp per quarter was the
//     critical path, ~600 dependent instructions per tile): the three accumulators -> bias, ReLU, split -> the pixel's
//     [hi 64 B | lo 64 B] half-record in a warp-private XOR-swizzled staging tile -> global, 64-byte segments that the
//     sibling warp completes to whole 128-byte lines in L2.
// The earlier tensor-core attempt (an im2col tensor in global memory + the generic 1 x 1 kernel) lost to the CUDA-core
// kernel on exactly those two points: the extra round trip of the im2col tensor and 64-byte store segments.
constexpr int kC0tTile = 128;                            // pixels per tile = UMMA M
constexpr int kC0tABytes = kC0tTile * 128;               // one operand plane of a tile: 128 rows x 128 B
constexpr int kC0tBBytes = 2 * 64 * 128;                 // [B_hi 64 rows | B_lo 64 rows] x 128 B
constexpr int kC0tStaging = 8 * 32 * 128;                // 32 pixel half-records of 128 B per epilogue warp
constexpr int kC0tProducers = 128;
constexpr int kC0tThreads = 13 * 32;                     // 4 producer warps, 8 epilogue warps, the MMA warp
constexpr int kC0tSmem = 1024 + 2 * 2 * kC0tABytes + kC0tBBytes + kC0tStaging + 256 /*barriers*/ + 256 /*bias*/;

__global__ void __launch_bounds__(kC0tThreads, 1)
conv0_tc_kernel(const float* __restrict__ img, const float* __restrict__ w /*[27][64]*/, const float* __restrict__ bias,
                __half* __restrict__ out, int B, int H, int W) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                // [2 stages][A_hi 16 KB | A_lo 16 KB]
  uint8_t* smem_b = smem + 2 * 2 * kC0tABytes;           // [B_hi | B_lo]
  uint8_t* staging = smem_b + kC0tBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kC0tStaging);
  uint64_t* a_full = bars;                               // [2]  producers -> MMA
  uint64_t* a_empty = bars + 2;                          // [2]  MMA commit -> producers
  uint64_t* tmem_full = bars + 4;                        // [2]  MMA commit -> epilogue
  uint64_t* tmem_empty = bars + 6;                       // [2]  epilogue (4 warps) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = W / kC0tTile, n_tiles = B * H * tiles_x;
  if (warp == 4 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full + s, kC0tProducers); mbar_init(a_empty + s, 1);
      mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 8);
    }
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc_512(tmem_slot);
  {
    // weight tile: row n = output channel, 32 K values (k = tap * 3 + c, zero from 27 on) = four 16-byte chunks at
    // (chunk ^ (n & 7)) of the 128-byte row; thread = (n, chunk)
    if (threadIdx.x < 256) {
      const int n = threadIdx.x & 63, j = threadIdx.x >> 6;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k0 = j * 8 + 2 * e;
        const float v0 = k0 < 27 ? __ldg(w + k0 * 64 + n) : 0.f, v1 = k0 + 1 < 27 ? __ldg(w + (k0 + 1) * 64 + n) : 0.f;
        split2(v0, v1, hi[e], lo[e]);
      }
      const uint32_t row = smem_u32(smem_b) + n * 128 + ((j ^ (n & 7)) << 4);
      st_shared_v4(row, hi[0], hi[1], hi[2], hi[3]);
      st_shared_v4(row + 64 * 128, lo[0], lo[1], lo[2], lo[3]);
    }
    if (threadIdx.x < 64) bias_s[threadIdx.x] = bias[threadIdx.x];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic stores -> the MMA's (async proxy) reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ===================== im2col producers =====================
    const int m = threadIdx.x;                                       // this thread's pixel of every tile
    auto load_tile = [&](int tile, float (&v)[27]) {
      if (tile >= n_tiles) return;
      const int xt = tile % tiles_x; int r = tile / tiles_x;
      const int y = r % H, b = r / H;
      const float* base = img + (size_t)b * 3 * H * W;
      const int x = xt * kC0tTile + m;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        const bool yok = yy >= 0 && yy < H;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          const bool ok = yok && xx >= 0 && xx < W;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            v[(ky * 3 + kx) * 3 + c] = ok ? __ldg(base + ((size_t)c * H + yy) * W + xx) : 0.f;
        }
      }
    };
    int sa = 0; uint32_t pa = 0;
    auto put_tile = [&](const float (&cur)[27]) {                    // split + store this thread's row of the operand tiles
      mbar_wait(a_empty + sa, pa ^ 1);
      const uint32_t row = smem_u32(smem_a) + sa * 2 * kC0tABytes + m * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k0 = j * 8 + 2 * e;
          split2(k0 < 27 ? cur[k0 < 27 ? k0 : 0] : 0.f, k0 + 1 < 27 ? cur[k0 + 1 < 27 ? k0 + 1 : 0] : 0.f, hi[e], lo[e]);
        }
        const uint32_t off = (uint32_t)((j ^ (m & 7)) << 4);
        st_shared_v4(row + off, hi[0], hi[1], hi[2], hi[3]);
        st_shared_v4(row + kC0tABytes + off, lo[0], lo[1], lo[2], lo[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(a_full + sa);
      if (++sa == 2) { sa = 0; pa ^= 1; }
    };
    // three register buffers rotate statically (the loop body handles three tiles): a tile's loads are issued two
    // tiles before its values are converted
    float t0[27], t1[27], t2[27];
    const int step = gridDim.x;
    int tile = blockIdx.x;
    load_tile(tile, t0); load_tile(tile + step, t1);
    while (tile < n_tiles) {
      load_tile(tile + 2 * step, t2); put_tile(t0); tile += step;
      if (tile >= n_tiles) break;
      load_tile(tile + 2 * step, t0); put_tile(t1); tile += step;
      if (tile >= n_tiles) break;
      load_tile(tile + 2 * step, t1); put_tile(t2); tile += step;
    }
  } else if (warp == 12) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_n128 = umma_idesc_f16(128);             // A_hi x [B_hi | B_lo]
    constexpr uint32_t idesc_n64 = umma_idesc_f16(64);               // A_lo x B_hi
    const uint64_t desc_b = umma_desc_sw128(smem_u32(smem_b));
    int sa = 0; uint32_t pa = 0; int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int as = t & 1; const uint32_t aphase = (t >> 1) & 1;
      mbar_wait(tmem_empty + as, aphase ^ 1);
      mbar_wait(a_full + sa, pa);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t acc0 = tmem_base + as * 256, acc1 = acc0 + 128;
        const uint64_t da = umma_desc_sw128(smem_u32(smem_a) + sa * 2 * kC0tABytes);
        const uint64_t dl = umma_desc_sw128(smem_u32(smem_a) + sa * 2 * kC0tABytes + kC0tABytes);
#pragma unroll
        for (int k = 0; k < 2; ++k) {                                // K = 32: two K = 16 steps, +32 B = +2 in the address field
          umma_f16(acc0, da + 2 * k, desc_b + 2 * k, idesc_n128, k != 0);
          umma_f16(acc1, dl + 2 * k, desc_b + 2 * k, idesc_n64, k != 0);
        }
        umma_commit(a_empty + sa);
        umma_commit(tmem_full + as);
      }
      __syncwarp();
      if (++sa == 2) { sa = 0; pa ^= 1; }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                                          // TMEM lane quarter = pixels [32 q, 32 q + 32) of the tile
    const int c0 = warp >= 8 ? 32 : 0;                               // this warp's 32 output channels
    const uint32_t stg = smem_u32(staging) + (warp - 4) * (32 * 128);
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int as = t & 1; const uint32_t aphase = (t >> 1) & 1;
      mbar_wait(tmem_full + as, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
      const uint32_t rec = stg + lane * 128;
      {
        uint32_t r0[32], r1[32], r2[32];
        tmem_ld_x32(taddr + c0, r0); tmem_ld_x32(taddr + 64 + c0, r1); tmem_ld_x32(taddr + 128 + c0, r2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + as);                 // the accumulators are in registers: the next tile's MMAs may run
#pragma unroll
        for (int j = 0; j < 4; ++j) {                                // 8 channels = one 16-byte piece of hi and of lo
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i0 = j * 8 + 2 * e;
            const float v0 = fmaf(__uint_as_float(r1[i0]) + __uint_as_float(r2[i0]), kLoInvScale, __uint_as_float(r0[i0])) + bias_s[c0 + i0];
            const float v1 = fmaf(__uint_as_float(r1[i0 + 1]) + __uint_as_float(r2[i0 + 1]), kLoInvScale, __uint_as_float(r0[i0 + 1])) + bias_s[c0 + i0 + 1];
            split2(fmaxf(v0, 0.f), fmaxf(v1, 0.f), hi[e], lo[e]);
          }
          st_shared_v4(rec + ((j ^ (lane & 7)) << 4), hi[0], hi[1], hi[2], hi[3]);
          st_shared_v4(rec + (((j + 4) ^ (lane & 7)) << 4), lo[0], lo[1], lo[2], lo[3]);
        }
      }
      __syncwarp();
      // write-out: four pixels per store instruction, each a 64-byte hi and a 64-byte lo segment of its record
      {
        const int sub = lane >> 3, piece = lane & 7;                 // piece 0-3: hi, 4-7: lo
        __half* gp = out + ((size_t)tile * kC0tTile + q * 32 + sub) * 128 + (piece >> 2) * 64 + c0 + (piece & 3) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int px = 4 * i + sub;
          const uint4 d = ld_shared_v4(stg + px * 128 + ((piece ^ (px & 7)) << 4));
          *reinterpret_cast<uint4*>(gp + (size_t)i * 512) = d;
        }
      }
      __syncwarp();                                                  // staging is rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) { tc_fence_after(); tmem_dealloc_512(tmem_base); }
}

// fp32 NHWC -> hi/lo activation planes (used by ha_conv3x3_nhwc, the single-layer test entry)
__global__ void split_act_kernel(const float* __restrict__ in, __half* __restrict__ out, int C, size_t n_px) {
  const size_t total = n_px * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / C; const int c = (int)(i % C);
    const float v = in[i];
    const __half h = __float2half_rn(v);
    out[p * 2 * C + c] = h;
    out[p * 2 * C + C + c] = __float2half_rn((v - __half2float(h)) * kLoScale);
  }
}

// train mode: x15 = maxpool2x2(x14) as the fp32 feature, and relu(x15) replicated 2x2 into the first 256 channels of cat1
// (what conv14's epilogue does in one go on the eval path).  One thread per pooled pixel and 8 channels.
__global__ void pool_x15_kernel(const float* __restrict__ x14, float* __restrict__ x15, __half* __restrict__ cat1, int B, int h, int w) {
  const size_t total = (size_t)B * (h / 2) * (w / 2) * 32;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 32) * 8;
    size_t p = i / 32;
    const int x = (int)(p % (w / 2)); p /= (w / 2);
    const int y = (int)(p % (h / 2)); const size_t b = p / (h / 2);
    float m[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* s = x14 + ((b * h + 2 * y + (k >> 1)) * w + 2 * x + (k & 1)) * 256 + c;
      const float4 v0 = *reinterpret_cast<const float4*>(s), v1 = *reinterpret_cast<const float4*>(s + 4);
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = k == 0 ? v[e] : fmaxf(m[e], v[e]);
    }
    float* d = x15 + ((b * (h / 2) + y) * (w / 2) + x) * 256 + c;
    *reinterpret_cast<float4*>(d) = make_float4(m[0], m[1], m[2], m[3]);
    *reinterpret_cast<float4*>(d + 4) = make_float4(m[4], m[5], m[6], m[7]);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(fmaxf(m[2 * e], 0.f), fmaxf(m[2 * e + 1], 0.f), hi[e], lo[e]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __half* o = cat1 + ((b * h + 2 * y + (k >> 1)) * w + 2 * x + (k & 1)) * (2 * 384) + c;
      *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(o + 384) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// VGGUnet_G2S (VGG.py:275-345): every map the decoders see is the [H, W] map re-interpreted as [2H, W/2] (row-major pixel
// order unchanged, so in NHWC the re-shape is free); only the x2 nearest upsample depends on the folded geometry.
// dst pixel (Y, X) of the folded fine map [2 hs][2 ws] = relu(src pixel (Y/2, X/2)) of the folded coarse map [hs][ws].
__global__ void relu_up2x_fold_kernel(const float* __restrict__ src, __half* __restrict__ dst, int pitch, int coff, int C, int B, int hs,
                                      int ws) {
  const int C8 = C / 8;
  const size_t total = (size_t)B * hs * ws * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8) * 8;
    size_t p = i / C8;
    const int x = (int)(p % ws); p /= ws;
    const int y = (int)(p % hs); const size_t b = p / hs;
    const float* s = src + ((b * hs + y) * ws + x) * C + c;
    const float4 v0 = *reinterpret_cast<const float4*>(s), v1 = *reinterpret_cast<const float4*>(s + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(fmaxf(v[2 * e], 0.f), fmaxf(v[2 * e + 1], 0.f), hi[e], lo[e]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __half* o = dst + ((b * 2 * hs + 2 * y + (k >> 1)) * (size_t)(2 * ws) + 2 * x + (k & 1)) * (2 * pitch) + coff + c;
      *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(o + pitch) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// cuTensorMapEncodeTiled is a pure function of its arguments and costs a few microseconds of host time; a U-Net forward
// needs 60 of them and the workspace pointers repeat from call to call, so the encoded maps are memoised (bounded).
struct MapKey {
  const void* base; unsigned long long d[5], s[4]; unsigned box[5]; int rank;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey); ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h;
  }
};
int encode_cached(CUtensorMap* m, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, const char* what) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> memo;
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.base = base; k.rank = rank;
  for (int i = 0; i < rank; ++i) { k.d[i] = dims[i]; k.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.s[i] = strides[i];
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = memo.find(k);
    if (it != memo.end()) { *m = it->second; return HA_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_cuda_error(cudaErrorUnknown, "cuTensorMapEncodeTiled entry point"); return HA_ECUDA; }
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_cuda_error(cudaErrorInvalidValue, what); return HA_ECUDA; }
  std::lock_guard<std::mutex> g(mu);
  if (memo.size() >= 1024) memo.clear();
  memo.emplace(k, *m);
  return HA_OK;
}

// activation tensor [B][H][W][2][pitch] (fp16), channel slice [coff, coff + cin): dims (C, plane, W, H, B)
static int make_act_map(CUtensorMap* m, const __half* base, int pitch, int coff, int cin, int B, int H, int W,
                        int box_w = kTileW, int box_h = kTileH) {
  cuuint64_t dims[5] = {(cuuint64_t)cin, 2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 4, (cuuint64_t)W * pitch * 4, (cuuint64_t)H * W * pitch * 4};
  cuuint32_t box[5] = {kBlockK, 1, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  return encode_cached(m, 5, base + coff, dims, strides, box, "cuTensorMapEncodeTiled(activations)");
}

// weights [9][cout_pad][cin_pad] fp16: dims (Cin, Cout, tap)
static int make_weight_map(CUtensorMap* m, const __half* base, int cin_pad, int cout_pad, int block_n, int n_taps) {
  cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)n_taps};
  cuuint64_t strides[2] = {(cuuint64_t)cin_pad * 2, (cuuint64_t)cin_pad * cout_pad * 2};
  cuuint32_t box[3] = {kBlockK, (cuuint32_t)block_n, 1};
  return encode_cached(m, 3, base, dims, strides, box, "cuTensorMapEncodeTiled(weights)");
}

// cudaFuncSetAttribute once per (kernel instantiation, device): `flag` is that instantiation's bit set of configured devices
template <typename K>
static int set_smem_once(K kern, int bytes, std::atomic<unsigned long long>& flag) {
  int dev = 0;
  HA_CUDA_TRY(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(flag.load(std::memory_order_acquire) & bit)) {
    HA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    flag.fetch_or(bit, std::memory_order_release);
  }
  return HA_OK;
}

template <int BLOCK_N, bool SPLIT>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tbh, const CUtensorMap& tbl, const TcConvArgs& a, cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N, SPLIT>;
  auto kern = conv3x3_tc_kernel<BLOCK_N, SPLIT>;
  static std::atomic<unsigned long long> configured{0};
  if (int rc = set_smem_once(kern, Cfg::kSmemBytes, configured)) return rc;
  const int grid = a.n_tiles < kNumSMs ? a.n_tiles : kNumSMs;
  kern<<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(ta, tbh, tbl, a);
  count_launches(1);
  return check_launch("conv3x3_tc_kernel");
}

// One 3x3 conv layer on the tensor cores.  in: activation planes (pitch/coff/cin), weights from the packed buffer.
// CTA-pair launch: 74 clusters of two on a full B200; fewer if the device cannot co-schedule that many pairs.
template <int BLOCK_N, bool RESIDENT>
static int launch_halo2(const CUtensorMap& ta, const CUtensorMap& tbh, const CUtensorMap& tbl, const CUtensorMap& tbh_half,
                        const TcConvArgs& a, cudaStream_t st) {
  using Cfg = Halo2Cfg<BLOCK_N>;
  auto kern = conv3x3_tc_halo2_kernel<BLOCK_N, RESIDENT>;
  static std::atomic<unsigned long long> configured{0};
  if (int rc = set_smem_once(kern, Cfg::kSmemBytes, configured)) return rc;
  static std::atomic<int> max_pairs{0};
  int mp = max_pairs.load(std::memory_order_acquire);
  if (mp == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = kNumSMs / 2; }
    mp = n < kNumSMs / 2 ? n : kNumSMs / 2;
    max_pairs.store(mp, std::memory_order_release);
  }
  const int n_pairs = a.n_tiles / 2;
  const int grid = 2 * (n_pairs < mp ? n_pairs : mp);
  kern<<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(ta, tbh, tbl, tbh_half, a);
  count_launches(1);
  return check_launch("conv3x3_tc_halo2_kernel");
}

int conv_tc(const __half* in, int in_pitch, int in_coff, int cin, const char* packed, const PackedConv& pc, int cout,
            bool has_bias, const TcOut& o, int B, int H, int W, bool split, cudaStream_t st, int n_taps, bool pair) {
  if ((W % kTileW) || (H % kTileH) || (cin % 8) || (in_coff % 8) || (in_pitch % 8)) return HA_EINVAL;
  const int block_n = cout >= 128 ? 128 : cout;
  if (block_n != 128 && block_n != 64 && block_n != 32 && block_n != 16) return HA_EINVAL;
  CUtensorMap ta, tbh, tbl;
  // Cout >= 32 in f16x3 mode, 16 x 8 pixel tiles in even number: the CTA-pair halo kernel (cta_group::2)
  if (pair && split && block_n >= 32 && n_taps == 9 && (W % kHaloTW) == 0 && (H % kHaloTH) == 0 && (cin % 64) == 0 &&
      (((W / kHaloTW) * (H / kHaloTH) * B) % 2) == 0) {
    CUtensorMap tbh_half;
    int rc = make_act_map(&ta, in, in_pitch, in_coff, cin, B, H, W, kHaloBoxW, kHaloBoxH);
    if (rc != HA_OK) return rc;
    const __half* whi = reinterpret_cast<const __half*>(packed + pc.hi);
    if ((rc = make_weight_map(&tbh, whi, pc.cin_pad, pc.cout_pad, block_n, n_taps)) != HA_OK) return rc;
    if ((rc = make_weight_map(&tbl, reinterpret_cast<const __half*>(packed + pc.lo), pc.cin_pad, pc.cout_pad, block_n, n_taps)) != HA_OK) return rc;
    if ((rc = make_weight_map(&tbh_half, whi, pc.cin_pad, pc.cout_pad, block_n / 2, n_taps)) != HA_OK) return rc;
    TcConvArgs a;
    a.bias = has_bias ? reinterpret_cast<const float*>(packed + pc.bias) : nullptr;
    a.act_full = o.act_full; a.af_pitch = o.af_pitch; a.af_coff = o.af_coff;
    a.act_pool = o.act_pool; a.ap_pitch = o.ap_pitch; a.ap_coff = o.ap_coff;
    a.act_up = o.act_up; a.au_pitch = o.au_pitch; a.au_coff = o.au_coff;
    a.feat = o.feat; a.feat_pooled = o.feat_pooled;
    a.B = B; a.H = H; a.W = W; a.cout = cout;
    a.n_kchunks = cin / kBlockK; a.n_taps = 9;
    a.tiles_x = W / kHaloTW; a.tiles_y = H / kHaloTH; a.tiles_n = cout / block_n;
    a.n_tiles = a.tiles_x * a.tiles_y * B * a.tiles_n;
    if (block_n == 128) return launch_halo2<128, false>(ta, tbh, tbl, tbh_half, a, st);
    if (block_n == 32) return launch_halo2<32, false>(ta, tbh, tbl, tbh_half, a, st);
    return (a.n_kchunks == 1 && a.tiles_n == 1) ? launch_halo2<64, true>(ta, tbh, tbl, tbh_half, a, st)
                                                : launch_halo2<64, false>(ta, tbh, tbl, tbh_half, a, st);
  }
  // Cout = 64 layers in f16x3 mode: halo-tile kernel (one activation load per 64-channel chunk instead of nine)
  const bool halo = split && cout == 64 && n_taps == 9 && (W % kHaloTW) == 0 && (H % kHaloTH) == 0 && (cin % 64) == 0;
  int rc = halo ? make_act_map(&ta, in, in_pitch, in_coff, cin, B, H, W, kHaloBoxW, kHaloBoxH)
                : make_act_map(&ta, in, in_pitch, in_coff, cin, B, H, W);
  if (rc != HA_OK) return rc;
  rc = make_weight_map(&tbh, reinterpret_cast<const __half*>(packed + pc.hi), pc.cin_pad, pc.cout_pad, block_n, n_taps);
  if (rc != HA_OK) return rc;
  rc = make_weight_map(&tbl, reinterpret_cast<const __half*>(packed + pc.lo), pc.cin_pad, pc.cout_pad, block_n, n_taps);
  if (rc != HA_OK) return rc;
  TcConvArgs a;
  a.bias = has_bias ? reinterpret_cast<const float*>(packed + pc.bias) : nullptr;
  a.act_full = o.act_full; a.af_pitch = o.af_pitch; a.af_coff = o.af_coff;
  a.act_pool = o.act_pool; a.ap_pitch = o.ap_pitch; a.ap_coff = o.ap_coff;
  a.act_up = o.act_up; a.au_pitch = o.au_pitch; a.au_coff = o.au_coff;
  a.feat = o.feat; a.feat_pooled = o.feat_pooled;
  a.B = B; a.H = H; a.W = W; a.cout = cout;
  a.n_kchunks = (cin + kBlockK - 1) / kBlockK;
  a.n_taps = n_taps;
  a.tiles_x = W / kTileW; a.tiles_y = H / kTileH; a.tiles_n = cout / block_n;
  a.n_tiles = a.tiles_x * a.tiles_y * a.tiles_n * B;
  if (halo) {
    a.tiles_x = W / kHaloTW; a.tiles_y = H / kHaloTH; a.tiles_n = 1;
    a.n_tiles = a.tiles_x * a.tiles_y * B;
    const int grid = a.n_tiles < kNumSMs ? a.n_tiles : kNumSMs;
    static std::atomic<unsigned long long> halo_configured{0};
    if (int rc2 = set_smem_once(conv3x3_tc_halo_kernel, kHaloSmemBytes, halo_configured)) return rc2;
    conv3x3_tc_halo_kernel<<<grid, kTcThreads, kHaloSmemBytes, st>>>(ta, tbh, tbl, a);
    count_launches(1);
    return check_launch("conv3x3_tc_halo_kernel");
  }
#define HA_TC_CASE(N)                                                              \
  case N: return split ? launch_tc<N, true>(ta, tbh, tbl, a, st) : launch_tc<N, false>(ta, tbh, tbl, a, st);
  switch (block_n) { HA_TC_CASE(128) HA_TC_CASE(64) HA_TC_CASE(32) HA_TC_CASE(16) }
#undef HA_TC_CASE
  return HA_EINVAL;
}

// Workspace layout of the tensor-core schedule (one sequential carve-up, shared with the backward pass, which finds the
// saved activations of a train-mode forward by repeating it on the same base pointer)
TcSaved vgg_tc_carve(Arena& ar, int B, int H, int W, int n_levels, bool train) {
  const size_t px1 = (size_t)B * H * W, px2 = px1 / 4, px4 = px1 / 16;
  auto act = [&](size_t px, int c) { return (__half*)ar.take(px * 2 * c * sizeof(__half)); };
  TcSaved s{};
  s.a1 = act(px1, 64);
  s.cat3 = n_levels == 4 ? act(px1, 128) : nullptr;
  s.cat2 = act(px2, 192);
  s.a5 = act(px2, 128);
  s.cat1 = act(px4, 384);
  s.a10 = act(px4, 256);
  s.a12 = act(px4, 256);
  s.d1 = act(px4, 128);
  s.d2 = act(px2, 64);
  s.d3 = n_levels == 4 ? act(px1, 32) : nullptr;
  if (train) {
    s.x2 = (float*)ar.take(px1 * 64 * sizeof(float));
    s.x7 = (float*)ar.take(px2 * 128 * sizeof(float));
    s.x14 = (float*)ar.take(px4 * 256 * sizeof(float));
  }
  return s;
}

// Tensor-core schedule of the U-Net; buffer names follow vgg.cu / VGG.py:121-158.
int vgg_forward_tc(const char* packed, const PackedLayout& L, const float* img, int B, int H, int W, int n_levels,
                   int precision, float* const* out_feat, Arena& ar, cudaStream_t st, TcSaved* saved, bool g2s) {
  const bool split = precision == HA_CONV_F16X3 || precision == HA_CONV_F16X3_1CTA;
  const bool pair = precision == HA_CONV_F16X3;
  const size_t px1 = (size_t)B * H * W, px2 = px1 / 4, px4 = px1 / 16;
  const TcSaved sv = vgg_tc_carve(ar, B, H, W, n_levels, saved != nullptr);
  __half *a1 = sv.a1, *cat3 = sv.cat3, *cat2 = sv.cat2, *a5 = sv.a5, *cat1 = sv.cat1, *a10 = sv.a10, *a12 = sv.a12;
  __half *d1 = sv.d1, *d2 = sv.d2, *d3 = sv.d3;
  float *x2 = sv.x2, *x7 = sv.x7, *x14 = sv.x14;
  if (saved) *saved = sv;
  if (ar.dry) return HA_OK;
  if (ar.off > ar.cap) return HA_ENOSPACE;
  if ((W % (4 * kTileW)) || (H % (4 * kTileH))) return HA_EINVAL;   // tiles must fit down to the 1/4 scale
  int rc;
#define HA_TRY(x) do { rc = (x); if (rc != HA_OK) return rc; } while (0)
  // conv0: on the tensor pipe like the other layers when the operands are split (conv0_tc_kernel: tiles of 128 pixels
  // of one image row), else on the CUDA cores in fp32 (conv0_kernel; W % 64 == 0 and H % 32 == 0 were checked above)
  const float* w0 = reinterpret_cast<const float*>(packed + L.c[L_CONV0].f32);
  const float* b0 = reinterpret_cast<const float*>(packed + L.c[L_CONV0].bias);
  if (split && (W % kC0tTile) == 0 && (long long)B * H * (W / kC0tTile) < (1ll << 31)) {
    static std::atomic<unsigned long long> c0t_configured{0};
    if (int rc0 = set_smem_once(conv0_tc_kernel, kC0tSmem, c0t_configured)) return rc0;
    const long long n_tiles = (long long)B * H * (W / kC0tTile);
    conv0_tc_kernel<<<(int)std::min<long long>(n_tiles, kNumSMs), kC0tThreads, kC0tSmem, st>>>(img, w0, b0, a1, B, H, W);
    count_launches(1);
    HA_TRY(check_launch("conv0_tc_kernel"));
  } else {
    static std::atomic<unsigned long long> c0_configured{0};
    if (int rc0 = set_smem_once(conv0_kernel, kC0Smem, c0_configured)) return rc0;
    conv0_kernel<<<dim3(W / kC0TW, H / kC0TH, B), 256, kC0Smem, st>>>(img, w0, b0, a1, H, W);
    count_launches(1);
    HA_TRY(check_launch("conv0_kernel"));
  }
  auto conv = [&](int li, const __half* in, int pitch, int coff, const TcOut& o, int h, int w) {
    return conv_tc(in, pitch, coff, kVggConvs[li].cin, packed, L.c[li], kVggConvs[li].cout, kVggConvs[li].has_bias != 0, o,
                   B, h, w, split, st, 9, pair);
  };
  TcOut o;
  o = TcOut(); o.act_pool = cat2; o.ap_pitch = 192; o.ap_coff = 128;                    // x4 = relu(pool(x2))
  if (n_levels == 4) { o.act_full = cat3; o.af_pitch = 128; o.af_coff = 64; }           // relu(x2) skip for dec3
  if (saved) o.feat = x2;
  HA_TRY(conv(L_CONV2, a1, 64, 0, o, H, W));
  o = TcOut(); o.act_full = a5; o.af_pitch = 128;
  HA_TRY(conv(L_CONV5, cat2, 192, 128, o, H / 2, W / 2));
  o = TcOut(); o.act_pool = cat1; o.ap_pitch = 384; o.ap_coff = 256;                    // x9 = relu(pool(x7))
  if (saved) o.feat = x7;
  HA_TRY(conv(L_CONV7, a5, 128, 0, o, H / 2, W / 2));
  o = TcOut(); o.act_full = a10; o.af_pitch = 256;
  HA_TRY(conv(L_CONV10, cat1, 384, 256, o, H / 4, W / 4));
  o = TcOut(); o.act_full = a12; o.af_pitch = 256;
  HA_TRY(conv(L_CONV12, a10, 256, 0, o, H / 4, W / 4));
  auto up_fold = [&](const float* src, __half* dst, int pitch, int C, int hs, int ws) {      // folded coarse dims
    const size_t n = (size_t)B * hs * ws * (C / 8);
    relu_up2x_fold_kernel<<<(unsigned)((n + 255) / 256 < (size_t)kNumSMs * 16 ? (n + 255) / 256 : (size_t)kNumSMs * 16), 256, 0, st>>>(
        src, dst, pitch, 0, C, B, hs, ws);
    count_launches(1);
  };
  // decoder geometry: VGGUnet runs on [H/4, W/4], [H/2, W/2], [H, W]; VGGUnet_G2S on the folded maps [H/2, W/8], [H, W/4], [2H, W/2]
  const int h4 = g2s ? H / 2 : H / 4, w4 = g2s ? W / 8 : W / 4;
  const int h2 = 2 * h4, w2 = 2 * w4, h1 = 4 * h4, w1 = 4 * w4;
  if (g2s && saved) return HA_EINVAL;
  if (saved) {                                                                          // x14 kept; x15 and its upsample by a kernel
    o = TcOut(); o.feat = x14;
    HA_TRY(conv(L_CONV14, a12, 256, 0, o, H / 4, W / 4));
    const size_t n = px4 / 4 * 32;
    pool_x15_kernel<<<(unsigned)((n + 255) / 256 < (size_t)kNumSMs * 16 ? (n + 255) / 256 : (size_t)kNumSMs * 16), 256, 0, st>>>(
        x14, out_feat[0], cat1, B, H / 4, W / 4);
    count_launches(1);
  } else if (g2s) {
    o = TcOut(); o.feat = out_feat[0]; o.feat_pooled = 1;                                                     // x15, no upsample here
    HA_TRY(conv(L_CONV14, a12, 256, 0, o, H / 4, W / 4));
    up_fold(out_feat[0], cat1, 384, 256, h4 / 2, w4 / 2);                                                     // x16 = up(x15_)  (VGG.py:303)
  } else {
    o = TcOut(); o.feat = out_feat[0]; o.feat_pooled = 1; o.act_up = cat1; o.au_pitch = 384; o.au_coff = 0;   // x15
    HA_TRY(conv(L_CONV14, a12, 256, 0, o, H / 4, W / 4));
  }
  o = TcOut(); o.act_full = d1; o.af_pitch = 128;
  HA_TRY(conv(L_DEC1A, cat1, 384, 0, o, h4, w4));
  o = TcOut(); o.feat = out_feat[1];                                                                          // x18
  if (!g2s) { o.act_up = cat2; o.au_pitch = 192; o.au_coff = 0; }
  HA_TRY(conv(L_DEC1B, d1, 128, 0, o, h4, w4));
  if (g2s) up_fold(out_feat[1], cat2, 192, 128, h4, w4);                                                      // x19 (:308)
  o = TcOut(); o.act_full = d2; o.af_pitch = 64;
  HA_TRY(conv(L_DEC2A, cat2, 192, 0, o, h2, w2));
  o = TcOut(); o.feat = out_feat[2];                                                                          // x21
  if (n_levels == 4 && !g2s) { o.act_up = cat3; o.au_pitch = 128; o.au_coff = 0; }
  HA_TRY(conv(L_DEC2B, d2, 64, 0, o, h2, w2));
  if (n_levels == 4) {
    if (g2s) up_fold(out_feat[2], cat3, 128, 64, h2, w2);                                                     // x22 (:312)
    o = TcOut(); o.act_full = d3; o.af_pitch = 32;
    HA_TRY(conv(L_DEC3A, cat3, 128, 0, o, h1, w1));
    o = TcOut(); o.feat = out_feat[3];                                                                        // x24
    HA_TRY(conv(L_DEC3B, d3, 32, 0, o, h1, w1));
  }
#undef HA_TRY
  return HA_OK;
}

}  // namespace ha

// ------------------------------------------------------------------------------ single-layer entry
static ha::PackedConv single_layout(int cin, int cout, size_t* total) {
  ha::PackedConv p;
  p.cin_pad = (int)ha::align_up(cin, 64);
  p.cout_pad = (int)ha::align_up(cout, 16);
  size_t off = 0;
  p.f32 = off; off = ha::align_up(off + (size_t)9 * cin * cout * 4, 256);
  p.hi = off; off = ha::align_up(off + (size_t)9 * p.cout_pad * p.cin_pad * 2, 256);
  p.lo = off; off = ha::align_up(off + (size_t)9 * p.cout_pad * p.cin_pad * 2, 256);
  p.bias = off; off = ha::align_up(off + (size_t)cout * 4, 256);
  *total = off;
  return p;
}

extern "C" size_t ha_conv3x3_workspace_bytes(int cin, int cout, int B, int H, int W) {
  if (cin <= 0 || cout <= 0 || B <= 0 || H <= 0 || W <= 0) return 0;
  size_t wbytes;
  single_layout(cin, cout, &wbytes);
  return wbytes + ha::align_up((size_t)B * H * W * 2 * cin * 2, 256) + 256;
}

extern "C" int ha_conv3x3_nhwc(const float* in_nhwc, int cin, const float* w_oihw, const float* bias, float* out_nhwc, int cout,
                               int B, int H, int W, int precision, void* ws, size_t ws_bytes, void* stream) {
  if (!in_nhwc || !w_oihw || !out_nhwc || !ws) return HA_EINVAL;
  if (ws_bytes < ha_conv3x3_workspace_bytes(cin, cout, B, H, W)) return HA_ENOSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  size_t wbytes;
  const ha::PackedConv pc = single_layout(cin, cout, &wbytes);
  char* base = reinterpret_cast<char*>(ws);
  ha::pack_conv_kernel<<<64, 256, 0, st>>>(w_oihw, bias, cin, cout, pc.cin_pad, pc.cout_pad, reinterpret_cast<float*>(base + pc.f32),
                                           reinterpret_cast<__half*>(base + pc.hi), reinterpret_cast<__half*>(base + pc.lo),
                                           reinterpret_cast<float*>(base + pc.bias));
  ha::count_launches(1);
  if (precision == HA_CONV_FP32_SIMT)
    return ha::conv_simt(in_nhwc, cin, 0, cin, reinterpret_cast<const float*>(base + pc.f32),
                         bias ? reinterpret_cast<const float*>(base + pc.bias) : nullptr, out_nhwc, cout, 0, cout, B, H, W, 0, st);
  if (precision != HA_CONV_F16X3 && precision != HA_CONV_F16 && precision != HA_CONV_F16X3_1CTA) return HA_EINVAL;
  __half* act = reinterpret_cast<__half*>(base + wbytes);
  const size_t n_px = (size_t)B * H * W;
  ha::split_act_kernel<<<ha::kNumSMs * 8, 256, 0, st>>>(in_nhwc, act, cin, n_px);
  ha::count_launches(1);
  ha::TcOut o;
  o.feat = out_nhwc;
  return ha::conv_tc(act, cin, 0, cin, base, pc, cout, bias != nullptr, o, B, H, W, precision != HA_CONV_F16, st, 9,
                     precision == HA_CONV_F16X3);
}
