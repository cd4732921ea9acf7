// Backward pass of the VGG16 U-Net (what `loss.backward()` of train_kitti.py:365 does inside
// VGGUnet.forward, VGG.py:121-203), sm_100a.  Input: the gradients w.r.t. the raw pyramid features x15 / x18 / x21 and the
// activations a train-mode forward kept (ha_vgg_forward_train); output: the gradients of the 13 convolution weights and 7
// biases in torch's OIHW layout.
//
// Per convolution layer  y = conv3x3(x, W) (+ b):
//  * data gradient  dx = conv3x3(dy, W flipped by 180 degrees with the channel roles swapped): the SAME tcgen05 / TMA
//    implicit-GEMM kernels as the forward pass (vgg_tc.cu conv_tc: f16x3 split, CTA pairs) on re-packed weights;
//  * weight gradient  dW[co][ci][ky][kx] = sum_p dy[p][co] x[p + (ky-1, kx-1)][ci]: a GEMM whose K dimension is the pixel
//    index.  Both operands are first transposed into [channel][padded pixel] planes (zero pixels all round every image,
//    row pitch W + 8, so a tap is a constant offset along the pixel axis and row wrap-around multiplies a zero of dy),
//    which makes them K-major: the operand tiles are plain 2-D TMA boxes with the 128-byte swizzle and the UMMA
//    descriptors are the ones the forward kernels use.  A TMA box must start on a 16-byte boundary (B200: an odd pixel
//    offset raises an illegal-instruction fault), so x is stored three times, pre-shifted by kx - 1 = -1 / 0 / +1 pixels,
//    and the ky shift is a whole padded row (a multiple of 8 pixels).  wgrad_tc_kernel: M = 128 output channels, N = a 64 (or 16) input-channel chunk, the three
//    kx taps of one ky per CTA (3 x 2 fp32 accumulators = 384 TMEM columns), split-K over the pixels, fp32 atomics into dW;
//  * fp16 range: every dy is multiplied by a power of two taken from its abs-max (device side, no host sync) before the
//    hi / lo split and the results are scaled back, so 1e-9-sized gradients keep 22 significant bits;
//  * ReLU masks come from the saved post-ReLU activations (x > 0), max-pool routing from the saved raw conv outputs (first
//    maximum in scan order, like torch's max_pool2d), the x2 nearest upsample becomes a 2x2 sum, concat becomes a split:
//    the decoder's first convolutions run two data-gradient convolutions (upsampled part / skip part).
#include <cuda.h>
#include <cuda_fp16.h>

#include <atomic>

#include "tc_ptx.cuh"
#include "vgg_common.cuh"

namespace ha {

int encode_cached(CUtensorMap* m, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, const char* what);

// ------------------------------------------------------------------------------ gradient scale (power of two)
// bits = float bits of max |g|; s = 2^(10 - exponent) puts the largest entry in [1024, 2048)
__device__ __forceinline__ float grad_scale(unsigned bits, float* inv) {
  const int e = (int)((bits >> 23) & 0xffu);
  int ex = (e == 0 || e == 255) ? 10 : e - 127;
  ex = ex < -100 ? -100 : (ex > 100 ? 100 : ex);
  *inv = exp2f((float)(ex - 10));
  return exp2f((float)(10 - ex));
}

__global__ void absmax_kernel(const float* __restrict__ g, size_t n, unsigned* __restrict__ bits) {
  float m = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(g);      // n is a multiple of 4 (channel counts are multiples of 16)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    const float a = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    if (a > m && a < INFINITY) m = a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));
}

__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((v0 - hf.x) * kLoScale, (v1 - hf.y) * kLoScale);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// fp32 NHWC gradient -> scaled hi / lo activation planes [P][2][C] (the format conv_tc reads); four channels per thread
__global__ void split_scaled_kernel(const float* __restrict__ g, const unsigned* __restrict__ bits, __half* __restrict__ out, int C,
                                    size_t n_px) {
  float inv;
  const float s = grad_scale(*bits, &inv);
  const int C4 = C / 4;
  const size_t total = n_px * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / C4; const int c = (int)(i % C4) * 4;
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    uint32_t h0, l0, h1, l1;
    split_pair(v.x * s, v.y * s, h0, l0);
    split_pair(v.z * s, v.w * s, h1, l1);
    *reinterpret_cast<uint2*>(out + p * 2 * C + c) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(out + p * 2 * C + C + c) = make_uint2(l0, l1);
  }
}

// bias gradient: column sums of dy [P][C]
__global__ void colsum_kernel(const float* __restrict__ g, size_t n_px, int C, float* __restrict__ db) {
  // blockDim.x = C (<= 256); each block sums a slice of the pixels
  const size_t per = (n_px + gridDim.x - 1) / gridDim.x;
  const size_t p0 = blockIdx.x * per, p1 = p0 + per < n_px ? p0 + per : n_px;
  float acc = 0.f;
  for (size_t p = p0; p < p1; ++p) acc += g[p * C + threadIdx.x];
  atomicAdd(db + threadIdx.x, acc);
}

// hi / lo activation planes [B][H][W][2][pitch] (channels [coff, coff + C)) -> K-major planes [shift][2][Cpad][Ppad] over the
// padded pixel index p = (b (H + 2) + y + 1)(W + 8) + x + 1; padding and the channel rows >= C are zero.  NSH == 3: copy s
// holds the pixel p + s - 1 at position p (the kx taps of the weight gradient); NSH == 1: the pixel itself.
// One CTA moves a 64-channel x 64-pixel tile (+ one pixel either side) through shared memory: 16-byte loads along the
// channels, 16-byte stores along the pixels, every source byte read once for all three copies.
// grid (Ppad / 64, Cpad / 64, 2 planes), 256 threads.
constexpr int kRowPad = 8;
constexpr int kTrPitch = 67;      // odd: channel rows 8 apart land in different banks
template <int NSH>
__global__ void __launch_bounds__(256) transpose_pad_kernel(const __half* __restrict__ src, int pitch, int coff, int C, int B, int H,
                                                            int W, __half* __restrict__ dst, int Cpad, size_t Ppad) {
  __shared__ __half tile[64 * kTrPitch];           // [channel][pixel p0 - 1 .. p0 + 64]
  const int plane = blockIdx.z, c0 = blockIdx.y * 64;
  const size_t p0 = (size_t)blockIdx.x * 64;
  const int Wp = W + kRowPad, Hp = H + 2;
  for (int item = threadIdx.x; item < 66 * 8; item += 256) {
    const int r = item >> 3, j = item & 7;         // tile pixel row (p0 - 1 + r), 8-channel chunk
    const long long ps = (long long)p0 - 1 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (ps >= 0 && c0 + 8 * j < C) {
      const size_t p = (size_t)ps;
      const int xx = (int)(p % Wp); const size_t t = p / Wp;
      const int yy = (int)(t % Hp); const size_t b = t / Hp;
      if (b < (size_t)B && xx >= 1 && xx <= W && yy >= 1 && yy <= H)
        v = *reinterpret_cast<const uint4*>(src + (((b * H + yy - 1) * W + xx - 1) * 2 + plane) * pitch + coff + c0 + 8 * j);
    }
    const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
    for (int e = 0; e < 8; ++e) tile[(8 * j + e) * kTrPitch + r] = h[e];
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NSH; ++s) {
    const int first = NSH == 3 ? s : 1;             // tile column of output pixel 0: p0 + (s - 1) -> column s
    for (int item = threadIdx.x; item < 64 * 8; item += 256) {
      const int c = item >> 3, g = item & 7;
      __align__(16) __half o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = tile[c * kTrPitch + 8 * g + k + first];
      *reinterpret_cast<uint4*>(dst + ((size_t)(s * 2 + plane) * Cpad + c0 + c) * Ppad + p0 + 8 * g) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// the input image (fp32 NCHW, 3 channels) as the K-major hi / lo planes [3 shifts][2][16][Ppad] of conv0's weight gradient
__global__ void image_pad_kernel(const float* __restrict__ img, int B, int H, int W, __half* __restrict__ dst, size_t Ppad) {
  const int Wp = W + kRowPad, Hp = H + 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 3 * Ppad; i += (size_t)gridDim.x * blockDim.x) {
    const int sh = (int)(i / Ppad);
    const size_t pd = i % Ppad;
    const long long ps = (long long)pd + sh - 1;
    const size_t p = ps < 0 ? (size_t)B * Hp * Wp : (size_t)ps;
    const int xx = (int)(p % Wp); size_t t = p / Wp;
    const int yy = (int)(t % Hp); const size_t b = t / Hp;
    const bool in = b < (size_t)B && xx >= 1 && xx <= W && yy >= 1 && yy <= H;
    for (int c = 0; c < 16; ++c) {
      float v = 0.f;
      if (in && c < 3) v = img[((b * 3 + c) * H + yy - 1) * W + xx - 1];
      const __half h = __float2half_rn(v);
      dst[((size_t)(sh * 2) * 16 + c) * Ppad + pd] = h;
      dst[((size_t)(sh * 2 + 1) * 16 + c) * Ppad + pd] = __float2half_rn((v - __half2float(h)) * kLoScale);
    }
  }
}

// out = add + R / s * [act > 0]      (R: data gradient of a convolution whose dy was scaled by s; act: the saved post-ReLU input);
// four channels per thread
__global__ void post_mask_kernel(const float* __restrict__ R, const unsigned* __restrict__ bits, const __half* __restrict__ act, int pitch,
                                 int coff, const float* __restrict__ add, float* __restrict__ out, int C, size_t n_px) {
  float inv;
  grad_scale(*bits, &inv);
  const int C4 = C / 4;
  const size_t total = n_px * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / C4; const int c = (int)(i % C4) * 4;
    const float4 r = reinterpret_cast<const float4*>(R)[i];
    float4 o = add ? reinterpret_cast<const float4*>(add)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    bool on[4] = {true, true, true, true};
    if (act) {
      const uint2 a = *reinterpret_cast<const uint2*>(act + p * 2 * pitch + coff + c);
      const __half2 a0 = *reinterpret_cast<const __half2*>(&a.x), a1 = *reinterpret_cast<const __half2*>(&a.y);
      on[0] = __low2float(a0) > 0.f; on[1] = __high2float(a0) > 0.f; on[2] = __low2float(a1) > 0.f; on[3] = __high2float(a1) > 0.f;
    }
    o.x += on[0] ? r.x * inv : 0.f; o.y += on[1] ? r.y * inv : 0.f; o.z += on[2] ? r.z * inv : 0.f; o.w += on[3] ? r.w * inv : 0.f;
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// backward of relu(nearest x2 upsample(v)) feeding a decoder: out[p'] = base[p'] + sum over the 2 x 2 block of R / s * [cat > 0]
__global__ void post_sumpool_kernel(const float* __restrict__ R, const unsigned* __restrict__ bits, const __half* __restrict__ cat, int pitch,
                                    int coff, const float* __restrict__ base, float* __restrict__ out, int C, int B, int h, int w) {
  float inv;
  grad_scale(*bits, &inv);
  const size_t total = (size_t)B * h * w * C;      // h, w: the COARSE size
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); size_t p = i / C;
    const int x = (int)(p % w); p /= w;
    const int y = (int)(p % h); const size_t b = p / h;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t q = (b * 2 * h + 2 * y + (k >> 1)) * (size_t)(2 * w) + 2 * x + (k & 1);
      if (__half2float(cat[q * 2 * pitch + coff + c]) > 0.f) acc += R[q * C + c];
    }
    out[i] = (base ? base[i] : 0.f) + acc * inv;
  }
}

// backward of maxpool2x2: the gradient of a pooled element goes to the first maximum of its window in scan order
// (torch max_pool2d); raw = the saved conv output [B][2h][2w][C], gp = [B][h][w][C]
__global__ void unpool_kernel(const float* __restrict__ gp, const float* __restrict__ raw, float* __restrict__ out, int C, int B, int h, int w) {
  const size_t total = (size_t)B * h * w * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); size_t p = i / C;
    const int x = (int)(p % w); p /= w;
    const int y = (int)(p % h); const size_t b = p / h;
    size_t q[4];
    float best = 0.f; int arg = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      q[k] = ((b * 2 * h + 2 * y + (k >> 1)) * (size_t)(2 * w) + 2 * x + (k & 1)) * C + c;
      const float v = raw[q[k]];
      if (k == 0 || v > best || (v != v && best == best)) { best = v; arg = k; }
    }
    const float g = gp[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) out[q[k]] = k == arg ? g : 0.f;
  }
}

// W'[j][co][ky][kx] = W[co][ci0 + j][2 - ky][2 - kx], j < n: the weights of the data-gradient convolution (OIHW again)
__global__ void flip_weights_kernel(const float* __restrict__ w, int cout, int cin, int ci0, int n, float* __restrict__ out) {
  const int total = n * cout * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, co = (i / 9) % cout, j = i / (9 * cout);
    out[i] = w[((size_t)co * cin + ci0 + j) * 9 + (8 - tap)];
  }
}

__global__ void add_kernel(float* __restrict__ a, const float* __restrict__ b, size_t n) {       // a += b, n % 4 == 0
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<float4*>(a)[i];
    const float4 y = reinterpret_cast<const float4*>(b)[i];
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    reinterpret_cast<float4*>(a)[i] = x;
  }
}

__global__ void zero_f32_kernel(float* __restrict__ p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

// ------------------------------------------------------------------------------ weight gradient on tcgen05
constexpr int kWgThreads = 256;
template <int N>
struct WgCfg {
  static constexpr int kABytes = 128 * 64 * 2;                      // [128 co][64 px] fp16
  static constexpr int kBBytes = N * 64 * 2;                        // [N ci][64 px] fp16, one plane of one tap
  static constexpr int kStageBytes = 2 * kABytes + 6 * kBBytes;     // A_hi | A_lo | 3 x (B_hi | B_lo)
  static constexpr int kStages = N == 64 ? 2 : 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

struct WgArgs {
  float* dw;               // [cout][cin][3][3], accumulated
  const unsigned* gbits;   // abs-max bits of dy (its scale)
  int cout, cin;
  int Wp;                  // padded row pitch W + 8 (a multiple of 8 pixels: TMA boxes start on 16-byte boundaries)
  int n_kblocks, kb_per_split, n_splits, n_ci_chunks, n_co_tiles;
};

template <int N>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x, const WgArgs a) {
  using Cfg = WgCfg<N>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTAs that run at the same time share a pixel range (split is the slowest index), so the operand tiles they all read
  // come from L2 once instead of from HBM once per (ky, channel chunk)
  int bid = blockIdx.x;
  const int ky = bid % 3; bid /= 3;
  const int cic = bid % a.n_ci_chunks; bid /= a.n_ci_chunks;
  const int cot = bid % a.n_co_tiles;
  const int split = bid / a.n_co_tiles;
  const int kb0 = split * a.kb_per_split;
  const int kb1 = min(a.n_kblocks, kb0 + a.kb_per_split);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_g); tma_prefetch_desc(&tmap_x); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t* st = smem + stage * Cfg::kStageBytes;
        mbar_expect_tx(full + stage, Cfg::kStageBytes);
        const int p0 = kb * 64;
        tma_load_3d(st, &tmap_g, full + stage, p0, cot * 128, 0);
        tma_load_3d(st + Cfg::kABytes, &tmap_g, full + stage, p0, cot * 128, 1);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {          // copy kx of x is pre-shifted by kx - 1 pixels; ky is a whole padded row
          uint8_t* bt = st + 2 * Cfg::kABytes + kx * 2 * Cfg::kBBytes;
          tma_load_3d(bt, &tmap_x, full + stage, p0 + (ky - 1) * a.Wp, cic * N, 2 * kx);
          tma_load_3d(bt + Cfg::kBBytes, &tmap_x, full + stage, p0 + (ky - 1) * a.Wp, cic * N, 2 * kx + 1);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc1 = umma_idesc_f16(N);          // dy_lo x x_hi
    constexpr uint32_t idesc2 = umma_idesc_f16(2 * N);      // dy_hi x [x_hi ; x_lo]
    const uint64_t desc_a_hi = umma_desc_sw128(smem_u32(smem));
    const uint64_t desc_a_lo = umma_desc_sw128(smem_u32(smem) + Cfg::kABytes);
    const uint64_t desc_b = umma_desc_sw128(smem_u32(smem) + 2 * Cfg::kABytes);
    int stage = 0; uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(full + stage, phase);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t so = (uint64_t)((uint32_t)(stage * Cfg::kStageBytes) >> 4);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          // accumulator columns of this tap: [0, N) = hi * hi, [N, 2N) = hi * lo + lo * hi
          const uint32_t acc = tmem_base + kx * 2 * N;
          const uint64_t db = desc_b + so + (uint64_t)((uint32_t)(kx * 2 * Cfg::kBBytes) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(acc, desc_a_hi + so + 2 * k, db + 2 * k, idesc2, (kb != kb0 || k != 0) ? 1u : 0u);
            umma_f16(acc + N, desc_a_lo + so + 2 * k, db + 2 * k, idesc1, 1u);
          }
        }
        umma_commit(empty + stage);
        if (kb == kb1 - 1) umma_commit(tmem_full);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float inv;
    grad_scale(*a.gbits, &inv);
    const int co = cot * 128 + q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int CH = N >= 32 ? 32 : 16;
#pragma unroll 1
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += CH) {
        uint32_t r0[32], r1[32];
        if (CH == 32) { tmem_ld_x32(taddr + kx * 2 * N + c0, r0); tmem_ld_x32(taddr + kx * 2 * N + N + c0, r1); }
        else { tmem_ld_x16(taddr + kx * 2 * N + c0, r0); tmem_ld_x16(taddr + kx * 2 * N + N + c0, r1); }
        tmem_ld_wait();
        if (co < a.cout) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int ci = cic * N + c0 + j;
            if (ci < a.cin) {
              const float v = fmaf(__uint_as_float(r1[j]), kLoInvScale, __uint_as_float(r0[j])) * inv;
              atomicAdd(a.dw + (((size_t)co * a.cin + ci) * 3 + ky) * 3 + kx, v);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_512(tmem_base); }
}

static int make_kmajor_map(CUtensorMap* m, const __half* base, size_t Ppad, int rows, int box_rows, int planes) {
  cuuint64_t dims[3] = {(cuuint64_t)Ppad, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)Ppad * 2, (cuuint64_t)Ppad * rows * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  return encode_cached(m, 3, base, dims, strides, box, "cuTensorMapEncodeTiled(wgrad operand)");
}

template <int N>
static int launch_wgrad(const __half* gT, int cout_pad, const __half* xT, int cin_pad, size_t Ppad, const WgArgs& a0, cudaStream_t st) {
  using Cfg = WgCfg<N>;
  auto kern = wgrad_tc_kernel<N>;
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  HA_CUDA_TRY(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    HA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    configured.fetch_or(bit, std::memory_order_release);
  }
  CUtensorMap tg, tx;
  int rc = make_kmajor_map(&tg, gT, Ppad, cout_pad, 128, 2);
  if (rc != HA_OK) return rc;
  if ((rc = make_kmajor_map(&tx, xT, Ppad, cin_pad, N, 6)) != HA_OK) return rc;
  WgArgs a = a0;
  a.n_kblocks = (int)(Ppad / 64);
  a.n_ci_chunks = cin_pad / N;
  a.n_co_tiles = cout_pad / 128;
  const int units = (cout_pad / 128) * a.n_ci_chunks * 3;
  int splits = (2 * kNumSMs + units - 1) / units;               // about two waves of CTAs
  if (splits > a.n_kblocks) splits = a.n_kblocks;
  if (splits < 1) splits = 1;
  a.kb_per_split = (a.n_kblocks + splits - 1) / splits;
  a.n_splits = (a.n_kblocks + a.kb_per_split - 1) / a.kb_per_split;   // every split owns at least one k-block
  kern<<<units * a.n_splits, kWgThreads, Cfg::kSmemBytes, st>>>(tg, tx, a);
  count_launches(1);
  return check_launch("wgrad_tc_kernel");
}

static inline int ew_grid(size_t total) {
  size_t b = (total + 255) / 256;
  return (int)(b < (size_t)kNumSMs * 16 ? (b ? b : 1) : (size_t)kNumSMs * 16);
}
static inline size_t padded_pixels(int B, int H, int W) { return align_up((size_t)B * (H + 2) * (W + kRowPad), 64); }

// packed weights of one data-gradient convolution (same layout as ha_conv3x3_nhwc's single_layout)
static PackedConv dgrad_layout(int cin, int cout, size_t* total) {
  PackedConv p;
  p.cin_pad = (int)align_up(cin, 64);
  p.cout_pad = (int)align_up(cout, 16);
  size_t off = 0;
  p.f32 = off; off = align_up(off + (size_t)9 * cin * cout * 4, 256);
  p.hi = off; off = align_up(off + (size_t)9 * p.cout_pad * p.cin_pad * 2, 256);
  p.lo = off; off = align_up(off + (size_t)9 * p.cout_pad * p.cin_pad * 2, 256);
  p.bias = off; off = align_up(off + (size_t)cout * 4, 256);
  *total = off;
  return p;
}

struct BwdWs {
  float* buf[6];          // gradient tensors, each up to B H W 64 floats
  __half* hs;             // scaled hi / lo planes of the current dy: [P][2][C]
  __half* gT;             // dy transposed: [2][CoutPad][Ppad]
  __half* xT;             // x transposed:  [2][CinPad][Ppad]
  float* wflip;           // OIHW scratch of a flipped weight slice
  char* wpack;            // packed weights of the current data-gradient convolution
  unsigned* bits;         // one abs-max word per layer (16)
  size_t total;
};
static BwdWs bwd_carve(void* ws, int B, int H, int W, int n_levels = 3) {
  BwdWs r;
  char* p = reinterpret_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* q = p + off; off = align_up(off + bytes, 256); return q; };
  const size_t px1 = (size_t)B * H * W, P1 = padded_pixels(B, H, W);
  for (int i = 0; i < 6; ++i) r.buf[i] = reinterpret_cast<float*>(take(px1 * 64 * sizeof(float)));
  const size_t P2 = padded_pixels(B, H / 2, W / 2), P4 = padded_pixels(B, H / 4, W / 4);
  auto max3 = [](size_t a, size_t b, size_t c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); };
  r.hs = reinterpret_cast<__half*>(take(px1 * 2 * 64 * sizeof(__half)));
  r.gT = reinterpret_cast<__half*>(take(max3(P1 * 128, P2 * 128, P4 * 256) * 2 * sizeof(__half)));   // dy^T: CoutPad rows per plane
  // x^T: 3 shifts x 2 planes x CinPad rows (level 4 adds cat3: 128 channels at full resolution)
  r.xT = reinterpret_cast<__half*>(take(max3(P1 * (n_levels == 4 ? 128 : 64), P2 * 192, P4 * 384) * 6 * sizeof(__half)));
  r.wflip = reinterpret_cast<float*>(take((size_t)256 * 256 * 9 * sizeof(float)));
  size_t wp;
  dgrad_layout(256, 256, &wp);
  r.wpack = take(wp);
  r.bits = reinterpret_cast<unsigned*>(take(16 * sizeof(unsigned)));
  r.total = off;
  return r;
}

// The three per-layer building blocks of the backward pass, on the caller's workspace
struct BwdOps {
  int B; cudaStream_t st; BwdWs w; int layer_no;
  // prepare dy of a layer: abs-max -> scaled hi / lo planes (data gradient) and their transpose (weight gradient); bias gradient
  const unsigned* prep(const float* gy, int C, int h, int wd, float* db) {
    unsigned* bits = w.bits + (layer_no++);
    const size_t n_px = (size_t)B * h * wd;
    absmax_kernel<<<ew_grid(n_px * C / 4), 256, 0, st>>>(gy, n_px * C, bits);
    split_scaled_kernel<<<ew_grid(n_px * C / 4), 256, 0, st>>>(gy, bits, w.hs, C, n_px);
    const int cpad = (int)align_up(C, 128);
    const size_t Pp = padded_pixels(B, h, wd);
    transpose_pad_kernel<1><<<dim3((unsigned)(Pp / 64), cpad / 64, 2), 256, 0, st>>>(w.hs, C, 0, C, B, h, wd, w.gT, cpad, Pp);
    count_launches(3);
    if (db) {
      zero_f32_kernel<<<1, 256, 0, st>>>(db, C);
      int blocks = (int)((n_px + 255) / 256);
      if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
      colsum_kernel<<<blocks, C, 0, st>>>(gy, n_px, C, db);
      count_launches(2);
    }
    return bits;
  }
  // weight gradient dw [cout][cin][3][3] of a layer whose input x is stored as hi / lo planes (channel slice of a concat buffer)
  int wgrad(int cin, int cout, float* dw, const __half* x, int pitch, int coff, int h, int wd, const unsigned* bits) {
    if (!dw) return HA_OK;
    const int cin_pad = (int)align_up(cin, 64), cout_pad = (int)align_up(cout, 128);
    const size_t Pp = padded_pixels(B, h, wd);
    transpose_pad_kernel<3><<<dim3((unsigned)(Pp / 64), cin_pad / 64, 2), 256, 0, st>>>(x, pitch, coff, cin, B, h, wd, w.xT, cin_pad, Pp);
    zero_f32_kernel<<<ew_grid((size_t)cout * cin * 9), 256, 0, st>>>(dw, (size_t)cout * cin * 9);
    count_launches(2);
    WgArgs a{};
    a.dw = dw; a.gbits = bits; a.cout = cout; a.cin = cin; a.Wp = wd + kRowPad;
    return launch_wgrad<64>(w.gT, cout_pad, w.xT, cin_pad, Pp, a, st);
  }
  // data gradient restricted to the input channels [ci0, ci0 + n): R = conv3x3(dy * s, W flipped) (still scaled by s)
  int dgrad(const float* w_oihw, int cin, int cout, int ci0, int n, int h, int wd, float* R) {
    flip_weights_kernel<<<ew_grid((size_t)n * cout * 9), 256, 0, st>>>(w_oihw, cout, cin, ci0, n, w.wflip);
    size_t wp;
    const PackedConv pc = dgrad_layout(cout, n, &wp);
    pack_conv_kernel<<<64, 256, 0, st>>>(w.wflip, nullptr, cout, n, pc.cin_pad, pc.cout_pad, reinterpret_cast<float*>(w.wpack + pc.f32),
                                         reinterpret_cast<__half*>(w.wpack + pc.hi), reinterpret_cast<__half*>(w.wpack + pc.lo),
                                         reinterpret_cast<float*>(w.wpack + pc.bias));
    count_launches(2);
    TcOut o;
    o.feat = R;
    return conv_tc(w.hs, cout, 0, cout, w.wpack, pc, n, false, o, B, h, wd, true, st, 9, true);
  }
};

}  // namespace ha

extern "C" size_t ha_vgg_backward_workspace_bytes(int B, int H, int W, int n_levels) {
  if (B <= 0 || H <= 0 || W <= 0 || (n_levels != 3 && n_levels != 4)) return 0;
  return ha::bwd_carve(nullptr, B, H, W, n_levels).total;
}

extern "C" int ha_vgg_backward(const HaVggStateDict* sd, const float* img_nchw, int B, int H, int W, int n_levels, void* fwd_ws,
                               const float* const* g_feat, const HaVggGrads* grads, void* ws, size_t ws_bytes, void* stream) {
  using namespace ha;
  if (!sd || !img_nchw || !fwd_ws || !g_feat || !grads || !ws) return HA_EINVAL;
  if ((n_levels != 3 && n_levels != 4) || B <= 0 || (W % 64) || (H % 32)) return HA_EINVAL;
  for (int l = 0; l < n_levels; ++l)
    if (!g_feat[l]) return HA_EINVAL;
  const BwdWs w = bwd_carve(ws, B, H, W, n_levels);
  if (ws_bytes < w.total) return HA_ENOSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const TcSaved sv = vgg_train_saved(reinterpret_cast<char*>(fwd_ws), B, H, W, n_levels);
  const size_t px1 = (size_t)B * H * W, px2 = px1 / 4, px4 = px1 / 16;
  HA_CUDA_TRY(cudaMemsetAsync(w.bits, 0, 16 * sizeof(unsigned), st));
  int rc;
#define HA_TRY(x) do { rc = (x); if (rc != HA_OK) return rc; } while (0)

  // ---- building blocks (BwdOps) bound to this network's layer table ---------------------------------------------------
  BwdOps ops{B, st, w, 0};
  auto prep = [&](const float* gy, int C, int h, int wd, int li) -> const unsigned* {
    return ops.prep(gy, C, h, wd, kVggConvs[li].has_bias ? grads->bias[li] : nullptr);
  };
  auto wgrad = [&](int li, const __half* x, int pitch, int coff, int h, int wd, const unsigned* bits) -> int {
    return ops.wgrad(kVggConvs[li].cin, kVggConvs[li].cout, grads->weight[li], x, pitch, coff, h, wd, bits);
  };
  auto dgrad = [&](int li, int ci0, int n, int h, int wd, float* R) -> int {
    return ops.dgrad(sd->weight[li], kVggConvs[li].cin, kVggConvs[li].cout, ci0, n, h, wd, R);
  };
  auto mask = [&](const float* R, const unsigned* bits, const __half* act, int pitch, int coff, const float* add, float* out, int C,
                  size_t n_px) {
    post_mask_kernel<<<ew_grid(n_px * C / 4), 256, 0, st>>>(R, bits, act, pitch, coff, add, out, C, n_px);
    count_launches(1);
  };
  auto sumpool = [&](const float* R, const unsigned* bits, const __half* cat, int pitch, int coff, const float* base, float* out, int C,
                     int h, int wd) {       // h, wd: coarse size
    post_sumpool_kernel<<<ew_grid((size_t)B * h * wd * C), 256, 0, st>>>(R, bits, cat, pitch, coff, base, out, C, B, h, wd);
    count_launches(1);
  };
  auto unpool = [&](const float* gp, const float* raw, float* out, int C, int h, int wd) {   // h, wd: pooled size
    unpool_kernel<<<ew_grid((size_t)B * h * wd * C), 256, 0, st>>>(gp, raw, out, C, B, h, wd);
    count_launches(1);
  };
  float *b0 = w.buf[0], *b1 = w.buf[1], *b2 = w.buf[2], *b3 = w.buf[3], *b4 = w.buf[4], *b5 = w.buf[5];
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
  const unsigned* bits;

  // ---- decoder 3 (level 4 only, VGG.py:154-157): x24 = dec3b(d3), d3 = relu(dec3a(cat3)), cat3 = [relu(up(x21)) | relu(x2)] ----
  const float* g21 = g_feat[2];
  if (n_levels == 4) {
    bits = prep(g_feat[3], 16, H, W, L_DEC3B);
    HA_TRY(wgrad(L_DEC3B, sv.d3, 32, 0, H, W, bits));
    HA_TRY(dgrad(L_DEC3B, 0, 32, H, W, b0));
    mask(b0, bits, sv.d3, 32, 0, nullptr, b1, 32, px1);
    bits = prep(b1, 32, H, W, L_DEC3A);
    HA_TRY(wgrad(L_DEC3A, sv.cat3, 128, 0, H, W, bits));
    HA_TRY(dgrad(L_DEC3A, 0, 64, H, W, b0));
    sumpool(b0, bits, sv.cat3, 128, 0, g_feat[2], b4, 64, H2, W2);                     // b4 = total gradient of x21 (b4 is free until decoder 1)
    HA_TRY(dgrad(L_DEC3A, 64, 64, H, W, b0));
    mask(b0, bits, sv.cat3, 128, 64, nullptr, b5, 64, px1);                            // b5 = gradient of x2 through the relu(x2) skip
    g21 = b4;
  }
  // ---- decoder 2 (VGG.py:149-152): x21 = dec2b(d2), d2 = relu(dec2a(cat2)), cat2 = [relu(up(x18)) | x4] ---------------
  bits = prep(g21, 64, H2, W2, L_DEC2B);
  HA_TRY(wgrad(L_DEC2B, sv.d2, 64, 0, H2, W2, bits));
  HA_TRY(dgrad(L_DEC2B, 0, 64, H2, W2, b0));
  mask(b0, bits, sv.d2, 64, 0, nullptr, b1, 64, px2);                                  // b1 = d loss / d (dec2a output)
  bits = prep(b1, 64, H2, W2, L_DEC2A);
  HA_TRY(wgrad(L_DEC2A, sv.cat2, 192, 0, H2, W2, bits));
  HA_TRY(dgrad(L_DEC2A, 0, 128, H2, W2, b0));
  sumpool(b0, bits, sv.cat2, 192, 0, g_feat[1], b2, 128, H4, W4);                      // b2 = total gradient of x18
  HA_TRY(dgrad(L_DEC2A, 128, 64, H2, W2, b0));
  mask(b0, bits, sv.cat2, 192, 128, nullptr, b3, 64, px2);                             // b3 = gradient of x4 through the skip
  // ---- decoder 1 (:144-147): x18 = dec1b(d1), d1 = relu(dec1a(cat1)), cat1 = [relu(up(x15)) | x9] ----------------------
  bits = prep(b2, 128, H4, W4, L_DEC1B);
  HA_TRY(wgrad(L_DEC1B, sv.d1, 128, 0, H4, W4, bits));
  HA_TRY(dgrad(L_DEC1B, 0, 128, H4, W4, b0));
  mask(b0, bits, sv.d1, 128, 0, nullptr, b1, 128, px4);
  bits = prep(b1, 128, H4, W4, L_DEC1A);
  HA_TRY(wgrad(L_DEC1A, sv.cat1, 384, 0, H4, W4, bits));
  HA_TRY(dgrad(L_DEC1A, 0, 256, H4, W4, b0));
  sumpool(b0, bits, sv.cat1, 384, 0, g_feat[0], b2, 256, H8, W8);                      // b2 = total gradient of x15
  HA_TRY(dgrad(L_DEC1A, 256, 128, H4, W4, b0));
  mask(b0, bits, sv.cat1, 384, 256, nullptr, b4, 128, px4);                            // b4 = gradient of x9 through the skip
  // ---- encoder block 2 (:137-141): x15 = pool(conv14(relu(conv12(relu(conv10(x9)))))) -----------------------------------
  unpool(b2, sv.x14, b1, 256, H8, W8);
  bits = prep(b1, 256, H4, W4, L_CONV14);
  HA_TRY(wgrad(L_CONV14, sv.a12, 256, 0, H4, W4, bits));
  HA_TRY(dgrad(L_CONV14, 0, 256, H4, W4, b0));
  mask(b0, bits, sv.a12, 256, 0, nullptr, b2, 256, px4);
  bits = prep(b2, 256, H4, W4, L_CONV12);
  HA_TRY(wgrad(L_CONV12, sv.a10, 256, 0, H4, W4, bits));
  HA_TRY(dgrad(L_CONV12, 0, 256, H4, W4, b0));
  mask(b0, bits, sv.a10, 256, 0, nullptr, b1, 256, px4);
  bits = prep(b1, 256, H4, W4, L_CONV10);
  HA_TRY(wgrad(L_CONV10, sv.cat1, 384, 256, H4, W4, bits));
  HA_TRY(dgrad(L_CONV10, 0, 128, H4, W4, b0));
  mask(b0, bits, sv.cat1, 384, 256, b4, b2, 128, px4);                                 // b2 = total gradient of x9 = relu(pool(x7))
  // ---- encoder block 1 (:129-134) -------------------------------------------------------------------------------
  unpool(b2, sv.x7, b1, 128, H4, W4);
  bits = prep(b1, 128, H2, W2, L_CONV7);
  HA_TRY(wgrad(L_CONV7, sv.a5, 128, 0, H2, W2, bits));
  HA_TRY(dgrad(L_CONV7, 0, 128, H2, W2, b0));
  mask(b0, bits, sv.a5, 128, 0, nullptr, b2, 128, px2);
  bits = prep(b2, 128, H2, W2, L_CONV5);
  HA_TRY(wgrad(L_CONV5, sv.cat2, 192, 128, H2, W2, bits));
  HA_TRY(dgrad(L_CONV5, 0, 64, H2, W2, b0));
  mask(b0, bits, sv.cat2, 192, 128, b3, b1, 64, px2);                                  // b1 = total gradient of x4 = relu(pool(x2))
  // ---- encoder block 0 (:123-128) -------------------------------------------------------------------------------
  unpool(b1, sv.x2, b2, 64, H2, W2);
  if (n_levels == 4) {                                                                   // + the skip into decoder 3
    add_kernel<<<ew_grid(px1 * 16), 256, 0, st>>>(b2, b5, px1 * 64);
    count_launches(1);
  }
  bits = prep(b2, 64, H, W, L_CONV2);
  HA_TRY(wgrad(L_CONV2, sv.a1, 64, 0, H, W, bits));
  HA_TRY(dgrad(L_CONV2, 0, 64, H, W, b0));
  mask(b0, bits, sv.a1, 64, 0, nullptr, b1, 64, px1);
  bits = prep(b1, 64, H, W, L_CONV0);
  if (grads->weight[L_CONV0]) {
    float* dw = grads->weight[L_CONV0];
    const size_t Pp = padded_pixels(B, H, W);
    image_pad_kernel<<<ew_grid(3 * Pp), 256, 0, st>>>(img_nchw, B, H, W, w.xT, Pp);
    zero_f32_kernel<<<ew_grid(64 * 3 * 9), 256, 0, st>>>(dw, 64 * 3 * 9);
    count_launches(2);
    WgArgs a{};
    a.dw = dw; a.gbits = bits; a.cout = 64; a.cin = 3; a.Wp = W + kRowPad;
    HA_TRY(launch_wgrad<16>(w.gT, 128, w.xT, 16, Pp, a, st));
  }
#undef HA_TRY
  return check_launch("ha_vgg_backward");
}

// ---- one convolution layer backward (the twin of ha_conv3x3_nhwc; the building blocks of ha_vgg_backward on their own)
extern "C" size_t ha_conv3x3_backward_workspace_bytes(int cin, int cout, int B, int H, int W) {
  if (cin <= 0 || cout <= 0 || cin > 384 || cout > 256 || B <= 0 || H <= 0 || W <= 0) return 0;
  // the U-Net carve-up is sized for 64 channels at H x W, 192 at half and 384 at quarter resolution: ask for a canvas on which
  // this layer's channel count fits
  const int scale = cin > 192 || cout > 128 ? 4 : (cin > 64 || cout > 64 ? 2 : 1);
  return ha::bwd_carve(nullptr, B, H * scale, W * scale).total + ha::align_up((size_t)B * H * W * 2 * cin * 2, 256);
}

extern "C" int ha_conv3x3_backward_nhwc(const float* x_nhwc, int cin, const float* w_oihw, const float* dy_nhwc, int cout, int B, int H,
                                        int W, float* dx_nhwc, float* dw_oihw, float* db, void* ws, size_t ws_bytes, void* stream) {
  using namespace ha;
  if (!x_nhwc || !w_oihw || !dy_nhwc || !ws || (cin % 64) || (cout % 64) || (W % 16) || (H % 16)) return HA_EINVAL;
  const size_t need = ha_conv3x3_backward_workspace_bytes(cin, cout, B, H, W);
  if (!need || ws_bytes < need) return HA_ENOSPACE;
  const int scale = cin > 192 || cout > 128 ? 4 : (cin > 64 || cout > 64 ? 2 : 1);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const BwdWs w = bwd_carve(ws, B, H * scale, W * scale);
  __half* xs = reinterpret_cast<__half*>(reinterpret_cast<char*>(ws) + w.total);
  const size_t n_px = (size_t)B * H * W;
  HA_CUDA_TRY(cudaMemsetAsync(w.bits, 0, 16 * sizeof(unsigned), st));
  split_act_kernel<<<kNumSMs * 8, 256, 0, st>>>(x_nhwc, xs, cin, n_px);
  count_launches(1);
  BwdOps ops{B, st, w, 0};
  const unsigned* bits = ops.prep(dy_nhwc, cout, H, W, db);
  int rc = ops.wgrad(cin, cout, dw_oihw, xs, cin, 0, H, W, bits);
  if (rc != HA_OK) return rc;
  if (dx_nhwc) {
    if (cin != 64 && cin != 128 && cin != 256) return HA_EINVAL;   // other widths (the concat inputs) are split by ha_vgg_backward's schedule
    if ((rc = ops.dgrad(w_oihw, cin, cout, 0, cin, H, W, w.buf[0])) != HA_OK) return rc;
    post_mask_kernel<<<ew_grid(n_px * cin / 4), 256, 0, st>>>(w.buf[0], bits, nullptr, 0, 0, nullptr, dx_nhwc, cin, n_px);
    count_launches(1);
  }
  return check_launch("ha_conv3x3_backward_nhwc");
}
