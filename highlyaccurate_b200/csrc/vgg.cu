// VGG16 U-Net feature extractor (VGG.py:13-203): weight packing, the CUDA-core fp32 convolution
// back end (validation path, HA_CONV_FP32_SIMT), the shared elementwise kernels (2x2 max-pool,
// nearest x2 upsample into concat slices, L2-norm scale, confidence heads) and the layer
// schedule.  The tensor-core back end lives in vgg_tc.cu and plugs into the same schedule.
//
// Data layout: every activation is NHWC; skip connections are not concatenated by a kernel —
// producers write straight into channel slices of the decoder's input buffer
// (cat1 = [up(x15) | x9], cat2 = [up(x18) | x4], cat3 = [up(x21) | x2]; VGG.py:144-155).
// ReLU commutes with max-pool and nearest upsampling, so the in-place ReLU quirks of the
// reference (VGG.py:126-128: the skip tensors are post-ReLU) are reproduced by storing
// relu(conv) once and pooling / upsampling that.
#include "vgg_common.cuh"

namespace ha {

// ---------------------------------------------------------------------------- weight packing
// src: OIHW fp32 [Cout][Cin][3][3] (torch Conv2d.weight).
__global__ void pack_conv_kernel(const float* __restrict__ w, const float* __restrict__ bias, int cin, int cout,
                                 int cin_pad, int cout_pad, float* __restrict__ f32, __half* __restrict__ hi,
                                 __half* __restrict__ lo, float* __restrict__ bias_out) {
  const int n_pad = 9 * cout_pad * cin_pad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
    const int ci = i % cin_pad, co = (i / cin_pad) % cout_pad, tap = i / (cin_pad * cout_pad);
    float v = 0.f;
    if (ci < cin && co < cout) v = w[((size_t)co * cin + ci) * 9 + tap];
    const __half h = __float2half_rn(v);
    const float rem = (v - __half2float(h)) * kLoScale;
    hi[i] = h;
    lo[i] = __float2half_rn(rem);
    if (ci < cin && co < cout) f32[((size_t)tap * cin + ci) * cout + co] = v;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cout; i += gridDim.x * blockDim.x)
    bias_out[i] = bias ? bias[i] : 0.f;
}

// ---------------------------------------------------------------------------- fp32 direct conv
// 3x3, pad 1, stride 1, NHWC in / NHWC out, both addressed as (pixel pitch, channel offset) so
// they can be channel slices of a concat buffer.  CTA tile: 16 rows x 8 cols x 64 output
// channels, 256 threads, each thread 8 pixels (one tile row) x 4 channels.
constexpr int kSimtTH = 16, kSimtTW = 8, kSimtTN = 64, kSimtKC = 8;

struct SimtConvArgs {
  const float* in; int in_pitch, in_coff, cin;
  const float* w;      // [9][cin][cout]
  const float* bias;   // [cout] or null
  float* out; int out_pitch, out_coff, cout;
  int B, H, W, relu_out;
};

__global__ void __launch_bounds__(256) conv3x3_simt_kernel(const SimtConvArgs a) {
  __shared__ float in_s[kSimtTH + 2][kSimtTW + 2][kSimtKC];
  __shared__ __align__(16) float w_s[9][kSimtKC][kSimtTN];
  const int tiles_x = (a.W + kSimtTW - 1) / kSimtTW;
  const int tx0 = (blockIdx.x % tiles_x) * kSimtTW, ty0 = (blockIdx.x / tiles_x) * kSimtTH;
  const int n0 = blockIdx.y * kSimtTN, b = blockIdx.z;
  const int tc = threadIdx.x & 15, tr = threadIdx.x >> 4;     // channel group (4 ch), tile row
  float acc[8][4];
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;

  const float* in_b = a.in + (size_t)b * a.H * a.W * a.in_pitch + a.in_coff;
  for (int k0 = 0; k0 < a.cin; k0 += kSimtKC) {
    // stage the input halo tile (zero outside the image / beyond cin)
    for (int i = threadIdx.x; i < (kSimtTH + 2) * (kSimtTW + 2) * kSimtKC; i += 256) {
      const int k = i % kSimtKC, x = (i / kSimtKC) % (kSimtTW + 2), y = i / (kSimtKC * (kSimtTW + 2));
      const int gy = ty0 + y - 1, gx = tx0 + x - 1;
      float v = 0.f;
      if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && k0 + k < a.cin)
        v = in_b[((size_t)gy * a.W + gx) * a.in_pitch + k0 + k];
      in_s[y][x][k] = v;
    }
    for (int i = threadIdx.x; i < 9 * kSimtKC * kSimtTN; i += 256) {
      const int n = i % kSimtTN, k = (i / kSimtTN) % kSimtKC, tap = i / (kSimtTN * kSimtKC);
      float v = 0.f;
      if (k0 + k < a.cin && n0 + n < a.cout) v = a.w[((size_t)tap * a.cin + k0 + k) * a.cout + n0 + n];
      w_s[tap][k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int k = 0; k < kSimtKC; ++k) {
          const float4 wv = *reinterpret_cast<const float4*>(&w_s[ky * 3 + kx][k][tc * 4]);
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const float v = in_s[tr + ky][p + kx][k];
            acc[p][0] += v * wv.x; acc[p][1] += v * wv.y; acc[p][2] += v * wv.z; acc[p][3] += v * wv.w;
          }
        }
    __syncthreads();
  }
  const int gy = ty0 + tr;
  if (gy >= a.H) return;
  float* out_b = a.out + (size_t)b * a.H * a.W * a.out_pitch + a.out_coff;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int gx = tx0 + p;
    if (gx >= a.W) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int n = n0 + tc * 4 + c;
      if (n >= a.cout) continue;
      float v = acc[p][c] + (a.bias ? a.bias[n] : 0.f);
      if (a.relu_out) v = fmaxf(v, 0.f);
      out_b[((size_t)gy * a.W + gx) * a.out_pitch + n] = v;
    }
  }
}

int conv_simt(const float* in, int in_pitch, int in_coff, int cin, const float* w, const float* bias, float* out,
                     int out_pitch, int out_coff, int cout, int B, int H, int W, int relu_out, cudaStream_t st) {
  SimtConvArgs a{in, in_pitch, in_coff, cin, w, bias, out, out_pitch, out_coff, cout, B, H, W, relu_out};
  dim3 grid(((W + kSimtTW - 1) / kSimtTW) * ((H + kSimtTH - 1) / kSimtTH), (cout + kSimtTN - 1) / kSimtTN, B);
  conv3x3_simt_kernel<<<grid, 256, 0, st>>>(a);
  count_launches(1);
  return check_launch("conv3x3_simt_kernel");
}

// ---------------------------------------------------------------------------- elementwise (fp32)
// 2x2 max-pool, stride 2 (VGG.py:126,134,141): in [B][H][W][pitch]+coff -> out [B][H/2][W/2][pitch]+coff
__global__ void pool2x2_kernel(const float* __restrict__ in, int in_pitch, int in_coff, float* __restrict__ out,
                               int out_pitch, int out_coff, int C, int H, int W, size_t total) {
  const int C4 = C / 4, Ho = H / 2, Wo = W / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    size_t p = i / C4;
    const int x = (int)(p % Wo); p /= Wo;
    const int y = (int)(p % Ho); const size_t b = p / Ho;
    const float* s = in + ((b * H + 2 * y) * W + 2 * x) * in_pitch + in_coff + c;
    const float4 v00 = *reinterpret_cast<const float4*>(s), v01 = *reinterpret_cast<const float4*>(s + in_pitch);
    const float4 v10 = *reinterpret_cast<const float4*>(s + (size_t)W * in_pitch);
    const float4 v11 = *reinterpret_cast<const float4*>(s + (size_t)W * in_pitch + in_pitch);
    float4 r;
    r.x = fmaxf(fmaxf(v00.x, v01.x), fmaxf(v10.x, v11.x)); r.y = fmaxf(fmaxf(v00.y, v01.y), fmaxf(v10.y, v11.y));
    r.z = fmaxf(fmaxf(v00.z, v01.z), fmaxf(v10.z, v11.z)); r.w = fmaxf(fmaxf(v00.w, v01.w), fmaxf(v10.w, v11.w));
    *reinterpret_cast<float4*>(out + ((b * Ho + y) * Wo + x) * out_pitch + out_coff + c) = r;
  }
}

// relu + nearest x2 upsample (VGG.py:144,149,154 F.interpolate mode='nearest' to exactly 2x) into
// a channel slice: in [B][h][w][C] -> out [B][2h][2w][pitch]+coff
__global__ void relu_up2x_kernel(const float* __restrict__ in, float* __restrict__ out, int out_pitch, int out_coff,
                                 int C, int h, int w, size_t total) {
  const int C4 = C / 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    size_t p = i / C4;
    const int x = (int)(p % (2 * w)); p /= (2 * w);
    const int y = (int)(p % (2 * h)); const size_t b = p / (2 * h);
    float4 v = *reinterpret_cast<const float4*>(in + ((b * h + y / 2) * w + x / 2) * C + c);
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    *reinterpret_cast<float4*>(out + ((b * 2 * h + y) * (size_t)(2 * w) + x) * out_pitch + out_coff + c) = v;
  }
}

// per-sample 1 / max(||x||_2, 1e-12) (VGG.py:511-514, F.normalize): two deterministic stages, all pyramid levels of a
// forward in one launch each (grid.z = level).
constexpr int kNormChunks = 64;
struct NormArgs { const float* x[HA_MAX_LEVELS]; float* scale[HA_MAX_LEVELS]; size_t n_per_sample[HA_MAX_LEVELS]; };
__global__ void sumsq_partial_kernel(const NormArgs a, int B, double* __restrict__ part) {
  const int b = blockIdx.y, l = blockIdx.z;
  if (!a.scale[l]) return;
  const size_t n_per_sample = a.n_per_sample[l];
  const float4* p = reinterpret_cast<const float4*>(a.x[l] + (size_t)b * n_per_sample);
  const size_t n4 = n_per_sample / 4;
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t lo = blockIdx.x * per, hi = min(n4, lo + per);
  float acc = 0.f;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float4 v = p[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  __shared__ double red[8];
  double r = warp_sum((double)acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = r;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    part[((size_t)l * B + b) * gridDim.x + blockIdx.x] = t;
  }
}
__global__ void norm_scale_kernel(const NormArgs a, const double* __restrict__ part, int chunks, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y;
  if (b >= B || !a.scale[l]) return;
  double t = 0;
  for (int i = 0; i < chunks; ++i) t += part[((size_t)l * B + b) * chunks + i];
  a.scale[l][b] = (float)(1.0 / fmax(sqrt(t), 1e-12));
}

// confidence head (VGG.py:62-81,160-163): sigmoid(-sigmoid(conv3x3_{C->1}(relu(x)))), no bias.
// One warp per output pixel, lanes across channels; memory-bound (N = 1).
__global__ void conf_head_kernel(const float* __restrict__ x, const float* __restrict__ w /*[9][C]*/, float* __restrict__ out,
                                 int C, int H, int W, size_t n_px) {
  const size_t px = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (px >= n_px) return;
  const int xx = (int)(px % W), yy = (int)((px / W) % H);
  const size_t b = px / ((size_t)W * H);
  float acc = 0.f;
  for (int ky = -1; ky <= 1; ++ky) {
    const int gy = yy + ky;
    if (gy < 0 || gy >= H) continue;
    for (int kx = -1; kx <= 1; ++kx) {
      const int gx = xx + kx;
      if (gx < 0 || gx >= W) continue;
      const float4* s = reinterpret_cast<const float4*>(x + ((b * H + gy) * W + gx) * C);
      const float4* wt = reinterpret_cast<const float4*>(w + ((ky + 1) * 3 + kx + 1) * C);
      for (int c = lane; c < C / 4; c += 32) {
        const float4 v = s[c], q = __ldg(wt + c);
        acc += fmaxf(v.x, 0.f) * q.x + fmaxf(v.y, 0.f) * q.y + fmaxf(v.z, 0.f) * q.z + fmaxf(v.w, 0.f) * q.w;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const float s1 = 1.f / (1.f + expf(-acc));
    out[px] = 1.f / (1.f + expf(s1));
  }
}

static inline int ew_blocks(size_t total) {
  size_t b = (total + 255) / 256;
  return (int)(b < (size_t)kNumSMs * 16 ? b : (size_t)kNumSMs * 16);
}

// fp32 CUDA-core schedule.  Buffers follow the names of VGG.py:121-158.
static int vgg_forward_simt(const char* packed, const PackedLayout& L, const float* img, int B, int H, int W, int n_levels,
                            float* const* out_feat, Arena& ar, cudaStream_t st) {
  const size_t px1 = (size_t)B * H * W, px2 = px1 / 4, px4 = px1 / 16;
  float* nhwc_img = (float*)ar.take(px1 * 3 * 4);
  float* a1 = (float*)ar.take(px1 * 64 * 4);
  float* cat3 = (float*)ar.take(px1 * 128 * 4);        // [up(relu(x21)) | relu(x2)]; x2 slice always used
  float* cat2 = (float*)ar.take(px2 * 192 * 4);        // [up(relu(x18)) | x4]
  float* a5 = (float*)ar.take(px2 * 128 * 4);
  float* x7 = (float*)ar.take(px2 * 128 * 4);
  float* cat1 = (float*)ar.take(px4 * 384 * 4);        // [up(relu(x15)) | x9]
  float* a10 = (float*)ar.take(px4 * 256 * 4);
  float* a12 = (float*)ar.take(px4 * 256 * 4);
  float* x14 = (float*)ar.take(px4 * 256 * 4);
  float* d1 = (float*)ar.take(px4 * 128 * 4);
  float* d2 = (float*)ar.take(px2 * 64 * 4);
  float* d3 = n_levels == 4 ? (float*)ar.take(px1 * 32 * 4) : nullptr;
  if (ar.dry) return HA_OK;
  if (ar.off > ar.cap) return HA_ENOSPACE;
  auto Wf = [&](int i) { return reinterpret_cast<const float*>(packed + L.c[i].f32); };
  auto Bs = [&](int i) { return kVggConvs[i].has_bias ? reinterpret_cast<const float*>(packed + L.c[i].bias) : nullptr; };
  int rc;
#define HA_TRY(x) do { rc = (x); if (rc != HA_OK) return rc; } while (0)
  HA_TRY(ha_nchw_to_nhwc(img, nhwc_img, B, 3, H, W, st));
  HA_TRY(conv_simt(nhwc_img, 3, 0, 3, Wf(L_CONV0), Bs(L_CONV0), a1, 64, 0, 64, B, H, W, 1, st));             // x1
  HA_TRY(conv_simt(a1, 64, 0, 64, Wf(L_CONV2), Bs(L_CONV2), cat3, 128, 64, 64, B, H, W, 1, st));             // relu(x2)
  pool2x2_kernel<<<ew_blocks(px2 * 16), 256, 0, st>>>(cat3, 128, 64, cat2, 192, 128, 64, H, W, px2 * 16); count_launches(1);     // x4
  HA_TRY(conv_simt(cat2, 192, 128, 64, Wf(L_CONV5), Bs(L_CONV5), a5, 128, 0, 128, B, H / 2, W / 2, 1, st));  // x6
  HA_TRY(conv_simt(a5, 128, 0, 128, Wf(L_CONV7), Bs(L_CONV7), x7, 128, 0, 128, B, H / 2, W / 2, 1, st));     // relu(x7)
  pool2x2_kernel<<<ew_blocks(px4 * 32), 256, 0, st>>>(x7, 128, 0, cat1, 384, 256, 128, H / 2, W / 2, px4 * 32); count_launches(1);  // x9
  HA_TRY(conv_simt(cat1, 384, 256, 128, Wf(L_CONV10), Bs(L_CONV10), a10, 256, 0, 256, B, H / 4, W / 4, 1, st));
  HA_TRY(conv_simt(a10, 256, 0, 256, Wf(L_CONV12), Bs(L_CONV12), a12, 256, 0, 256, B, H / 4, W / 4, 1, st));
  HA_TRY(conv_simt(a12, 256, 0, 256, Wf(L_CONV14), Bs(L_CONV14), x14, 256, 0, 256, B, H / 4, W / 4, 0, st)); // x14 (no relu)
  pool2x2_kernel<<<ew_blocks(px4 / 4 * 64), 256, 0, st>>>(x14, 256, 0, out_feat[0], 256, 0, 256, H / 4, W / 4, px4 / 4 * 64); count_launches(1);  // x15
  relu_up2x_kernel<<<ew_blocks(px4 * 64), 256, 0, st>>>(out_feat[0], cat1, 384, 0, 256, H / 8, W / 8, px4 * 64); count_launches(1);
  HA_TRY(conv_simt(cat1, 384, 0, 384, Wf(L_DEC1A), nullptr, d1, 128, 0, 128, B, H / 4, W / 4, 1, st));
  HA_TRY(conv_simt(d1, 128, 0, 128, Wf(L_DEC1B), nullptr, out_feat[1], 128, 0, 128, B, H / 4, W / 4, 0, st));   // x18
  relu_up2x_kernel<<<ew_blocks(px2 * 32), 256, 0, st>>>(out_feat[1], cat2, 192, 0, 128, H / 4, W / 4, px2 * 32); count_launches(1);
  HA_TRY(conv_simt(cat2, 192, 0, 192, Wf(L_DEC2A), nullptr, d2, 64, 0, 64, B, H / 2, W / 2, 1, st));
  HA_TRY(conv_simt(d2, 64, 0, 64, Wf(L_DEC2B), nullptr, out_feat[2], 64, 0, 64, B, H / 2, W / 2, 0, st));        // x21
  if (n_levels == 4) {
    relu_up2x_kernel<<<ew_blocks(px1 * 16), 256, 0, st>>>(out_feat[2], cat3, 128, 0, 64, H / 2, W / 2, px1 * 16); count_launches(1);
    HA_TRY(conv_simt(cat3, 128, 0, 128, Wf(L_DEC3A), nullptr, d3, 32, 0, 32, B, H, W, 1, st));
    HA_TRY(conv_simt(d3, 32, 0, 32, Wf(L_DEC3B), nullptr, out_feat[3], 16, 0, 16, B, H, W, 0, st));              // x24
  }
#undef HA_TRY
  return check_launch("vgg_forward_simt");
}

static int vgg_run(const char* packed, const float* img, int B, int H, int W, int n_levels, int precision,
                   float* const* out_feat, float* const* out_scale, float* const* out_conf, Arena& ar, cudaStream_t st,
                   TcSaved* saved = nullptr, bool g2s = false) {
  const PackedLayout L = vgg_packed_layout();
  double* norm_part = (double*)ar.take((size_t)HA_MAX_LEVELS * B * kNormChunks * sizeof(double));
  int rc;
  if (precision == HA_CONV_FP32_SIMT) rc = g2s ? HA_EINVAL : vgg_forward_simt(packed, L, img, B, H, W, n_levels, out_feat, ar, st);
  else rc = vgg_forward_tc(packed, L, img, B, H, W, n_levels, precision, out_feat, ar, st, saved, g2s);
  if (rc != HA_OK || ar.dry) return rc;
  static const int chans[4] = {256, 128, 64, 16};
  NormArgs na;
  bool any_scale = false;
  for (int l = 0; l < HA_MAX_LEVELS; ++l) {
    na.x[l] = nullptr; na.scale[l] = nullptr; na.n_per_sample[l] = 0;
    if (l < n_levels && out_scale && out_scale[l]) {
      na.x[l] = out_feat[l]; na.scale[l] = out_scale[l];
      na.n_per_sample[l] = (size_t)(H >> (3 - l)) * (W >> (3 - l)) * chans[l];
      any_scale = true;
    }
  }
  if (any_scale) {
    sumsq_partial_kernel<<<dim3(kNormChunks, B, n_levels), 256, 0, st>>>(na, B, norm_part);
    norm_scale_kernel<<<dim3((B + 127) / 128, n_levels), 128, 0, st>>>(na, norm_part, kNormChunks, B);
    count_launches(2);
  }
  for (int l = 0; l < n_levels; ++l) {
    int h = H >> (3 - l), w = W >> (3 - l);
    const int C = chans[l];
    if (g2s && l > 0) { h *= 2; w /= 2; }        // VGG.py:326-329: c0 is taken from the un-folded x15, c1.. from the folded decoder maps
    if (out_conf && out_conf[l]) {
      const size_t n_px = (size_t)B * h * w;
      conf_head_kernel<<<(unsigned)((n_px * 32 + 255) / 256), 256, 0, st>>>(
          out_feat[l], reinterpret_cast<const float*>(packed + L.c[L_CONF0 + l].f32), out_conf[l], C, h, w, n_px);
      count_launches(1);
    }
  }
  return check_launch("vgg_run");
}

// where a train-mode forward (ha_vgg_forward_train) left the activations in its workspace
TcSaved vgg_train_saved(char* ws, int B, int H, int W, int n_levels) {
  Arena ar{ws, 0, ~(size_t)0, false};
  ar.take((size_t)HA_MAX_LEVELS * B * kNormChunks * sizeof(double));
  return vgg_tc_carve(ar, B, H, W, n_levels, true);
}

}  // namespace ha

extern "C" size_t ha_vgg_packed_weight_bytes(void) { return ha::vgg_packed_layout().total; }

extern "C" int ha_vgg_pack_weights(const HaVggStateDict* sd, void* packed, size_t packed_bytes, void* stream) {
  if (!sd || !packed) return HA_EINVAL;
  const ha::PackedLayout L = ha::vgg_packed_layout();
  if (packed_bytes < L.total) return HA_ENOSPACE;
  char* base = reinterpret_cast<char*>(packed);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int i = 0; i < HA_VGG_N_CONV; ++i) {
    if (!sd->weight[i]) return HA_EINVAL;
    const ha::ConvSpec& s = ha::kVggConvs[i];
    const ha::PackedConv& p = L.c[i];
    ha::pack_conv_kernel<<<64, 256, 0, st>>>(sd->weight[i], s.has_bias ? sd->bias[i] : nullptr, s.cin, s.cout, p.cin_pad,
                                             p.cout_pad, reinterpret_cast<float*>(base + p.f32),
                                             reinterpret_cast<__half*>(base + p.hi), reinterpret_cast<__half*>(base + p.lo),
                                             reinterpret_cast<float*>(base + p.bias));
  }
  ha::count_launches(HA_VGG_N_CONV);
  return ha::check_launch("pack_conv_kernel");
}

static bool vgg_shape_ok(int B, int H, int W, int n_levels, int precision) {
  if (B <= 0 || H <= 0 || W <= 0 || (H % 8) || (W % 8)) return false;
  if (n_levels != 3 && n_levels != 4) return false;
  return precision == HA_CONV_FP32_SIMT || precision == HA_CONV_F16X3 || precision == HA_CONV_F16 || precision == HA_CONV_F16X3_1CTA;
}

extern "C" size_t ha_vgg_workspace_bytes(int B, int H, int W, int n_levels, int precision) {
  if (!vgg_shape_ok(B, H, W, n_levels, precision)) return 0;
  ha::Arena ar{nullptr, 0, 0, true};
  ha::vgg_run(nullptr, nullptr, B, H, W, n_levels, precision, nullptr, nullptr, nullptr, ar, nullptr);
  return ha::align_up(ar.off, 256);
}

extern "C" int ha_vgg_forward(const void* packed_weights, const float* img_nchw, int B, int H, int W, int n_levels,
                              int precision, float* const* out_feat, float* const* out_scale, float* const* out_conf,
                              void* ws, size_t ws_bytes, void* stream) {
  if (!packed_weights || !img_nchw || !out_feat || !ws) return HA_EINVAL;
  if (!vgg_shape_ok(B, H, W, n_levels, precision)) return HA_EINVAL;
  for (int l = 0; l < n_levels; ++l)
    if (!out_feat[l]) return HA_EINVAL;
  ha::Arena ar{reinterpret_cast<char*>(ws), 0, ws_bytes, false};
  return ha::vgg_run(reinterpret_cast<const char*>(packed_weights), img_nchw, B, H, W, n_levels, precision, out_feat,
                     out_scale, out_conf, ar, reinterpret_cast<cudaStream_t>(stream));
}

// ---- train-mode forward: the same tcgen05 schedule, keeping what the backward pass needs (vgg_backward.cu)
extern "C" size_t ha_vgg_train_workspace_bytes(int B, int H, int W, int n_levels) {
  if (!vgg_shape_ok(B, H, W, n_levels, HA_CONV_F16X3)) return 0;
  ha::Arena ar{nullptr, 0, 0, true};
  ha::TcSaved sv;
  ha::vgg_run(nullptr, nullptr, B, H, W, n_levels, HA_CONV_F16X3, nullptr, nullptr, nullptr, ar, nullptr, &sv);
  return ha::align_up(ar.off, 256);
}

extern "C" int ha_vgg_forward_train(const void* packed_weights, const float* img_nchw, int B, int H, int W, int n_levels,
                                    float* const* out_feat, float* const* out_scale, float* const* out_conf, void* ws,
                                    size_t ws_bytes, void* stream) {
  if (!packed_weights || !img_nchw || !out_feat || !ws) return HA_EINVAL;
  if (!vgg_shape_ok(B, H, W, n_levels, HA_CONV_F16X3)) return HA_EINVAL;
  for (int l = 0; l < n_levels; ++l)
    if (!out_feat[l]) return HA_EINVAL;
  ha::Arena ar{reinterpret_cast<char*>(ws), 0, ws_bytes, false};
  ha::TcSaved sv;
  return ha::vgg_run(reinterpret_cast<const char*>(packed_weights), img_nchw, B, H, W, n_levels, HA_CONV_F16X3, out_feat,
                     out_scale, out_conf, ar, reinterpret_cast<cudaStream_t>(stream), &sv);
}

// ---- VGGUnet_G2S (VGG.py:206-345; the ground branch of LM_G2SP --proj nn): same weights and encoder, decoders on the folded maps
extern "C" int ha_vgg_g2s_forward(const void* packed_weights, const float* img_nchw, int B, int H, int W, int n_levels, int precision,
                                  float* const* out_feat, float* const* out_scale, float* const* out_conf, void* ws, size_t ws_bytes,
                                  void* stream) {
  if (!packed_weights || !img_nchw || !out_feat || !ws) return HA_EINVAL;
  if (!vgg_shape_ok(B, H, W, n_levels, precision) || precision == HA_CONV_FP32_SIMT || (W % 128) || (H % 32)) return HA_EINVAL;
  for (int l = 0; l < n_levels; ++l)
    if (!out_feat[l]) return HA_EINVAL;
  ha::Arena ar{reinterpret_cast<char*>(ws), 0, ws_bytes, false};
  return ha::vgg_run(reinterpret_cast<const char*>(packed_weights), img_nchw, B, H, W, n_levels, precision, out_feat, out_scale,
                     out_conf, ar, reinterpret_cast<cudaStream_t>(stream), nullptr, true);
}
