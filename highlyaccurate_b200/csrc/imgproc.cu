// Input pipeline on the GPU: the per-sample image preparation the reference does on the CPU with
// PIL + torchvision (dataLoader/KITTI_dataset.py:128-157, :256-288; dataLoader/Ford_dataset.py:178-209):
//   Image.rotate (nearest)  ->  Image.transform(AFFINE, BILINEAR)  ->  ...  ->  TF.center_crop  ->  Resize  ->  ToTensor.
// Every PIL stage rounds to uint8, so the stages cannot be merged algebraically; each one is one HBM-bound launch over the
// batch (4 bytes per pixel in, 4 out, coalesced 32-bit accesses), and the arithmetic is Pillow's, bit for bit
// (libImaging Geometry.c affine_fixed / bilinear_filter32RGB, Resample.c 8bpc passes; restated and pinned against PIL in
// oracle/imgproc.py + tests/test_imgproc.py):
//   * nearest: 16.16 fixed point, coefficients FIX(v) = floor(v * 65536 + 0.5), source pixel = accumulator >> 16;
//   * bilinear: float64 without contraction (the x86-64 wheels have no FMA), clamped neighbours, truncation to uint8;
//   * resize: normalised float64 filter weights -> 22-bit fixed point, horizontal then vertical pass, uint8 in between;
//   * ToTensor: uint8 -> fp32 / 255 (correctly rounded division), HWC -> CHW; the centre crop is fused into whichever
//     kernel writes the tensor.
#include "common.cuh"

namespace ha {

constexpr int kImgTX = 32, kImgTY = 8;

__device__ __forceinline__ int pil_floor(double v) { return v >= 0.0 ? (int)v : (int)floor(v); }
__device__ __forceinline__ int pil_fix(double v) { return pil_floor(__dadd_rn(__dmul_rn(v, 65536.0), 0.5)); }

template <int BPP>
__device__ __forceinline__ uchar4 load_px(const uint8_t* __restrict__ img, size_t idx) {
  if (BPP == 4) return reinterpret_cast<const uchar4*>(img)[idx];
  const uint8_t* p = img + idx * 3;
  return make_uchar4(p[0], p[1], p[2], 0);
}

__device__ __forceinline__ void store_tensor(float* __restrict__ dst, int b, int side_h, int side_w, int oy, int ox, uchar4 v) {
  const size_t plane = (size_t)side_h * side_w;
  float* d = dst + (size_t)b * 3 * plane + (size_t)oy * side_w + ox;
  d[0] = __fdiv_rn((float)v.x, 255.f);            // ToTensor: .div(255) in fp32
  d[plane] = __fdiv_rn((float)v.y, 255.f);
  d[2 * plane] = __fdiv_rn((float)v.z, 255.f);
}

// One output pixel per thread.  dst_u8: the whole H x W image as RGBX; dst_f32: the centre crop as a [3][side][side] tensor
// (the grid then covers the crop window only).
template <bool BILIN, int BPP>
__global__ void __launch_bounds__(kImgTX * kImgTY)
img_affine_kernel(const uint8_t* __restrict__ src, uchar4* __restrict__ dst_u8, float* __restrict__ dst_f32,
                  const double* __restrict__ coef, int H, int W, int crop_top, int crop_left, int side) {
  const int b = blockIdx.z;
  const int ox = blockIdx.x * kImgTX + threadIdx.x, oy = blockIdx.y * kImgTY + threadIdx.y;
  const int x = ox + crop_left, y = oy + crop_top;
  if (dst_f32 ? (ox >= side || oy >= side) : (x >= W || y >= H)) return;
  const double* m = coef + (size_t)b * 6;
  const uint8_t* img = src + (size_t)b * H * W * BPP;
  uchar4 out = make_uchar4(0, 0, 0, 0);
  if (!BILIN) {
    // affine_fixed(): the C loop adds a0 / a3 per column and a1 / a4 per row to 32-bit accumulators
    const int a0 = pil_fix(m[0]), a1 = pil_fix(m[1]), a3 = pil_fix(m[3]), a4 = pil_fix(m[4]);
    const int a2 = pil_fix(__dadd_rn(__dadd_rn(m[2], __dmul_rn(m[0], 0.5)), __dmul_rn(m[1], 0.5)));
    const int a5 = pil_fix(__dadd_rn(__dadd_rn(m[5], __dmul_rn(m[3], 0.5)), __dmul_rn(m[4], 0.5)));
    const int xx = (int)((unsigned)a2 + (unsigned)y * (unsigned)a1 + (unsigned)x * (unsigned)a0);
    const int yy = (int)((unsigned)a5 + (unsigned)y * (unsigned)a4 + (unsigned)x * (unsigned)a3);
    const int xin = xx >> 16, yin = yy >> 16;
    if (xin >= 0 && xin < W && yin >= 0 && yin < H) out = load_px<BPP>(img, (size_t)yin * W + xin);
  } else {
    // affine_transform() + bilinear_filter32RGB()
    const double xc = (double)x + 0.5, yc = (double)y + 0.5;
    double xin = __dadd_rn(__dadd_rn(__dmul_rn(m[0], xc), __dmul_rn(m[1], yc)), m[2]);
    double yin = __dadd_rn(__dadd_rn(__dmul_rn(m[3], xc), __dmul_rn(m[4], yc)), m[5]);
    if (!(xin < 0.0 || xin >= (double)W || yin < 0.0 || yin >= (double)H)) {
      xin = __dsub_rn(xin, 0.5);
      yin = __dsub_rn(yin, 0.5);
      const int fx = pil_floor(xin), fy = pil_floor(yin);
      const double dx = __dsub_rn(xin, (double)fx), dy = __dsub_rn(yin, (double)fy);
      const int x0 = fx < 0 ? 0 : (fx < W ? fx : W - 1), x1 = fx + 1 < 0 ? 0 : (fx + 1 < W ? fx + 1 : W - 1);
      const int y0 = fy < 0 ? 0 : (fy < H ? fy : H - 1);
      const uchar4 p00 = load_px<BPP>(img, (size_t)y0 * W + x0), p01 = load_px<BPP>(img, (size_t)y0 * W + x1);
      const bool has2 = fy + 1 >= 0 && fy + 1 < H;
      uchar4 p10 = p00, p11 = p01;
      if (has2) { p10 = load_px<BPP>(img, (size_t)(fy + 1) * W + x0); p11 = load_px<BPP>(img, (size_t)(fy + 1) * W + x1); }
      auto lerp2 = [&](int a, int bb, int c, int d) -> uint8_t {
        const double v1 = __dadd_rn((double)a, __dmul_rn((double)(bb - a), dx));
        const double v2 = has2 ? __dadd_rn((double)c, __dmul_rn((double)(d - c), dx)) : v1;
        return (uint8_t)(int)__dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), dy));
      };
      out.x = lerp2(p00.x, p01.x, p10.x, p11.x);
      out.y = lerp2(p00.y, p01.y, p10.y, p11.y);
      out.z = lerp2(p00.z, p01.z, p10.z, p11.z);
    }
  }
  if (dst_f32) store_tensor(dst_f32, b, side, side, oy, ox, out);
  else dst_u8[((size_t)b * H + y) * W + x] = out;
}

// precompute_coeffs() + normalize_coeffs_8bpc() of Resample.c for the bilinear filter (support 1), one output index per thread
__global__ void img_resample_coef_kernel(int in_size, int out_size, int ksize, int* __restrict__ kk, int* __restrict__ bounds) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  if (xx >= out_size) return;
  const double scale = __ddiv_rn((double)in_size, (double)out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                      // filterp->support (1.0) * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double w[16];
  double ww = 0.0;
  for (int x = 0; x < xmax; ++x) {
    double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (t < 0.0) t = -t;
    w[x] = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
    ww = __dadd_rn(ww, w[x]);
  }
  int* k = kk + (size_t)xx * ksize;
  for (int x = 0; x < ksize; ++x) {
    double v = x < xmax ? w[x] : 0.0;
    if (x < xmax && ww != 0.0) v = __ddiv_rn(v, ww);
    v = __dmul_rn(v, 4194304.0);                            // 1 << PRECISION_BITS, PRECISION_BITS = 32 - 8 - 2
    k[x] = v < 0.0 ? (int)__dadd_rn(-0.5, v) : (int)__dadd_rn(0.5, v);
  }
  bounds[xx * 2] = xmin;
  bounds[xx * 2 + 1] = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= 22;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ImagingResampleHorizontal_8bpc: [B][H][W] -> [B][H][out_w] RGBX
template <int BPP>
__global__ void __launch_bounds__(kImgTX * kImgTY)
img_resample_h_kernel(const uint8_t* __restrict__ src, uchar4* __restrict__ dst, const int* __restrict__ kk,
                      const int* __restrict__ bounds, int ksize, int H, int W, int out_w) {
  const int b = blockIdx.z;
  const int xx = blockIdx.x * kImgTX + threadIdx.x, y = blockIdx.y * kImgTY + threadIdx.y;
  if (xx >= out_w || y >= H) return;
  const int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
  const int* k = kk + (size_t)xx * ksize;
  const uint8_t* img = src + (size_t)b * H * W * BPP;
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
  for (int x = 0; x < xmax; ++x) {
    const uchar4 p = load_px<BPP>(img, (size_t)y * W + xmin + x);
    const int c = k[x];
    s0 += p.x * c; s1 += p.y * c; s2 += p.z * c;
  }
  dst[((size_t)b * H + y) * out_w + xx] = make_uchar4(clip8(s0), clip8(s1), clip8(s2), 0);
}

// ImagingResampleVertical_8bpc, writing the ToTensor result: [B][H][W] -> [B][3][out_h][W] fp32
template <int BPP>
__global__ void __launch_bounds__(kImgTX * kImgTY)
img_resample_v_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, const int* __restrict__ kk,
                      const int* __restrict__ bounds, int ksize, int H, int W, int out_h) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * kImgTX + threadIdx.x, yy = blockIdx.y * kImgTY + threadIdx.y;
  if (x >= W || yy >= out_h) return;
  const int ymin = bounds[yy * 2], ymax = bounds[yy * 2 + 1];
  const int* k = kk + (size_t)yy * ksize;
  const uint8_t* img = src + (size_t)b * H * W * BPP;
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
  for (int y = 0; y < ymax; ++y) {
    const uchar4 p = load_px<BPP>(img, (size_t)(ymin + y) * W + x);
    const int c = k[y];
    s0 += p.x * c; s1 += p.y * c; s2 += p.z * c;
  }
  store_tensor(dst, b, out_h, W, yy, x, make_uchar4(clip8(s0), clip8(s1), clip8(s2), 0));
}

// ToTensor alone (both sizes already match: Image.resize returns a copy)
template <int BPP>
__global__ void __launch_bounds__(kImgTX * kImgTY)
img_to_tensor_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int H, int W) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * kImgTX + threadIdx.x, y = blockIdx.y * kImgTY + threadIdx.y;
  if (x >= W || y >= H) return;
  store_tensor(dst, b, H, W, y, x, load_px<BPP>(src + (size_t)b * H * W * BPP, (size_t)y * W + x));
}

static int resample_ksize(int in_size, int out_size) {
  const double scale = (double)in_size / out_size;
  const double support = scale < 1.0 ? 1.0 : scale;
  return (int)ceil(support) * 2 + 1;
}

struct ResizeWs { int* kk_h; int* bd_h; int* kk_v; int* bd_v; uint8_t* tmp; size_t total; int ks_h, ks_v; };
static ResizeWs resize_ws_carve(void* ws, int B, int H, int W, int out_h, int out_w) {
  ResizeWs r;
  char* p = reinterpret_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t bytes) { char* q = p + off; off = (off + bytes + 255) / 256 * 256; return q; };
  r.ks_h = resample_ksize(W, out_w);
  r.ks_v = resample_ksize(H, out_h);
  r.kk_h = reinterpret_cast<int*>(take((size_t)out_w * r.ks_h * 4));
  r.bd_h = reinterpret_cast<int*>(take((size_t)out_w * 8));
  r.kk_v = reinterpret_cast<int*>(take((size_t)out_h * r.ks_v * 4));
  r.bd_v = reinterpret_cast<int*>(take((size_t)out_h * 8));
  r.tmp = reinterpret_cast<uint8_t*>(take((size_t)B * H * out_w * 4));
  r.total = off;
  return r;
}

}  // namespace ha

// ------------------------------------------------------------------------------------ C ABI
extern "C" int ha_img_affine_u8(const uint8_t* src, int src_pixel_bytes, uint8_t* dst_rgbx, float* dst_chw, int crop_side, int B,
                                int H, int W, const double* coef, int resample, void* stream) {
  using namespace ha;
  if (!src || !coef || (!dst_rgbx == !dst_chw) || B <= 0 || H <= 0 || W <= 0 || B > 65535) return HA_EINVAL;
  if ((src_pixel_bytes != 3 && src_pixel_bytes != 4) || (resample != 0 && resample != 2)) return HA_EINVAL;
  if (H >= 16384 || W >= 16384) return HA_EINVAL;          // keeps Geometry.c's check_fixed() true for every sane transform
  if (src_pixel_bytes == 4 && ((uintptr_t)src & 3)) return HA_EINVAL;
  int top = 0, left = 0, side = 0;
  if (dst_chw) {
    if (crop_side <= 0 || crop_side > H || crop_side > W) return HA_EINVAL;
    side = crop_side;
    top = (int)nearbyint((H - side) / 2.0);                 // torchvision center_crop: int(round(x / 2.0)), half to even
    left = (int)nearbyint((W - side) / 2.0);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 block(kImgTX, kImgTY);
  const dim3 grid(((dst_chw ? side : W) + kImgTX - 1) / kImgTX, ((dst_chw ? side : H) + kImgTY - 1) / kImgTY, B);
  uchar4* d8 = reinterpret_cast<uchar4*>(dst_rgbx);
#define HA_IMG_LAUNCH(BIL, BPP) img_affine_kernel<BIL, BPP><<<grid, block, 0, st>>>(src, d8, dst_chw, coef, H, W, top, left, side)
  if (resample == 2) { if (src_pixel_bytes == 3) HA_IMG_LAUNCH(true, 3); else HA_IMG_LAUNCH(true, 4); }
  else { if (src_pixel_bytes == 3) HA_IMG_LAUNCH(false, 3); else HA_IMG_LAUNCH(false, 4); }
#undef HA_IMG_LAUNCH
  count_launches(1);
  return check_launch("img_affine_kernel");
}

extern "C" size_t ha_img_resize_workspace_bytes(int B, int H, int W, int out_h, int out_w) {
  if (B <= 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0) return 0;
  return ha::resize_ws_carve(nullptr, B, H, W, out_h, out_w).total;
}

extern "C" int ha_img_resize_to_tensor(const uint8_t* src, int src_pixel_bytes, int B, int H, int W, int out_h, int out_w,
                                       float* dst_chw, void* ws, size_t ws_bytes, void* stream) {
  using namespace ha;
  if (!src || !dst_chw || !ws || B <= 0 || B > 65535 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0) return HA_EINVAL;
  if (src_pixel_bytes != 3 && src_pixel_bytes != 4) return HA_EINVAL;
  const ResizeWs r = resize_ws_carve(ws, B, H, W, out_h, out_w);
  if (ws_bytes < r.total) return HA_ENOSPACE;
  if (r.ks_h > 16 || r.ks_v > 16) return HA_EINVAL;        // down-scaling by more than 7.5x is outside the datasets' range
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 block(kImgTX, kImgTY);
  const uint8_t* cur = src;
  int cur_bpp = src_pixel_bytes, cur_w = W;
  if (W != out_w) {
    img_resample_coef_kernel<<<(out_w + 127) / 128, 128, 0, st>>>(W, out_w, r.ks_h, r.kk_h, r.bd_h);
    const dim3 grid((out_w + kImgTX - 1) / kImgTX, (H + kImgTY - 1) / kImgTY, B);
    if (cur_bpp == 3) img_resample_h_kernel<3><<<grid, block, 0, st>>>(cur, reinterpret_cast<uchar4*>(r.tmp), r.kk_h, r.bd_h, r.ks_h, H, W, out_w);
    else img_resample_h_kernel<4><<<grid, block, 0, st>>>(cur, reinterpret_cast<uchar4*>(r.tmp), r.kk_h, r.bd_h, r.ks_h, H, W, out_w);
    count_launches(2);
    cur = r.tmp; cur_bpp = 4; cur_w = out_w;
  }
  const dim3 grid((cur_w + kImgTX - 1) / kImgTX, (out_h + kImgTY - 1) / kImgTY, B);
  if (H != out_h) {
    img_resample_coef_kernel<<<(out_h + 127) / 128, 128, 0, st>>>(H, out_h, r.ks_v, r.kk_v, r.bd_v);
    if (cur_bpp == 3) img_resample_v_kernel<3><<<grid, block, 0, st>>>(cur, dst_chw, r.kk_v, r.bd_v, r.ks_v, H, cur_w, out_h);
    else img_resample_v_kernel<4><<<grid, block, 0, st>>>(cur, dst_chw, r.kk_v, r.bd_v, r.ks_v, H, cur_w, out_h);
    count_launches(2);
  } else {
    if (cur_bpp == 3) img_to_tensor_kernel<3><<<grid, block, 0, st>>>(cur, dst_chw, H, cur_w);
    else img_to_tensor_kernel<4><<<grid, block, 0, st>>>(cur, dst_chw, H, cur_w);
    count_launches(1);
  }
  return check_launch("img_resample kernels");
}
