// Library-level entry points of libha_b200.so: version, errors, device check, layout helpers.
#include <string.h>

#include "common.cuh"

namespace ha {

static thread_local char g_cuda_err[512] = "";
static unsigned long long g_launches = 0;   // statistics only

void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

void set_cuda_error(cudaError_t e, const char* what) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

void set_error_text(const char* text) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s", text); }

// [B][C][HW] <-> [B][HW][C] through a 32x33 shared tile: coalesced on both sides.
template <bool TO_NHWC>
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int rows = TO_NHWC ? C : HW, cols = TO_NHWC ? HW : C;    // src is [rows][cols]
  const float* s = src + (size_t)b * C * HW;
  float* d = dst + (size_t)b * C * HW;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = s[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) d[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

template <bool TO_NHWC>
static int transpose(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
  if (!src || !dst || B <= 0 || C <= 0 || H <= 0 || W <= 0) return HA_EINVAL;
  const int HW = H * W;
  const int rows = TO_NHWC ? C : HW, cols = TO_NHWC ? HW : C;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, B), block(32, 8);
  if (grid.y > 65535 || grid.z > 65535) return HA_EINVAL;
  transpose_kernel<TO_NHWC><<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, C, HW);
  count_launches(1);
  return check_launch("transpose_kernel");
}

}  // namespace ha

extern "C" int ha_version(void) { return 3; }
extern "C" unsigned long long ha_launch_count(void) { return __atomic_load_n(&ha::g_launches, __ATOMIC_RELAXED); }

extern "C" const char* ha_error_string(int code) {
  switch (code) {
    case HA_OK: return "ok";
    case HA_EINVAL: return "invalid argument (shape, alignment or unsupported channel count)";
    case HA_ENOSPACE: return "workspace too small";
    case HA_ECUDA: return "CUDA call failed (see ha_last_cuda_error)";
    case HA_EUNSUPPORTED: return "device is not compute capability 10.x (B200, sm_100a)";
    case HA_ECOMM: return "NCCL unavailable or an NCCL call failed (see ha_last_cuda_error)";
    default: return "unknown error";
  }
}

extern "C" const char* ha_last_cuda_error(void) { return ha::g_cuda_err; }

extern "C" int ha_device_check(int device) {
  cudaDeviceProp prop;
  HA_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  return prop.major == 10 ? HA_OK : HA_EUNSUPPORTED;
}

extern "C" int ha_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
  return ha::transpose<true>(src, dst, B, C, H, W, stream);
}
extern "C" int ha_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
  return ha::transpose<false>(src, dst, B, C, H, W, stream);
}
