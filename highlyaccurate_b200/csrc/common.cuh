// Shared helpers for the sm_100a kernels of libha_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ha_b200.h"

namespace ha {

void set_cuda_error(cudaError_t e, const char* what);
void count_launches(int n);   // statistics only: kernels launched by this library (ha_launch_count)
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e, what); return HA_ECUDA; }
  return HA_OK;
}
#define HA_CUDA_TRY(expr)                                             \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) { ha::set_cuda_error(_e, #expr); return HA_ECUDA; } \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float4 ldg_nc_stream(const float4* p) {
  // streamed-once data (ground features): read-only path, do not allocate in L1
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_nc(const float4* p) { return __ldg(p); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ha
