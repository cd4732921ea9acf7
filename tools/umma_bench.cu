// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (M = 128 per CTA, K = 16) against N, from static shared-memory
// operands (no TMA in the loop).  Answers: is there a per-instruction floor that makes small-N MMAs inefficient?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_bench tools/umma_bench.cu && build/umma_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: N1 only; mode 1: alternate (A0 x B, N1) and (A1 x B, N2) like the f16x3 kernels
template <int N1, int N2, int MODE>
__global__ void __launch_bounds__(128, 1) bench(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (2 * 16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint64_t a0 = desc_sw128(smem_u32(smem)), a1 = desc_sw128(smem_u32(smem) + 16384), b = desc_sw128(smem_u32(smem) + 32768);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma(tmem, a0 + 2 * k, b + 2 * k, idesc_f16(N1), 1);
        if (MODE == 1) umma(tmem + 256, a1 + 2 * k, b + 2 * k, idesc_f16(N2), 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N1, int N2, int MODE>
void run(const char* what, int grid) {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 2 * 16384 + 32768 + 1024, iters = 2000;
  cudaFuncSetAttribute(bench<N1, N2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  bench<N1, N2, MODE><<<grid, 128, smem>>>(50, d);
  bench<N1, N2, MODE><<<grid, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const int n_mma = iters * 4 * (MODE == 1 ? 2 : 1);
  const double math = MODE == 1 ? (128.0 * N1 / 256 + 128.0 * N2 / 256) / 2 : 128.0 * N1 / 256;
  printf("%-34s grid %3d: %7.1f cycles per MMA (math floor %5.1f)  %s\n", what, grid, (double)h / n_mma, math, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, 0, 0>("N=64", grid);
    run<128, 0, 0>("N=128", grid);
    run<256, 0, 0>("N=256", grid);
    run<128, 64, 1>("N=128 then N=64 (Cout-64 f16x3)", grid);
    run<256, 128, 1>("N=256 then N=128 (Cout-128 f16x3)", grid);
    run<256, 256, 1>("N=256 then N=256", grid);
  }
  return 0;
}
