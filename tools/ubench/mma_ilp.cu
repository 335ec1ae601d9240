// Micro-benchmark: tcgen05.mma throughput vs number of independent accumulators (D tiles) and M.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../stcd_b200/csrc/ptx.cuh"
using namespace stcd;

__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t idesc_mn(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
template <int NACC>
__global__ void __launch_bounds__(128) k(int m_tile, int n_tile, int reps, int same_a, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tb;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_mn(m_tile, n_tile);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 128 * 1024);
    uint64_t bd = desc_nosw(b, n_tile * 16, 128);
    uint64_t ad[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) ad[j] = desc_nosw(a + (same_a ? 0 : j * 6144), 2880, 160);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
#pragma unroll
      for (int j = 0; j < NACC; ++j) umma_bf16(tb + j * (512 / NACC), ad[j], bd, idesc, 1);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int NACC>
void run(long long* d, int m, int n, int same_a) {
  const int reps = 256;
  cudaFuncSetAttribute(k<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  k<NACC><<<148, 128, 170 * 1024>>>(m, n, reps, same_a, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("M=%3d N=%3d nacc=%d same_a=%d  issue %.1f  total %.1f cyc/mma %s\n", m, n, NACC, same_a, (double)h[0] / (reps * NACC),
         (double)h[1] / (reps * NACC), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  for (int m : {128, 64})
    for (int n : {16, 32, 64, 128}) {
      for (int same : {0, 1}) {
        run<1>(d, m, n, same); run<2>(d, m, n, same); run<4>(d, m, n, same); if (n <= 64) run<8>(d, m, n, same);
      }
    }
  return 0;
}
