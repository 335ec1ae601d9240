// Micro-benchmark: does mbarrier try_wait spinning by other warps slow down tcgen05.mma issue/execution?
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
// spin_mode: 0 none, 1 all lanes of 5 warps spin on try_wait, 2 one lane per warp spins, 3 all lanes with nanosleep(64)
__global__ void __launch_bounds__(224) k(int n_tile, int spin_mode, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, never;
  __shared__ uint32_t tb;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 2) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(n_tile);
      const uint32_t a = smem_u32(smem), b = smem_u32(smem + 48 * 1024);
      uint64_t ad = desc_nosw(a, 2880, 160), bd = desc_nosw(b, n_tile * 16, 128);
      long long t0 = clock64();
      for (int i = 0; i < reps; ++i) umma_bf16(tb, ad + (i % 9), bd, idesc, 1);
      long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      mbar_arrive(&never);   // release spinners
    }
    __syncwarp();
  } else if (spin_mode != 0) {
    if (spin_mode == 1 || (spin_mode == 2 && lane == 0)) {
      while (!mbar_try_wait(&never, 0)) {}
    } else if (spin_mode == 3) {
      while (!mbar_try_wait(&never, 0)) __nanosleep(64);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 256;
  for (int mode = 0; mode < 4; ++mode)
    for (int n : {16, 128}) {
      k<<<148, 224, 100 * 1024>>>(n, mode, reps, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("spin_mode=%d N=%3d  issue %.1f cyc/mma  total %.1f cyc/mma %s\n", mode, n, (double)h[0] / reps, (double)h[1] / reps,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
