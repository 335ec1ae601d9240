// Micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16, SS mode) by N, for the two operand layouts of the conv kernel:
//   layout 0: 16 x 8 tile with an 18 x 10 halo box (SBO = 160 B, LBO = 2880 B), 9 tap offsets
//   layout 1: horizontally folded 8 x 16 tile, 10 x 16 box (SBO = 128 B, LBO = 2560 B), 3 tap offsets (256 B apart)
// The issue loop is fully unrolled with precomputed descriptors (a single thread retires ~1 dependent instruction per 10 cycles,
// so anything computed inside the loop would be what is measured).  NACC independent accumulators.  One CTA per SM.
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
template <int NACC, int TAPS>
__global__ void __launch_bounds__(128) k(int n_tile, int layout, int reps, long long* out) {
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
    const uint32_t idesc = make_idesc_bf16(n_tile);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    uint64_t ad[TAPS], bd[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
      const uint32_t aoff = layout == 0 ? ((t / 3) * 10 + (t % 3)) * 16 : t * 256;
      ad[t] = layout == 0 ? desc_nosw(a + aoff, 2880, 160) : desc_nosw(a + aoff, 2560, 128);
      bd[t] = desc_nosw(b + t * n_tile * 32, n_tile * 16, 128);
    }
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
#pragma unroll
      for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int j = 0; j < NACC; ++j) umma_bf16(tb + j * (512 / NACC), ad[t] + j * 640, bd[t], idesc, 1);
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

template <int NACC, int TAPS>
void run(long long* d, int n, int layout) {
  const int reps = 64;
  cudaFuncSetAttribute(k<NACC, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  k<NACC, TAPS><<<148, 128, 170 * 1024>>>(n, layout, reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double per = (double)h[1] / (reps * NACC * TAPS);
  printf("layout=%d N=%3d nacc=%d taps=%d  issue %.1f  total %.1f cyc/mma  -> %.0f MAC/clk (%.0f%% of 4096)  %s\n", layout, n, NACC, TAPS,
         (double)h[0] / (reps * NACC * TAPS), per, 128.0 * n * 16 / per, 100.0 * 128.0 * n * 16 / per / 4096.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  for (int n : {16, 32, 48, 64, 96, 128, 192, 256}) {
    run<1, 9>(d, n, 0);
    run<2, 9>(d, n, 0);
    run<1, 3>(d, n, 1);
    if (n <= 128) run<2, 3>(d, n, 1);
  }
  return 0;
}
