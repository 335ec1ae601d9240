// Micro-benchmark: tcgen05.mma issue/execute cost per instruction for the operand layouts the conv
// kernel uses (un-swizzled K-major with halo strides vs 128B-swizzled), by N.  One CTA per SM.
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

// mode 0: nosw halo (SBO=160, LBO=2880)  1: nosw dense (SBO=128, LBO=2048)  2: sw128 (row 128B, SBO 1024)
__global__ void __launch_bounds__(128) k(int n_tile, int mode, int reps, int fence_each, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tb;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(n_tile);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 48 * 1024);
    uint64_t ad, bd;
    if (mode == 0) { ad = desc_nosw(a, 2880, 160); bd = desc_nosw(b, n_tile * 16, 128); }
    else if (mode == 1) { ad = desc_nosw(a, 2048, 128); bd = desc_nosw(b, n_tile * 16, 128); }
    else if (mode == 3) { ad = desc_nosw(a, 3072, 128); bd = desc_nosw(b, n_tile * 16, 128); }
    else { ad = make_kmajor_desc(a, 128, 1024); bd = make_kmajor_desc(b, 128, 1024); }
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      if (fence_each) tc_fence_after();
      umma_bf16(tb, ad + (mode == 0 ? (i % 9) : (mode == 3 ? (i % 3) * 32 : 0)), bd, idesc, 1);
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

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 256;
  for (int fence = 0; fence < 1; ++fence)
    for (int mode : {0, 1, 3, 2})
      for (int n : {16, 32, 48, 64, 96, 128, 192, 256}) {
        for (int ctas : {1, 148}) {
          k<<<ctas, 128, 100 * 1024>>>(n, mode, reps, fence, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("fence=%d mode=%d N=%3d ctas=%3d  issue %.1f cyc/mma  total %.1f cyc/mma  (floor %.0f) %s\n", fence, mode, n, ctas,
                 (double)h[0] / reps, (double)h[1] / reps, 128.0 * n / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
      }
  return 0;
}
