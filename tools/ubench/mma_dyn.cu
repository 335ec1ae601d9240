// Micro-benchmark: what does the conv kernel's issue loop pay per tcgen05.mma beyond the MMA itself?
// Same unrolled 9-tap / 3-tap groups as mma_n.cu (45-64 cycles per MMA with loop-invariant descriptors), plus, per group ("chunk"):
//   dyn = 1     the A descriptors get a stage offset that changes every chunk (ring of 4 stages): fresh R2UR per MMA
//   commit = 1  one tcgen05.commit per chunk onto an mbarrier ring (what releases the A stage in the kernel)
//   wait = 1    one mbarrier try_wait per chunk (on a barrier that has already completed: the a_full wait when data is early)
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
__global__ void __launch_bounds__(128) k(int n_tile, int reps, int dyn, int commit, int wait, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar, ring[4], done_bar;
  __shared__ uint32_t tb;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&done_bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&ring[i], 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(&tb, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) mbar_arrive(&done_bar);     // phase 0 of done_bar is complete: waits on it return at once
  __syncthreads();
  if (threadIdx.x < 32) {                           // the whole warp runs the loop, one elected lane issues (as in the kernel)
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(n_tile);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 96 * 1024);
    uint32_t alo[TAPS], blo[TAPS];
    const uint32_t a_hi = (uint32_t)(desc_nosw(0, 2880, 160) >> 32), b_hi = (uint32_t)(desc_nosw(0, n_tile * 16, 128) >> 32);
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
      alo[t] = (uint32_t)desc_nosw(a + ((t / 3) * 10 + (t % 3)) * 16, 2880, 160);
      blo[t] = (uint32_t)desc_nosw(b + t * n_tile * 32, n_tile * 16, 128);
    }
    long long t0 = clock64();
    int s = 0;
    for (int i = 0; i < reps; ++i) {
      if (wait) { mbar_wait(&done_bar, 0); tc_fence_after(); }
      const uint32_t st = dyn ? (uint32_t)s * (23040u >> 4) : 0u;
      if (leader) {
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
#pragma unroll
          for (int j = 0; j < NACC; ++j) umma_bf16_lohi(tb + j * (512 / NACC), alo[t] + st + j * 40, a_hi, blo[t], b_hi, idesc, 1);
      }
      if (commit && leader) umma_commit(&ring[s]);
      if (++s == 4) s = 0;
    }
    long long t1 = clock64();
    if (leader) umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int NACC, int TAPS>
void run(long long* d, int n) {
  const int reps = 64;
  cudaFuncSetAttribute(k<NACC, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int mode = 0; mode < 8; ++mode) {
    const int dyn = mode & 1, commit = (mode >> 1) & 1, wait = (mode >> 2) & 1;
    k<NACC, TAPS><<<148, 128, 200 * 1024>>>(n, reps, dyn, commit, wait, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d nacc=%d taps=%d dyn=%d commit=%d wait=%d  issue %.1f  total %.1f cyc/mma  %s\n", n, NACC, TAPS, dyn, commit, wait,
           (double)h[0] / (reps * NACC * TAPS), (double)h[1] / (reps * NACC * TAPS), e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  for (int n : {16, 64, 96, 128}) {
    run<1, 9>(d, n);
    run<2, 9>(d, n);
    run<2, 3>(d, n);
  }
  return 0;
}
