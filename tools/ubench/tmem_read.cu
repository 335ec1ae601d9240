// Micro-benchmark: TMEM read bandwidth of tcgen05.ld (32x32b.x16: 32 lanes x 16 columns x 4 B = 2 KB per instruction) with
// 1, 2 and 4 warps of one CTA reading their own lane quarters, and with two CTAs per SM.  The horizontally folded convs read 3x
// the accumulator columns: whether that is 24*cs or 6*cs cycles per tile decides where the folding pays.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../stcd_b200/csrc/ptx.cuh"
using namespace stcd;

__global__ void __launch_bounds__(128) k(int warps, int reps, int cols, long long* out, float* sink) {
  __shared__ uint32_t tb;
  if (threadIdx.x < 32) { tmem_alloc(&tb, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5;
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    const uint32_t base = tb + (static_cast<uint32_t>(warp * 32) << 16);
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      for (int c = 0; c < cols; c += 64) {        // four loads in flight, then one wait
        uint32_t a[16], b[16], c2[16], d[16];
        tmem_ld16(base + c, a);
        tmem_ld16(base + c + 16, b);
        tmem_ld16(base + c + 32, c2);
        tmem_ld16(base + c + 48, d);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += __uint_as_float(a[j] ^ b[j] ^ c2[j] ^ d[j]);
      }
    }
    t1 = clock64();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  float* sink; cudaMalloc(&sink, 4096);
  const int reps = 200, cols = 256;
  for (int ctas : {148, 296})
    for (int warps : {1, 2, 4}) {
      k<<<ctas, 128>>>(warps, reps, cols, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double loads = (double)reps * cols / 16;
      printf("ctas=%d (%d per SM) warps=%d: %.1f cycles per LDTM.x16 per warp -> %.0f B/clk per CTA, %.0f B/clk per SM  %s\n", ctas, ctas / 148, warps,
             h / loads, warps * 2048.0 * loads / h, (ctas / 148) * warps * 2048.0 * loads / h, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
