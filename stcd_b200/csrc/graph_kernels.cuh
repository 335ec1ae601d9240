// ViG Grapher graph ops (gcn_lib, imported by models/pyramid_vig.py:17; semantics: SURVEY.md App. D):
// dense dilated kNN graph construction and the max-relative aggregation, on the reference's own layout
// (fp32 node features [B][C][N], N = H*W nodes; int64 neighbour tables).
//
// The neighbour ORDER is what the downstream max-relative features depend on, so the distance is computed in
// fp32 on the CUDA cores with the reference's expression (|x|^2 - 2 x.y + |y|^2 on L2-normalised nodes, plus the
// relative-position bias) rather than in bf16 on the tensor pipe: a bf16 distance reorders near-equidistant
// neighbours.  The work is small (N x M x C MACs with M <= 256) and the kernel keeps the whole [64 x M] distance
// tile in registers -- the dense [B, N, M] tensor of the reference never exists.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace stcd {

// denom[b][n] = max(||x[b, :, n]||_2, 1e-12)   (F.normalize(x, p=2, dim=1)).  One thread per node, coalesced over n.
__global__ void __launch_bounds__(256) node_norm_kernel(const float* __restrict__ x, float* __restrict__ denom, int B, int C,
                                                        int N) {
  const size_t total = static_cast<size_t>(B) * N;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / N, n = i - b * N;
    const float* p = x + b * C * N + n;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(p + static_cast<size_t>(c) * N);
      s = fmaf(v, v, s);
    }
    denom[i] = fmaxf(sqrtf(s), 1e-12f);
  }
}

constexpr int kKnnQ = 64;     // queries per CTA (8 per warp)
constexpr int kKnnM = 256;    // max keys (8 per lane)
constexpr int kKnnCC = 32;    // channels per shared-memory chunk

// grid (ceil(N / 64), B), 256 threads.  nn_idx[b][n][t] = index of the (t*dilation)-th nearest key of query n.
__global__ void __launch_bounds__(256) knn_graph_kernel(const float* __restrict__ x, const float* __restrict__ xden,
                                                        const float* __restrict__ y, const float* __restrict__ yden,
                                                        const float* __restrict__ relpos, int C, int N, int M, int k,
                                                        int dilation, long long* __restrict__ nn_idx) {
  __shared__ float xs[kKnnQ][kKnnCC + 1];
  __shared__ float ys[kKnnM][kKnnCC + 1];
  const int b = blockIdx.y, n0 = blockIdx.x * kKnnQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + static_cast<size_t>(b) * C * N;
  const float* yb = y + static_cast<size_t>(b) * C * M;
  float dot[8][8], xsq[8], ysq[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    xsq[q] = 0.f;
    ysq[q] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) dot[q][j] = 0.f;
  }
  for (int c0 = 0; c0 < C; c0 += kKnnCC) {
    __syncthreads();
    for (int i = threadIdx.x; i < kKnnQ * kKnnCC; i += blockDim.x) {
      const int q = i % kKnnQ, cc = i / kKnnQ;
      const int n = n0 + q, c = c0 + cc;
      xs[q][cc] = (n < N && c < C) ? __fdiv_rn(__ldg(xb + static_cast<size_t>(c) * N + n), __ldg(xden + static_cast<size_t>(b) * N + n)) : 0.f;
    }
    for (int i = threadIdx.x; i < kKnnM * kKnnCC; i += blockDim.x) {
      const int j = i % kKnnM, cc = i / kKnnM;
      const int c = c0 + cc;
      ys[j][cc] = (j < M && c < C) ? __fdiv_rn(__ldg(yb + static_cast<size_t>(c) * M + j), __ldg(yden + static_cast<size_t>(b) * M + j)) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int cc = 0; cc < kKnnCC; ++cc) {
      float xv[8], yv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) xv[q] = xs[warp * 8 + q][cc];
#pragma unroll
      for (int j = 0; j < 8; ++j) yv[j] = ys[lane + 32 * j][cc];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        xsq[q] = fmaf(xv[q], xv[q], xsq[q]);
#pragma unroll
        for (int j = 0; j < 8; ++j) dot[q][j] = fmaf(xv[q], yv[j], dot[q][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ysq[j] = fmaf(yv[j], yv[j], ysq[j]);
    }
  }
  const int kd = k * dilation;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int n = n0 + warp * 8 + q;
    if (n >= N) continue;      // warp-uniform
    float dist[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = lane + 32 * j;
      // the reference's order of operations: (x_sq + (-2 * inner)) + y_sq, then + relative_pos
      float dv = __fadd_rn(__fadd_rn(xsq[q], -2.f * dot[q][j]), ysq[j]);
      if (relpos != nullptr && key < M) dv = __fadd_rn(dv, __ldg(relpos + static_cast<size_t>(n) * M + key));
      dist[j] = key < M ? dv : CUDART_INF_F;
    }
    long long* out = nn_idx + (static_cast<size_t>(b) * N + n) * k;
    for (int t = 0; t < kd; ++t) {
      float best = dist[0];
      int bi = lane;
#pragma unroll
      for (int j = 1; j < 8; ++j)
        if (dist[j] < best) {
          best = dist[j];
          bi = lane + 32 * j;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < best || (ov == best && oi < bi)) {
          best = ov;
          bi = oi;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (lane + 32 * j == bi) dist[j] = CUDART_INF_F;
      if (lane == 0 && (t % dilation) == 0) out[t / dilation] = bi;
    }
  }
}

// Max-relative aggregation (MRConv2d): m[b][c][n] = max_t ( y[b][c][nn_idx[b][n][t]] - x[b][c][n] ).
// interleave = 0: out [B][C][N] = m;  interleave = 1: out [B][2C][N] with channel 2c = x, 2c+1 = m (the
// channel-interleaved input of MRConv2d's grouped 1x1 conv).  One thread per (b, c, n), coalesced over n; the
// neighbour table of a node is re-read per channel from L1/L2, the y gathers stay inside one M-float row.
__global__ void __launch_bounds__(256) max_relative_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           const long long* __restrict__ nn_idx, int B, int C, int N, int M,
                                                           int k, int interleave, float* __restrict__ out) {
  const size_t total = static_cast<size_t>(B) * C * N;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t n = i % N;
    const size_t bc = i / N;
    const size_t c = bc % C, b = bc / C;
    const float xv = __ldg(x + i);
    const float* yrow = y + (b * C + c) * M;
    const long long* idx = nn_idx + (b * N + n) * k;
    float m = -CUDART_INF_F;
    for (int t = 0; t < k; ++t) {
      const long long j = __ldg(idx + t);
      m = fmaxf(m, __fsub_rn(__ldg(yrow + j), xv));
    }
    if (interleave) {
      float* o = out + (b * 2 * C + 2 * c) * N + n;
      o[0] = xv;
      o[N] = m;
    } else {
      out[i] = m;
    }
  }
}

}  // namespace stcd
