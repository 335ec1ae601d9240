// K12: BIT's token path (models/networks.py:359-394,414-428; blocks in models/help_funcs.py) on a 32-channel bf16 feature
// map [img][4][h*w][8], fp32 arithmetic.
//
//   bit_tokenizer_kernel   one CTA per image: tokens[l] = sum_n softmax_n(conv_a(x))[l, n] x[n]  (online softmax, one read)
//   bit_token_mixer_kernel one CTA per pair: + learned positions, the transformer encoder over the pair's 2L tokens, then per
//                          decoder layer and stream the collapsed cross-attention matrices
//                            A[c][(h, j)]  = scale * sum_d Wq[h*dh + d][c] * k[j][h*dh + d]       k = Wk LN(m)
//                            Bm[(h, j)][c] =         sum_d Wout[c][h*dh + d] * v[j][h*dh + d]     v = Wv LN(m)
//                          so that softmax_j(q_h . k_hj * scale) v_hj Wout^T == softmax_groups(LN(x) A) Bm.
//   bit_decoder_kernel     one thread per pixel, its 32 channels in registers through every decoder layer:
//                          x += softmax_groups(LN(x) A) Bm + bout;  x += W2 gelu(W1 LN(x) + b1) + b2
// C = 32 channels, L = 4 tokens, 8 heads (heads * L = 32 scores per pixel), mlp = 64: the only configuration the reference
// registers (models/networks.py:174-182); the host checks it.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "transformer_kernels.cuh"

namespace stcd {

constexpr int kBitC = 32, kBitL = 4, kBitHeads = 8, kBitMlp = 64, kBitHJ = kBitHeads * kBitL;
constexpr int kBitCoef = 2 * kBitC * kBitHJ;   // floats per (image, decoder layer): A | Bm

__host__ __device__ constexpr int bit_enc_size(int inner) {
  return 2 * kBitC + 3 * inner * kBitC + kBitC * inner + kBitC + 2 * kBitC + kBitMlp * kBitC + kBitMlp + kBitC * kBitMlp + kBitC;
}
__host__ __device__ constexpr int bit_dec_size(int inner) {
  return 2 * kBitC + 3 * inner * kBitC + kBitC * inner + kBitC + 2 * kBitC + kBitC * kBitMlp + kBitMlp + kBitMlp * kBitC + kBitC;
}
// offset of the per-pixel part of a decoder layer (bout | ln2_g | ln2_b | w1t | b1 | w2t | b2) and its length
__host__ __device__ constexpr int bit_dec_tail_off(int inner) { return 2 * kBitC + 3 * inner * kBitC + kBitC * inner; }
constexpr int kBitDecTail = kBitC + 2 * kBitC + kBitC * kBitMlp + kBitMlp + kBitMlp * kBitC + kBitC;

__device__ __forceinline__ float bit_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// ------------------------------------------------------------------------------------------ tokenizer
// grid = images, 256 threads.  State per thread and token: running max m, sum s, weighted channel sums acc[32].
__global__ void __launch_bounds__(256) bit_tokenizer_kernel(const __nv_bfloat16* __restrict__ src, const float* __restrict__ conv_a,
                                                            float* __restrict__ tokens, int src_c8, int hw) {
  __shared__ float s_wa[kBitL * kBitC];
  __shared__ float s_m[8][kBitL], s_s[8][kBitL], s_acc[8][kBitL][kBitC];
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < kBitL * kBitC) s_wa[threadIdx.x] = conv_a[threadIdx.x];
  __syncthreads();
  float m[kBitL], s[kBitL], acc[kBitL][kBitC];
#pragma unroll
  for (int l = 0; l < kBitL; ++l) {
    m[l] = -1e30f;
    s[l] = 0.f;
#pragma unroll
    for (int c = 0; c < kBitC; ++c) acc[l][c] = 0.f;
  }
  const __nv_bfloat16* base = src + static_cast<size_t>(img) * src_c8 * hw * 8;
  for (int n = threadIdx.x; n < hw; n += blockDim.x) {
    float x[kBitC];
#pragma unroll
    for (int g = 0; g < 4; ++g) unpack8(__ldg(reinterpret_cast<const uint4*>(base + (static_cast<size_t>(g) * hw + n) * 8)), x + 8 * g);
#pragma unroll
    for (int l = 0; l < kBitL; ++l) {
      float a = 0.f;
#pragma unroll
      for (int c = 0; c < kBitC; ++c) a = fmaf(s_wa[l * kBitC + c], x[c], a);
      const float mn = fmaxf(m[l], a);
      const float r = expf(m[l] - mn), e = expf(a - mn);
      s[l] = s[l] * r + e;
#pragma unroll
      for (int c = 0; c < kBitC; ++c) acc[l][c] = fmaf(acc[l][c], r, e * x[c]);
      m[l] = mn;
    }
  }
  // butterfly over the warp (fixed order), then over the 8 warps through shared memory
#pragma unroll
  for (int l = 0; l < kBitL; ++l) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[l], o), so = __shfl_xor_sync(0xffffffffu, s[l], o);
      const float mn = fmaxf(m[l], mo);
      const float r = expf(m[l] - mn), ro = expf(mo - mn);
      s[l] = s[l] * r + so * ro;
#pragma unroll
      for (int c = 0; c < kBitC; ++c) acc[l][c] = acc[l][c] * r + __shfl_xor_sync(0xffffffffu, acc[l][c], o) * ro;
      m[l] = mn;
    }
    if (lane == 0) {
      s_m[warp][l] = m[l];
      s_s[warp][l] = s[l];
    }
#pragma unroll
    for (int c = 0; c < kBitC; ++c)
      if (lane == c) s_acc[warp][l][c] = acc[l][c];
  }
  __syncthreads();
  if (threadIdx.x < kBitL * kBitC) {
    const int l = threadIdx.x / kBitC, c = threadIdx.x % kBitC;
    float M = -1e30f;
    for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w][l]);
    float S = 0.f, A = 0.f;
    for (int w = 0; w < 8; ++w) {
      const float r = expf(s_m[w][l] - M);
      S = fmaf(s_s[w][l], r, S);
      A = fmaf(s_acc[w][l][c], r, A);
    }
    tokens[(static_cast<size_t>(img) * kBitL + l) * kBitC + c] = A / S;
  }
}

// ------------------------------------------------------------------------------------------ token mixer
__device__ __forceinline__ void bit_ln_rows(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ g,
                                            const float* __restrict__ b, int rows) {
  // one warp per row, lane = channel (C = 32)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rows; r += 8) {
    const float v = in[r * kBitC + lane];
    float mean = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mean += __shfl_xor_sync(0xffffffffu, mean, o);
    mean *= (1.f / kBitC);
    const float d = v - mean;
    float var = d * d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    out[r * kBitC + lane] = d * rsqrtf(var * (1.f / kBitC) + 1e-5f) * __ldg(g + lane) + __ldg(b + lane);
  }
}

// grid = pairs (chunk), 256 threads; dynamic smem: big[8 * 3 * inner_max] floats
__global__ void __launch_bounds__(256) bit_token_mixer_kernel(const float* __restrict__ tokens, const float* __restrict__ pos,
                                                              const float* __restrict__ enc, const float* __restrict__ dec,
                                                              float* __restrict__ coef, int chunk, int n_enc, int n_dec,
                                                              int inner_e, int inner_d, float scale) {
  extern __shared__ float s_big[];                 // encoder: qkv [8][3*inner_e]; decoder: k | v [2][4][inner_d]
  __shared__ float s_t[2 * kBitL * kBitC], s_y[2 * kBitL * kBitC], s_attn[kBitHeads][2 * kBitL][2 * kBitL], s_h[2 * kBitL * kBitMlp];
  const int pair = blockIdx.x, tid = threadIdx.x;
  constexpr int R = 2 * kBitL;                      // 8 token rows per pair
  {
    const int r = tid / kBitC, c = tid % kBitC;     // 256 threads == R * C
    const int img = (r < kBitL ? 0 : chunk) + pair;
    s_t[tid] = tokens[(static_cast<size_t>(img) * kBitL + (r % kBitL)) * kBitC + c] + __ldg(pos + tid);
  }
  __syncthreads();
  for (int l = 0; l < n_enc; ++l) {
    const float* P = enc + static_cast<size_t>(l) * bit_enc_size(inner_e);
    const float* ln1_g = P;
    const float* ln1_b = ln1_g + kBitC;
    const float* wqkv = ln1_b + kBitC;
    const float* wout = wqkv + 3 * inner_e * kBitC;
    const float* bout = wout + kBitC * inner_e;
    const float* ln2_g = bout + kBitC;
    const float* ln2_b = ln2_g + kBitC;
    const float* w1 = ln2_b + kBitC;
    const float* b1 = w1 + kBitMlp * kBitC;
    const float* w2 = b1 + kBitMlp;
    const float* b2 = w2 + kBitC * kBitMlp;
    bit_ln_rows(s_t, s_y, ln1_g, ln1_b, R);
    __syncthreads();
    for (int j = tid; j < 3 * inner_e; j += blockDim.x) {       // qkv[r][j] = y[r] . Wqkv[j]
      float w[kBitC];
#pragma unroll
      for (int q = 0; q < kBitC / 4; ++q) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(wqkv + static_cast<size_t>(j) * kBitC) + q);
        w[4 * q] = t4.x, w[4 * q + 1] = t4.y, w[4 * q + 2] = t4.z, w[4 * q + 3] = t4.w;
      }
      for (int r = 0; r < R; ++r) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < kBitC; ++c) a = fmaf(s_y[r * kBitC + c], w[c], a);
        s_big[r * 3 * inner_e + j] = a;
      }
    }
    __syncthreads();
    const int dh = inner_e / kBitHeads;
    if (tid < kBitHeads * R) {                                  // one (head, query) per thread: 8 scores, softmax
      const int h = tid / R, i = tid % R;
      float d[R], mx = -1e30f;
      for (int j = 0; j < R; ++j) {
        float a = 0.f;
        for (int e = 0; e < dh; ++e) a = fmaf(s_big[i * 3 * inner_e + h * dh + e], s_big[j * 3 * inner_e + inner_e + h * dh + e], a);
        d[j] = a * scale;
        mx = fmaxf(mx, d[j]);
      }
      float sum = 0.f;
      for (int j = 0; j < R; ++j) {
        d[j] = expf(d[j] - mx);
        sum += d[j];
      }
      for (int j = 0; j < R; ++j) s_attn[h][i][j] = d[j] / sum;
    }
    __syncthreads();
    // o[i][col] = sum_j attn[h(col)][i][j] v[j][col], written over q (each (i, col) is read by nobody else after the scores)
    for (int it = tid; it < R * inner_e; it += blockDim.x) {
      const int i = it / inner_e, col = it % inner_e, h = col / dh;
      float a = 0.f;
      for (int j = 0; j < R; ++j) a = fmaf(s_attn[h][i][j], s_big[j * 3 * inner_e + 2 * inner_e + col], a);
      s_big[i * 3 * inner_e + col] = a;
    }
    __syncthreads();
    {
      const int r = tid / kBitC, c = tid % kBitC;
      float a = __ldg(bout + c);
      for (int j = 0; j < inner_e; ++j) a = fmaf(s_big[r * 3 * inner_e + j], __ldg(wout + static_cast<size_t>(c) * inner_e + j), a);
      s_t[tid] += a;
    }
    __syncthreads();
    bit_ln_rows(s_t, s_y, ln2_g, ln2_b, R);
    __syncthreads();
    for (int it = tid; it < R * kBitMlp; it += blockDim.x) {
      const int r = it / kBitMlp, k = it % kBitMlp;
      float a = __ldg(b1 + k);
#pragma unroll
      for (int c = 0; c < kBitC; ++c) a = fmaf(s_y[r * kBitC + c], __ldg(w1 + k * kBitC + c), a);
      s_h[it] = bit_gelu(a);
    }
    __syncthreads();
    {
      const int r = tid / kBitC, c = tid % kBitC;
      float a = __ldg(b2 + c);
#pragma unroll
      for (int k = 0; k < kBitMlp; ++k) a = fmaf(s_h[r * kBitMlp + k], __ldg(w2 + c * kBitMlp + k), a);
      s_t[tid] += a;
    }
    __syncthreads();
  }
  // decoder: the tokens are the memory of every layer (they do not change); collapse each layer's cross-attention
  const int dh = inner_d / kBitHeads;
  for (int l = 0; l < n_dec; ++l) {
    const float* P = dec + static_cast<size_t>(l) * bit_dec_size(inner_d);
    const float* ln1_g = P;
    const float* ln1_b = ln1_g + kBitC;
    const float* wq = ln1_b + kBitC;
    const float* wk = wq + inner_d * kBitC;
    const float* wv = wk + inner_d * kBitC;
    const float* wout = wv + inner_d * kBitC;
    bit_ln_rows(s_t, s_y, ln1_g, ln1_b, R);                     // PreNorm2: the layer's norm on the memory too
    __syncthreads();
    float* s_k = s_big;                                         // [R][inner_d]
    float* s_v = s_big + R * inner_d;
    for (int j = tid; j < 2 * inner_d; j += blockDim.x) {
      const float* wrow = (j < inner_d ? wk + static_cast<size_t>(j) * kBitC : wv + static_cast<size_t>(j - inner_d) * kBitC);
      float w[kBitC];
#pragma unroll
      for (int q = 0; q < kBitC / 4; ++q) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(wrow) + q);
        w[4 * q] = t4.x, w[4 * q + 1] = t4.y, w[4 * q + 2] = t4.z, w[4 * q + 3] = t4.w;
      }
      for (int r = 0; r < R; ++r) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < kBitC; ++c) a = fmaf(s_y[r * kBitC + c], w[c], a);
        (j < inner_d ? s_k[r * inner_d + j] : s_v[r * inner_d + j - inner_d]) = a;
      }
    }
    __syncthreads();
    for (int it = tid; it < 2 * kBitC * kBitHJ; it += blockDim.x) {       // (stream, c, hj): A and Bm
      const int s = it / (kBitC * kBitHJ), rem = it % (kBitC * kBitHJ);
      const int hj = rem / kBitC, c = rem % kBitC;                       // lanes vary c: Wq reads coalesce
      const int h = hj / kBitL, j = hj % kBitL, row = s * kBitL + j;
      float a = 0.f, b = 0.f;
      for (int e = 0; e < dh; ++e) {
        a = fmaf(__ldg(wq + static_cast<size_t>(h * dh + e) * kBitC + c), s_k[row * inner_d + h * dh + e], a);
        b = fmaf(__ldg(wout + static_cast<size_t>(c) * inner_d + h * dh + e), s_v[row * inner_d + h * dh + e], b);
      }
      float* o = coef + (static_cast<size_t>(s * chunk + pair) * n_dec + l) * kBitCoef;
      o[c * kBitHJ + hj] = a * scale;
      o[kBitC * kBitHJ + hj * kBitC + c] = b;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ decoder
__device__ __forceinline__ void bit_ln32(const float (&x)[kBitC], float (&y)[kBitC], const float* __restrict__ g, const float* __restrict__ b) {
  float mean = 0.f;
#pragma unroll
  for (int c = 0; c < kBitC; ++c) mean += x[c];
  mean *= (1.f / kBitC);
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < kBitC; ++c) var = fmaf(x[c] - mean, x[c] - mean, var);
  const float rs = rsqrtf(var * (1.f / kBitC) + 1e-5f);
#pragma unroll
  for (int c = 0; c < kBitC; ++c) y[c] = (x[c] - mean) * rs * g[c] + b[c];
}

// out[N] += in[K] . W[K][N] (row-major rows in shared memory, broadcast float4 reads)
template <int K, int N>
__device__ __forceinline__ void bit_matvec(const float (&in)[K], const float* __restrict__ W, float (&out)[N]) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int n4 = 0; n4 < N / 4; ++n4) {
      const float4 w = *reinterpret_cast<const float4*>(W + k * N + 4 * n4);
      out[4 * n4] = fmaf(in[k], w.x, out[4 * n4]);
      out[4 * n4 + 1] = fmaf(in[k], w.y, out[4 * n4 + 1]);
      out[4 * n4 + 2] = fmaf(in[k], w.z, out[4 * n4 + 2]);
      out[4 * n4 + 3] = fmaf(in[k], w.w, out[4 * n4 + 3]);
    }
  }
}

// grid (pixel blocks, images), 256 threads
__global__ void __launch_bounds__(256) bit_decoder_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                          const float* __restrict__ dec, const float* __restrict__ coef, int src_c8,
                                                          int dst_c8, int hw, int n_dec, int inner_d, int softmax) {
  __shared__ __align__(16) float s_ln1[2 * kBitC];
  __shared__ __align__(16) float s_coef[kBitCoef];
  __shared__ __align__(16) float s_tail[kBitDecTail];
  const int img = blockIdx.y, pix = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = pix < hw;
  float x[kBitC];
  if (live) {
    const __nv_bfloat16* b = src + static_cast<size_t>(img) * src_c8 * hw * 8;
#pragma unroll
    for (int g = 0; g < 4; ++g) unpack8(__ldg(reinterpret_cast<const uint4*>(b + (static_cast<size_t>(g) * hw + pix) * 8)), x + 8 * g);
  } else {
#pragma unroll
    for (int c = 0; c < kBitC; ++c) x[c] = 0.f;
  }
  for (int l = 0; l < n_dec; ++l) {
    const float* P = dec + static_cast<size_t>(l) * bit_dec_size(inner_d);
    const float* cf = coef + (static_cast<size_t>(img) * n_dec + l) * kBitCoef;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * kBitC; i += blockDim.x) s_ln1[i] = __ldg(P + i);
    for (int i = threadIdx.x; i < kBitCoef / 4; i += blockDim.x)
      reinterpret_cast<float4*>(s_coef)[i] = __ldg(reinterpret_cast<const float4*>(cf) + i);
    for (int i = threadIdx.x; i < kBitDecTail; i += blockDim.x) s_tail[i] = __ldg(P + bit_dec_tail_off(inner_d) + i);
    __syncthreads();
    const float* bout = s_tail;
    const float* ln2_g = bout + kBitC;
    const float* ln2_b = ln2_g + kBitC;
    const float* w1t = ln2_b + kBitC;
    const float* b1 = w1t + kBitC * kBitMlp;
    const float* w2t = b1 + kBitMlp;
    const float* b2 = w2t + kBitMlp * kBitC;
    float y[kBitC];
    bit_ln32(x, y, s_ln1, s_ln1 + kBitC);
    {
      float d[kBitHJ];
#pragma unroll
      for (int i = 0; i < kBitHJ; ++i) d[i] = 0.f;
      bit_matvec<kBitC, kBitHJ>(y, s_coef, d);
      if (softmax) {
#pragma unroll
        for (int h = 0; h < kBitHeads; ++h) {
          const float mx = fmaxf(fmaxf(d[4 * h], d[4 * h + 1]), fmaxf(d[4 * h + 2], d[4 * h + 3]));
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < kBitL; ++j) {
            d[4 * h + j] = __expf(d[4 * h + j] - mx);
            sum += d[4 * h + j];
          }
          const float inv = 1.f / sum;
#pragma unroll
          for (int j = 0; j < kBitL; ++j) d[4 * h + j] *= inv;
        }
      }
#pragma unroll
      for (int c = 0; c < kBitC; ++c) x[c] += bout[c];
      bit_matvec<kBitHJ, kBitC>(d, s_coef + kBitC * kBitHJ, x);
    }
    bit_ln32(x, y, ln2_g, ln2_b);
    {
      float hbuf[kBitMlp];
#pragma unroll
      for (int k = 0; k < kBitMlp; ++k) hbuf[k] = b1[k];
      bit_matvec<kBitC, kBitMlp>(y, w1t, hbuf);
#pragma unroll
      for (int k = 0; k < kBitMlp; ++k) hbuf[k] = bit_gelu(hbuf[k]);
#pragma unroll
      for (int c = 0; c < kBitC; ++c) x[c] += b2[c];
      bit_matvec<kBitMlp, kBitC>(hbuf, w2t, x);
    }
  }
  if (live) {
    __nv_bfloat16* o = dst + static_cast<size_t>(img) * dst_c8 * hw * 8;
#pragma unroll
    for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(o + (static_cast<size_t>(g) * hw + pix) * 8) = pack8(x + 8 * g);
  }
}

}  // namespace stcd
