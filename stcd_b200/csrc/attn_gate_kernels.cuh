// K13: DSIFN's attention gates (models/DSIFN.py:24-51) as bandwidth kernels on bf16 plan tensors [img][c/8][h*w][8], fp32 arithmetic.
//
//   channel attention  `ca(x) * x` over a VIRTUAL concat x = cat(decoder map, T1 feature, T2 feature) (:138-140 ...): per segment
//                      chan_stats_kernel (sum + max per image and channel, one warp per (image, 8 channels, pixel range)), once
//                      ca_fc_kernel (sigmoid(fc2(relu(fc1 avg)) + fc2(relu(fc1 max))), one CTA per image), per segment
//                      ca_apply_kernel, which writes the scaled segment into the materialised concat the next conv reads.
//   spatial attention  `bn(sa(x) * x)` (:131-132 ...): sa_stats_kernel (mean and max over the channels of every pixel) and
//                      sa_apply_kernel (7x7 conv over the 2-channel statistics map staged in shared memory, sigmoid, gate,
//                      folded BatchNorm).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "transformer_kernels.cuh"

namespace stcd {

constexpr int kCaMaxC = 2048, kCaMaxH = 256;

// one warp per (pixel range, 8-channel group, image) of ONE segment; partial_{sum,max}[(img * ranges + r) * c_tot + c_off + ...]
__global__ void __launch_bounds__(256) chan_stats_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ psum,
                                                         float* __restrict__ pmax, int c_seg, int src_c8, int hw, int ranges, int n_items,
                                                         int c_tot, int c_off) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item >= n_items) return;
  const int g8 = c_seg >> 3;
  const int r = item % ranges, g = (item / ranges) % g8, b = item / (ranges * g8);
  const int per = (hw + ranges - 1) / ranges;
  const int p0 = r * per, p1 = min(hw, p0 + per);
  float s[8], m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f, m[j] = -3.0e38f;
  const __nv_bfloat16* base = src + (static_cast<size_t>(b) * src_c8 + g) * static_cast<size_t>(hw) * 8;
#pragma unroll 4
  for (int px = p0 + lane; px < p1; px += 32) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(px) * 8)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += v[j], m[j] = fmaxf(m[j], v[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
      m[j] = fmaxf(m[j], __shfl_xor_sync(0xffffffffu, m[j], o));
    }
  }
  if (lane < 8) {
    float a = s[0], mm = m[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) a = lane == j ? s[j] : a, mm = lane == j ? m[j] : mm;
    const size_t o = (static_cast<size_t>(b) * ranges + r) * c_tot + c_off + g * 8 + lane;
    psum[o] = a;
    pmax[o] = mm;
  }
}

// grid (ceil(C / 256), images), 256 threads: gate[img][c] = sigmoid(fc2 (relu(fc1 avg) + relu(fc1 max)))  (fc2 is linear and
// bias-free).  Every CTA redoes the hidden layer (hid x C MACs, warp-cooperative, coalesced fc1 rows) and finishes 256 channels;
// fc2t is fc2 transposed to [hid][C] so those reads coalesce too.
__global__ void __launch_bounds__(256) ca_fc_kernel(const float* __restrict__ psum, const float* __restrict__ pmax, const float* __restrict__ fc1,
                                                    const float* __restrict__ fc2t, float* __restrict__ gate, int C, int hid, int hw, int ranges) {
  __shared__ float s_avg[kCaMaxC], s_max[kCaMaxC], s_hid[kCaMaxH];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, m = -3.0e38f;
    for (int r = 0; r < ranges; ++r) {
      a += psum[(static_cast<size_t>(b) * ranges + r) * C + c];
      m = fmaxf(m, pmax[(static_cast<size_t>(b) * ranges + r) * C + c]);
    }
    s_avg[c] = a / static_cast<float>(hw);
    s_max[c] = m;
  }
  __syncthreads();
  for (int u = warp; u < hid; u += 8) {
    float a = 0.f, m = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float w = __ldg(fc1 + static_cast<size_t>(u) * C + c);
      a = fmaf(w, s_avg[c], a);
      m = fmaf(w, s_max[c], m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      m += __shfl_xor_sync(0xffffffffu, m, o);
    }
    if (lane == 0) s_hid[u] = fmaxf(a, 0.f) + fmaxf(m, 0.f);
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    float a = 0.f;
    for (int u = 0; u < hid; ++u) a = fmaf(__ldg(fc2t + static_cast<size_t>(u) * C + c), s_hid[u], a);
    gate[static_cast<size_t>(b) * C + c] = 1.f / (1.f + expf(-a));
  }
}

// dst[img][c_off/8 + g][pix] = src[img][g][pix] * gate[img][c_off + 8 g ..]: one segment, grid-stride over 16-byte vectors
__global__ void __launch_bounds__(256) ca_apply_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                       const float* __restrict__ gate, int imgs, int c_seg, int src_c8, int dst_c8, int hw,
                                                       int c_tot, int c_off) {
  const int g8 = c_seg >> 3;
  const size_t total = static_cast<size_t>(imgs) * g8 * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int pix = static_cast<int>(i % hw);
    const int g = static_cast<int>((i / hw) % g8), b = static_cast<int>(i / (static_cast<size_t>(hw) * g8));
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + ((static_cast<size_t>(b) * src_c8 + g) * hw + pix) * 8)), v);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + static_cast<size_t>(b) * c_tot + c_off + g * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + static_cast<size_t>(b) * c_tot + c_off + g * 8) + 1);
    v[0] *= g0.x, v[1] *= g0.y, v[2] *= g0.z, v[3] *= g0.w, v[4] *= g1.x, v[5] *= g1.y, v[6] *= g1.z, v[7] *= g1.w;
    *reinterpret_cast<uint4*>(dst + ((static_cast<size_t>(b) * dst_c8 + (c_off >> 3) + g) * hw + pix) * 8) = pack8(v);
  }
}

// stats[img][pix] = (mean_c x, max_c x); one thread per pixel
__global__ void __launch_bounds__(256) sa_stats_kernel(const __nv_bfloat16* __restrict__ src, float2* __restrict__ stats, int C, int src_c8, int hw) {
  const int b = blockIdx.y, pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= hw) return;
  const __nv_bfloat16* s = src + (static_cast<size_t>(b) * src_c8 * hw + pix) * 8;
  float sum = 0.f, mx = -3.0e38f;
  for (int g = 0; g < (C >> 3); ++g) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(s + static_cast<size_t>(g) * hw * 8)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += v[j], mx = fmaxf(mx, v[j]);
  }
  stats[static_cast<size_t>(b) * hw + pix] = make_float2(sum / static_cast<float>(C), mx);
}

// block (32, 8) pixels; grid (ceil(w/32), ceil(h/8), images).  wgt: [2][7][7] (mean plane, max plane), scale/shift: folded BN
constexpr int kSaTW = 32, kSaTH = 8;
__global__ void __launch_bounds__(kSaTW* kSaTH) sa_apply_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                const float2* __restrict__ stats, const float* __restrict__ wgt,
                                                                const float* __restrict__ scale, const float* __restrict__ shift, int C,
                                                                int src_c8, int dst_c8, int h, int w) {
  __shared__ float2 s_t[kSaTH + 6][kSaTW + 6];
  __shared__ float s_w[98];
  const int b = blockIdx.z, x0 = blockIdx.x * kSaTW, y0 = blockIdx.y * kSaTH, hw = h * w;
  const int tid = threadIdx.y * kSaTW + threadIdx.x;
  if (tid < 98) s_w[tid] = wgt[tid];
  for (int i = tid; i < (kSaTH + 6) * (kSaTW + 6); i += kSaTW * kSaTH) {
    const int ty = i / (kSaTW + 6), tx = i % (kSaTW + 6);
    const int gy = y0 + ty - 3, gx = x0 + tx - 3;
    s_t[ty][tx] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? stats[static_cast<size_t>(b) * hw + gy * w + gx] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= w || y >= h) return;
  float a = 0.f;
#pragma unroll
  for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const float2 t = s_t[threadIdx.y + ky][threadIdx.x + kx];
      a = fmaf(t.x, s_w[ky * 7 + kx], a);
      a = fmaf(t.y, s_w[49 + ky * 7 + kx], a);
    }
  }
  const float gs = 1.f / (1.f + expf(-a));
  const int pix = y * w + x;
  const __nv_bfloat16* s = src + (static_cast<size_t>(b) * src_c8 * hw + pix) * 8;
  __nv_bfloat16* o = dst + (static_cast<size_t>(b) * dst_c8 * hw + pix) * 8;
  for (int g = 0; g < (C >> 3); ++g) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(s + static_cast<size_t>(g) * hw * 8)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] * gs, __ldg(scale + g * 8 + j), __ldg(shift + g * 8 + j));
    *reinterpret_cast<uint4*>(o + static_cast<size_t>(g) * hw * 8) = pack8(v);
  }
}

// ------------------------------------------------------------------------------------------ ChangeGNNV2 decoder gates
// Global_Local's global branch (models/ChangeVIG.py:377-385): out = sigmoid(ch[c] * sp[pixel]) * x with
//   ch[c] = relu((w0[c] avg_c + w1[c] max_c) * cs[c] + ct[c])      (grouped (2,1) conv over [avg; max], bias + BatchNorm folded)
//   sp    = relu(conv5x5([mean_c x, max_c x]) + b)
// prm: w0[C] | w1[C] | cs[C] | ct[C] | wsp[2][5][5] | b.  block (32, 8) pixels; grid (ceil(w/32), ceil(h/8), images).
constexpr int kGlMaxC = 512;
__global__ void __launch_bounds__(kSaTW* kSaTH) gl_apply_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                const float2* __restrict__ stats, const float* __restrict__ psum,
                                                                const float* __restrict__ pmax, const float* __restrict__ prm, int C,
                                                                int src_c8, int dst_c8, int h, int w, int ranges) {
  __shared__ float2 s_t[kSaTH + 4][kSaTW + 4];
  __shared__ float s_w[51];
  __shared__ float s_ch[kGlMaxC];
  const int b = blockIdx.z, x0 = blockIdx.x * kSaTW, y0 = blockIdx.y * kSaTH, hw = h * w;
  const int tid = threadIdx.y * kSaTW + threadIdx.x;
  if (tid < 51) s_w[tid] = prm[4 * C + tid];
  for (int c = tid; c < C; c += kSaTW * kSaTH) {
    float a = 0.f, m = -3.0e38f;
    for (int r = 0; r < ranges; ++r) {
      a += psum[(static_cast<size_t>(b) * ranges + r) * C + c];
      m = fmaxf(m, pmax[(static_cast<size_t>(b) * ranges + r) * C + c]);
    }
    a /= static_cast<float>(hw);
    s_ch[c] = fmaxf(fmaf(fmaf(prm[c], a, prm[C + c] * m), prm[2 * C + c], prm[3 * C + c]), 0.f);
  }
  for (int i = tid; i < (kSaTH + 4) * (kSaTW + 4); i += kSaTW * kSaTH) {
    const int ty = i / (kSaTW + 4), tx = i % (kSaTW + 4);
    const int gy = y0 + ty - 2, gx = x0 + tx - 2;
    s_t[ty][tx] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? stats[static_cast<size_t>(b) * hw + gy * w + gx] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= w || y >= h) return;
  float a = s_w[50];
#pragma unroll
  for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 5; ++kx) {
      const float2 t = s_t[threadIdx.y + ky][threadIdx.x + kx];
      a = fmaf(t.x, s_w[ky * 5 + kx], a);
      a = fmaf(t.y, s_w[25 + ky * 5 + kx], a);
    }
  }
  const float sp = fmaxf(a, 0.f);
  const int pix = y * w + x;
  const __nv_bfloat16* s = src + (static_cast<size_t>(b) * src_c8 * hw + pix) * 8;
  __nv_bfloat16* o = dst + (static_cast<size_t>(b) * dst_c8 * hw + pix) * 8;
  for (int g = 0; g < (C >> 3); ++g) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(s + static_cast<size_t>(g) * hw * 8)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= 1.f / (1.f + expf(-s_ch[g * 8 + j] * sp));
    *reinterpret_cast<uint4*>(o + static_cast<size_t>(g) * hw * 8) = pack8(v);
  }
}

// VFFM (models/ChangeVIG.py:452-460): xo = 2 low wei + 2 high (1 - wei), wei = sigmoid(g_avg[c] + g_max[c] + local[pixel, c]);
// g_* = conv-BN-ReLU-conv-BN (1x1, on the pooled vector of mixed = low + high), BatchNorm and biases folded to scale / shift.
// prm: avg branch then max branch, each  w1[inter][C] | s1[inter] | t1[inter] | w2t[inter][C] | s2[C] | t2[C].
// grid (pixel blocks, images), 256 threads; every CTA redoes the two small MLPs of its image.
__global__ void __launch_bounds__(256) vffm_apply_kernel(const __nv_bfloat16* __restrict__ low, const __nv_bfloat16* __restrict__ high,
                                                         const __nv_bfloat16* __restrict__ local, __nv_bfloat16* __restrict__ dst,
                                                         const float* __restrict__ psum, const float* __restrict__ pmax,
                                                         const float* __restrict__ prm, int C, int inter, int hw, int ranges,
                                                         int pix_per_block) {
  __shared__ float s_v[2][kGlMaxC], s_h[2][kGlMaxC / 4], s_g[kGlMaxC];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g8 = C >> 3;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, m = -3.0e38f;
    for (int r = 0; r < ranges; ++r) {
      a += psum[(static_cast<size_t>(b) * ranges + r) * C + c];
      m = fmaxf(m, pmax[(static_cast<size_t>(b) * ranges + r) * C + c]);
    }
    s_v[0][c] = a / static_cast<float>(hw);
    s_v[1][c] = m;
  }
  __syncthreads();
  const int bsz = 2 * inter * C + 2 * inter + 2 * C;            // floats per branch
  for (int it = warp; it < 2 * inter; it += 8) {
    const int br = it / inter, u = it % inter;
    const float* P = prm + br * bsz;
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(__ldg(P + static_cast<size_t>(u) * C + c), s_v[br][c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) s_h[br][u] = fmaxf(fmaf(a, P[inter * C + u], P[inter * C + inter + u]), 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float tot = 0.f;
    for (int br = 0; br < 2; ++br) {
      const float* P = prm + br * bsz + inter * C + 2 * inter;   // w2t | s2 | t2
      float a = 0.f;
      for (int u = 0; u < inter; ++u) a = fmaf(__ldg(P + static_cast<size_t>(u) * C + c), s_h[br][u], a);
      tot += fmaf(a, P[inter * C + c], P[inter * C + C + c]);
    }
    s_g[c] = tot;
  }
  __syncthreads();
  const int p_begin = blockIdx.x * pix_per_block;
  const int n_pix = min(hw, p_begin + pix_per_block) - p_begin;
  for (int it = threadIdx.x; it < n_pix * g8; it += blockDim.x) {
    const int g = it / n_pix, pix = p_begin + it - g * n_pix;
    const size_t off = ((static_cast<size_t>(b) * g8 + g) * hw + pix) * 8;
    float lo[8], hi[8], lc[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(low + off)), lo);
    unpack8(__ldg(reinterpret_cast<const uint4*>(high + off)), hi);
    unpack8(__ldg(reinterpret_cast<const uint4*>(local + off)), lc);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float wei = 1.f / (1.f + expf(-(s_g[g * 8 + j] + lc[j])));
      lo[j] = 2.f * lo[j] * wei + 2.f * hi[j] * (1.f - wei);
    }
    *reinterpret_cast<uint4*>(dst + off) = pack8(lo);
  }
}

// csam_V20 (models/ChangeVIG.py:956-994): out = bt((sigmoid(ch[c]) + sigmoid(sp[pixel])) * x),
//   ch = liner2(relu(liner1(gelu(BN(grouped (2,1) conv over [avg; max])))))        per image
//   sp = conv3x3(relu(conv3x3([mean_c x, max_c x])))                                both bias-free, zero padding between them
// prm: w_avg[C] | w_max[C] | cs[C] | ct[C] | l1[hid][C] | l2t[hid][C] | b2[C] | bt_s[C] | bt_t[C] | w21[2][3][3] | w22[3][3].
// block (32, 8) pixels; grid (ceil(w/32), ceil(h/8), images); every CTA redoes the C x hid channel MLP of its image.
__global__ void __launch_bounds__(kSaTW* kSaTH) csam_apply_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                  const float2* __restrict__ stats, const float* __restrict__ psum,
                                                                  const float* __restrict__ pmax, const float* __restrict__ prm, int C,
                                                                  int hid, int src_c8, int dst_c8, int h, int w, int ranges) {
  __shared__ float2 s_t[kSaTH + 4][kSaTW + 4];
  __shared__ float s_m[kSaTH + 2][kSaTW + 2];
  __shared__ float s_w[27];
  __shared__ float s_v[kGlMaxC], s_ch[kGlMaxC], s_h[kGlMaxC / 4];
  const int b = blockIdx.z, x0 = blockIdx.x * kSaTW, y0 = blockIdx.y * kSaTH, hw = h * w;
  const int tid = threadIdx.y * kSaTW + threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* l1 = prm + 4 * C;
  const float* l2t = l1 + static_cast<size_t>(hid) * C;
  const float* b2 = l2t + static_cast<size_t>(hid) * C;
  const float* bt_s = b2 + C;
  const float* bt_t = bt_s + C;
  if (tid < 27) s_w[tid] = bt_t[C + tid];
  for (int c = tid; c < C; c += kSaTW * kSaTH) {
    float a = 0.f, m = -3.0e38f;
    for (int r = 0; r < ranges; ++r) {
      a += psum[(static_cast<size_t>(b) * ranges + r) * C + c];
      m = fmaxf(m, pmax[(static_cast<size_t>(b) * ranges + r) * C + c]);
    }
    a /= static_cast<float>(hw);
    const float z = fmaf(fmaf(prm[c], a, prm[C + c] * m), prm[2 * C + c], prm[3 * C + c]);
    s_v[c] = 0.5f * z * (1.f + erff(z * 0.70710678118654752f));
  }
  for (int i = tid; i < (kSaTH + 4) * (kSaTW + 4); i += kSaTW * kSaTH) {
    const int ty = i / (kSaTW + 4), tx = i % (kSaTW + 4);
    const int gy = y0 + ty - 2, gx = x0 + tx - 2;
    s_t[ty][tx] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? stats[static_cast<size_t>(b) * hw + gy * w + gx] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int u = warp; u < hid; u += 8) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(__ldg(l1 + static_cast<size_t>(u) * C + c), s_v[c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) s_h[u] = fmaxf(a, 0.f);
  }
  // first 3x3 conv + ReLU on the tile + 1 halo; positions outside the image are the second conv's zero padding
  for (int i = tid; i < (kSaTH + 2) * (kSaTW + 2); i += kSaTW * kSaTH) {
    const int ty = i / (kSaTW + 2), tx = i % (kSaTW + 2);
    const int gy = y0 + ty - 1, gx = x0 + tx - 1;
    float a = 0.f;
    if (gy >= 0 && gy < h && gx >= 0 && gx < w) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float2 t = s_t[ty + ky][tx + kx];
          a = fmaf(t.x, s_w[ky * 3 + kx], a);
          a = fmaf(t.y, s_w[9 + ky * 3 + kx], a);
        }
      }
      a = fmaxf(a, 0.f);
    }
    s_m[ty][tx] = a;
  }
  __syncthreads();
  for (int c = tid; c < C; c += kSaTW * kSaTH) {
    float a = b2[c];
    for (int u = 0; u < hid; ++u) a = fmaf(__ldg(l2t + static_cast<size_t>(u) * C + c), s_h[u], a);
    s_ch[c] = 1.f / (1.f + expf(-a));
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= w || y >= h) return;
  float a = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) a = fmaf(s_m[threadIdx.y + ky][threadIdx.x + kx], s_w[18 + ky * 3 + kx], a);
  }
  const float sp = 1.f / (1.f + expf(-a));
  const int pix = y * w + x;
  const __nv_bfloat16* s = src + (static_cast<size_t>(b) * src_c8 * hw + pix) * 8;
  __nv_bfloat16* o = dst + (static_cast<size_t>(b) * dst_c8 * hw + pix) * 8;
  for (int g = 0; g < (C >> 3); ++g) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(s + static_cast<size_t>(g) * hw * 8)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf((s_ch[g * 8 + j] + sp) * v[j], __ldg(bt_s + g * 8 + j), __ldg(bt_t + g * 8 + j));
    *reinterpret_cast<uint4*>(o + static_cast<size_t>(g) * hw * 8) = pack8(v);
  }
}

}  // namespace stcd
