// K1: persistent, warp-specialised implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels x N channels] (TMEM, fp32) += A[128 x 16] * W[N x 16]^T   per filter tap and 16 channels
//
// Data layout.  Activations live in HBM as bf16 [img][c/8][h][w][8].  One TMA box
// {8 ch, box_w, box_h, kc/8, 1 img} therefore lands in shared memory as [kc/8][box_h][box_w][8 ch]:
// the un-swizzled K-major "core matrix" layout of tcgen05.mma (8 rows x 16 bytes contiguous), with
//   SBO (next 8 tile rows = next tile line)  = box_w * 16 bytes
//   LBO (next 8 channels)                    = box_h * box_w * 16 bytes
// A CTA tile is 16 x 8 output pixels (M = 128; 8 consecutive pixels of a line form a core matrix),
// so the A operand of filter tap (ty, tx) is the SAME box at byte offset (ty*box_w + tx)*16: a 3x3
// conv fetches its 18x10 halo once per K-chunk and issues 9 MMAs from it (L2->SM traffic 1.4x
// instead of 9x).  Out-of-bounds box pixels are zero-filled by TMA (= the conv's zero padding).
// Strided convs use element strides in the tensor map and one box per tap.
//
// Weights are host-packed per (N tile, phase) in consumption order as [block][kc/8][n_tile][8]
// (block = one tap of one chunk; SBO = 128 B, LBO = n_tile*16 B) and arrive by 1-D bulk copies:
// either ALL blocks once per CTA (weight-stationary: the CTA then streams pixel tiles past them)
// or through a ring when they do not fit.
//
// Warp roles (7 warps): 0 = A producer (TMA), 1 = W producer (bulk copy), 2 = MMA issuer
// (one lane; also owns TMEM alloc/dealloc), 3..6 = epilogue (TMEM -> registers -> folded BN /
// ReLU / residual / |f1-f2| / 2x2 max-pool -> global).  Accumulators are double-buffered in TMEM
// so the epilogue of tile i overlaps the MMAs of tile i+1; CTAs are persistent over their tiles.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace stcd {

constexpr int kTileH = 16;
constexpr int kTileW = 8;
constexpr int kMaxSrc = 6;
constexpr int kMaxPhase = 4;
constexpr int kMaxAStages = 8;
constexpr int kMaxWStages = 16;
constexpr int kMaxChunks = 128;  // per phase
constexpr int kMaxTaps = 512;    // per phase
constexpr int kConvThreads = 224;

struct Chunk {
  int16_t src, c0;
  int16_t by, bx;
  int32_t n_off;
  int16_t tap_begin, n_taps;
};
struct Tap {
  int16_t ty, tx;
};
struct PhaseInfo {
  int32_t chunk_begin, chunk_count, oy, ox, w_block, n_blocks;
};

struct TmapPack {
  CUtensorMap src[kMaxSrc];
};

struct ConvParams {
  int32_t hg, wg, tiles_x, tiles_y, n_img, pair_off, n_tiles;  // n_tiles = tiles_x*tiles_y*n_img
  int32_t n_ntiles;                                            // N tiles (grid.y = n_phase * n_ntiles)
  int32_t src_sy[kMaxSrc], src_sx[kMaxSrc], src_pw[kMaxSrc], src_ph[kMaxSrc];
  int32_t osy, osx, ho, wo;
  int32_t n_phase;
  PhaseInfo phase[kMaxPhase];
  const Chunk* chunks;
  const Tap* taps;
  int32_t kc, n_tile, cout, mt;
  int32_t a_stages, w_stages, w_resident;
  uint32_t a_stage_bytes, a_sub_bytes, wblk_bytes, tmem_cols, acc_cols;
  const uint8_t* wpack;        // device: [n_ntiles][total blocks] blocks of wblk_bytes
  int32_t blocks_per_ntile;    // total blocks over all phases
  const float* scale;
  const float* shift;
  const float* scale2;
  const float* shift2;
  int32_t relu;
  const __nv_bfloat16* res;
  int32_t res_c8;
  __nv_bfloat16* out0;
  int32_t out0_c8, out0_coff;
  __nv_bfloat16* out_raw;
  int32_t out_raw_c8;
  __nv_bfloat16* out_pool;
  int32_t out_pool_c8;
  __nv_bfloat16* out_diff;
  int32_t out_diff_c8;
  float* out_f32;
  int32_t n_valid;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void unpack8_bf16(uint4 q, float* v) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    v[2 * j] = __low2float(h);
    v[2 * j + 1] = __high2float(h);
  }
}
// element offset of (img, channel-chunk c8, y, x) in a [img][C8][H][W][8] tensor
__device__ __forceinline__ size_t nc8_off(int n, int c8, int C8, int H, int W, int y, int x) {
  return (((static_cast<size_t>(n) * C8 + c8) * H + y) * W + x) * 8;
}

// Un-swizzled K-major operand descriptor: start, LBO (K direction), SBO (M/N direction), bytes.
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  return d;                              // layout type 0 = no swizzle
}

__global__ void __launch_bounds__(kConvThreads, 2) conv_ws_kernel(const __grid_constant__ TmapPack tm,
                                                                  const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  __shared__ __align__(8) uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ Chunk s_chunks[kMaxChunks];
  __shared__ Tap s_taps[kMaxTaps];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int ph = blockIdx.y / p.n_ntiles;
  const int nt = blockIdx.y - ph * p.n_ntiles;
  const int n0 = nt * p.n_tile;
  const PhaseInfo phase = p.phase[ph];
  const int my_tiles = (p.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* smem_w = smem;  // W region first, then the A ring
  const uint32_t w_region = (p.w_resident ? static_cast<uint32_t>(phase.n_blocks) : static_cast<uint32_t>(p.w_stages)) * p.wblk_bytes;
  uint8_t* smem_a = smem + ((w_region + 127u) & ~127u);
  const uint8_t* wsrc = p.wpack + (static_cast<size_t>(nt) * p.blocks_per_ntile + phase.w_block) * p.wblk_bytes;

  for (int i = threadIdx.x; i < phase.chunk_count; i += blockDim.x) s_chunks[i] = p.chunks[phase.chunk_begin + i];
  {
    const int tap0 = p.chunks[phase.chunk_begin].tap_begin;
    for (int i = threadIdx.x; i < phase.n_blocks; i += blockDim.x) s_taps[i] = p.taps[tap0 + i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < kMaxWStages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int tap_base = s_chunks[0].tap_begin;  // taps of this phase are contiguous from here

  if (warp == 0) {
    // ============================== A producer (TMA) ==============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = 0; t < my_tiles; ++t) {
        int tile = blockIdx.x + t * gridDim.x;
        const int tile_x = tile % p.tiles_x;
        tile /= p.tiles_x;
        const int tile_y = tile % p.tiles_y;
        const int img = tile / p.tiles_y;
        const int x0 = tile_x * kTileW, y0 = tile_y * kTileH;
        for (int c = 0; c < phase.chunk_count; ++c, ++it) {
          const Chunk ch = s_chunks[c];
          const int s = it % p.a_stages;
          const uint32_t par = (it / p.a_stages) & 1;
          mbar_wait(&a_empty[s], par ^ 1);
          const uint32_t box_bytes = static_cast<uint32_t>(p.kc / 8) * p.src_ph[ch.src] * p.src_pw[ch.src] * 16u;
          mbar_expect_tx(&a_full[s], p.mt * box_bytes);
          uint8_t* dst = smem_a + s * p.a_stage_bytes;
          for (int m = 0; m < p.mt; ++m)
            tma_load_5d(dst + m * p.a_sub_bytes, &tm.src[ch.src], &a_full[s], 0, x0 * p.src_sx[ch.src] + ch.bx,
                        y0 * p.src_sy[ch.src] + ch.by, ch.c0 >> 3, img + ch.n_off + m * p.pair_off);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== W producer (bulk copies) ==============================
    if (lane == 0) {
      if (p.w_resident) {
        // everything once: barrier slot s covers blocks [s*per, (s+1)*per) so MMAs can start early
        const int per = (phase.n_blocks + kMaxWStages - 1) / kMaxWStages;
        for (int s = 0; s < kMaxWStages; ++s) {
          const int b0 = s * per, b1 = min(phase.n_blocks, b0 + per);
          if (b0 >= b1) break;
          mbar_expect_tx(&w_full[s], static_cast<uint32_t>(b1 - b0) * p.wblk_bytes);
          for (int b = b0; b < b1; ++b)
            bulk_load(smem_w + static_cast<size_t>(b) * p.wblk_bytes, wsrc + static_cast<size_t>(b) * p.wblk_bytes,
                      p.wblk_bytes, &w_full[s]);
        }
      } else {
        uint32_t it = 0;
        for (int t = 0; t < my_tiles; ++t) {
          for (int b = 0; b < phase.n_blocks; ++b, ++it) {
            const int s = it % p.w_stages;
            const uint32_t par = (it / p.w_stages) & 1;
            mbar_wait(&w_empty[s], par ^ 1);
            mbar_expect_tx(&w_full[s], p.wblk_bytes);
            bulk_load(smem_w + static_cast<size_t>(s) * p.wblk_bytes, wsrc + static_cast<size_t>(b) * p.wblk_bytes,
                      p.wblk_bytes, &w_full[s]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(p.n_tile);
      const int ksteps = p.kc / 16;
      const uint32_t w_lbo = static_cast<uint32_t>(p.n_tile) * 16u, w_sbo = 128u;
      const int per = (phase.n_blocks + kMaxWStages - 1) / kMaxWStages;
      uint32_t a_it = 0, w_it = 0;
      int w_ready = 0;  // resident mode: barrier slots already waited for
      for (int t = 0; t < my_tiles; ++t) {
        const int acc = t & 1;
        mbar_wait(&acc_empty[acc], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * p.acc_cols;
        uint32_t first = 1;
        int blk = 0;
        for (int c = 0; c < phase.chunk_count; ++c, ++a_it) {
          const Chunk ch = s_chunks[c];
          const int s = a_it % p.a_stages;
          mbar_wait(&a_full[s], (a_it / p.a_stages) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem_a + s * p.a_stage_bytes);
          const uint32_t pw = p.src_pw[ch.src];
          const uint32_t a_sbo = pw * 16u, a_lbo = pw * p.src_ph[ch.src] * 16u;
          for (int k = 0; k < ch.n_taps; ++k, ++blk) {
            const Tap tp = s_taps[ch.tap_begin - tap_base + k];
            uint32_t b_base;
            int ws = 0;
            if (p.w_resident) {
              const int need = blk / per;
              while (w_ready <= need) {
                mbar_wait(&w_full[w_ready], 0);
                ++w_ready;
              }
              tc_fence_after();
              b_base = smem_u32(smem_w + static_cast<size_t>(blk) * p.wblk_bytes);
            } else {
              ws = w_it % p.w_stages;
              mbar_wait(&w_full[ws], (w_it / p.w_stages) & 1);
              tc_fence_after();
              b_base = smem_u32(smem_w + static_cast<size_t>(ws) * p.wblk_bytes);
              ++w_it;
            }
            const uint32_t a_tap = a_base + (static_cast<uint32_t>(tp.ty) * pw + tp.tx) * 16u;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t bdesc = make_desc_nosw(b_base + ks * 2 * w_lbo, w_lbo, w_sbo);
              for (int m = 0; m < p.mt; ++m) {
                const uint32_t a_addr = a_tap + m * p.a_sub_bytes + ks * 2 * a_lbo;
                const uint64_t adesc = make_desc_nosw(a_addr, a_lbo, a_sbo);
                umma_bf16(d_tmem + m * p.n_tile, adesc, bdesc, idesc, first ? 0u : 1u);
              }
              first = 0;
            }
            if (!p.w_resident) umma_commit(&w_empty[ws]);
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue (warps 3..6) ==============================
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int ty = 4 * wq + (lane >> 3);
    const int tx = lane & 7;
    const size_t img_pix = static_cast<size_t>(p.ho) * p.wo;
    for (int t = 0; t < my_tiles; ++t) {
      int tile = blockIdx.x + t * gridDim.x;
      const int tile_x = tile % p.tiles_x;
      tile /= p.tiles_x;
      const int tile_y = tile % p.tiles_y;
      const int img = tile / p.tiles_y;
      const int gy = tile_y * kTileH + ty, gx = tile_x * kTileW + tx;
      const bool valid = (gy < p.hg) && (gx < p.wg);
      const int oy = gy * p.osy + phase.oy, ox = gx * p.osx + phase.ox;
      const int acc = t & 1;
      mbar_wait(&acc_full[acc], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t tlane = tmem_base + acc * p.acc_cols + (static_cast<uint32_t>(wq * 32) << 16);

      for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        const int ch = n0 + c0;
        if (ch >= p.cout) break;
        const int nv = min(16, p.cout - ch);  // valid channels in this 16-group (bf16 outputs: 8 or 16)
        float v[2][16];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          if (m >= p.mt) break;
          uint32_t raw[16];
          tmem_ld16(tlane + m * p.n_tile + c0, raw);
          tmem_wait_ld();
          const int n = img + m * p.pair_off;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[m][j] = fmaf(__uint_as_float(raw[j]), __ldg(p.scale + ch + j), __ldg(p.shift + ch + j));
          if (p.out_raw != nullptr && valid) {
            __nv_bfloat16* o = p.out_raw + nc8_off(n, ch >> 3, p.out_raw_c8, p.ho, p.wo, oy, ox);
            *reinterpret_cast<uint4*>(o) = pack8_bf16(v[m]);
            if (nv > 8) *reinterpret_cast<uint4*>(o + img_pix * 8) = pack8_bf16(v[m] + 8);
          }
          if (p.scale2 != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[m][j] = fmaf(v[m][j], __ldg(p.scale2 + ch + j), __ldg(p.shift2 + ch + j));
          }
          if (p.res != nullptr && valid) {
            const __nv_bfloat16* r = p.res + nc8_off(n, ch >> 3, p.res_c8, p.ho, p.wo, oy, ox);
            float rv[16];
            unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(r)), rv);
            if (nv > 8) unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(r + img_pix * 8)), rv + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[m][j] += (j < nv) ? rv[j] : 0.f;
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[m][j] = fmaxf(v[m][j], 0.f);
          }
          if (p.out0 != nullptr && valid) {
            __nv_bfloat16* o = p.out0 + nc8_off(n, (p.out0_coff + ch) >> 3, p.out0_c8, p.ho, p.wo, oy, ox);
            *reinterpret_cast<uint4*>(o) = pack8_bf16(v[m]);
            if (nv > 8) *reinterpret_cast<uint4*>(o + img_pix * 8) = pack8_bf16(v[m] + 8);
          }
          if (p.out_f32 != nullptr && valid && n < p.n_valid) {
            for (int j = 0; j < nv; ++j)
              p.out_f32[(static_cast<size_t>(n) * p.cout + ch + j) * img_pix + static_cast<size_t>(oy) * p.wo + ox] = v[m][j];
          }
          if (p.out_pool != nullptr) {
            float q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = fmaxf(v[m][j], __shfl_xor_sync(0xffffffffu, v[m][j], 1));
              q[j] = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 8));
            }
            if (valid && ((lane & 9) == 0)) {
              __nv_bfloat16* o = p.out_pool + nc8_off(n, ch >> 3, p.out_pool_c8, p.ho >> 1, p.wo >> 1, oy >> 1, ox >> 1);
              *reinterpret_cast<uint4*>(o) = pack8_bf16(q);
              if (nv > 8) *reinterpret_cast<uint4*>(o + (img_pix >> 2) * 8) = pack8_bf16(q + 8);
            }
          }
        }
        if (p.mt == 2 && p.out_diff != nullptr && valid) {
          float d[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) d[j] = fabsf(v[0][j] - v[1][j]);
          __nv_bfloat16* o = p.out_diff + nc8_off(img, ch >> 3, p.out_diff_c8, p.ho, p.wo, oy, ox);
          *reinterpret_cast<uint4*>(o) = pack8_bf16(d);
          if (nv > 8) *reinterpret_cast<uint4*>(o + img_pix * 8) = pack8_bf16(d + 8);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace stcd
