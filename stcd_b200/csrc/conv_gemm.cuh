// K1: implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels x N channels] (TMEM, fp32) = sum over K-blocks  A[128 x kc] * W[N x kc]^T
//
// One CTA computes one 8x16-pixel output tile (M = 128) of one image (or of a Siamese image
// pair sharing the weight tiles) for one output phase.  A K-block is `kc` channels of one
// source tensor at one filter tap: a 4-D TMA box {kc, 16, 8, 1} over the NHWC tensor whose
// out-of-bounds pixels are zero-filled by the TMA unit (= the conv's zero padding), landing in
// shared memory already in the swizzled K-major layout tcgen05.mma consumes.  The list of
// K-blocks (the "K-program") is data, so the same kernel executes 3x3/1x1 convs, strided convs,
// ConvTranspose2d phases and virtual channel-concats (torch.cat) of up to 6 sources.
//
// Warp roles: warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer, then all four warps
// run the epilogue (TMEM -> registers -> folded BN / ReLU / residual / |f1-f2| / 2x2 max-pool
// -> global).  Several CTAs are resident per SM (small tiles), so one CTA's epilogue overlaps
// another's main loop.
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace stcd {

constexpr int kTileH = 8;
constexpr int kTileW = 16;
constexpr int kMaxSrc = 6;
constexpr int kMaxPhase = 4;
constexpr int kMaxStages = 12;
constexpr int kMaxKProg = 256;  // K-blocks per phase

struct KEntry {
  int16_t src, dy, dx, c0;
  int32_t n_off;
  int32_t wk;
};
struct PhaseInfo {
  int32_t k_begin, k_count, oy, ox, w_row;
};

struct TmapPack {
  CUtensorMap src[kMaxSrc];
  CUtensorMap w;
};

struct ConvParams {
  int32_t hg, wg, tiles_x, tiles_y, n_img, pair_off;
  int32_t src_sy[kMaxSrc], src_sx[kMaxSrc];
  int32_t osy, osx, ho, wo;
  int32_t n_phase;
  PhaseInfo phase[kMaxPhase];
  const KEntry* kprog;
  int32_t kc, n_tile, cout;
  int32_t stages, group;
  uint32_t a_bytes, b_bytes, sub_bytes, tmem_cols;  // b_bytes = bytes TMA delivers; sub_bytes is 1024-aligned
  const float* scale;
  const float* shift;
  const float* scale2;
  const float* shift2;
  int32_t relu;
  const __nv_bfloat16* res;
  int32_t res_c;
  __nv_bfloat16* out0;
  int32_t out0_c, out0_coff;
  __nv_bfloat16* out_raw;
  int32_t out_raw_c;
  __nv_bfloat16* out_pool;
  int32_t out_pool_c;
  __nv_bfloat16* out_diff;
  int32_t out_diff_c;
  float* out_f32;
  int32_t n_valid;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, const float* v, int nv) {
  // nv in {8, 16}: number of valid channels of this 16-chunk
  uint4 lo = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  *reinterpret_cast<uint4*>(dst) = lo;
  if (nv > 8) {
    uint4 hi =
        make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
    *reinterpret_cast<uint4*>(dst + 8) = hi;
  }
}
__device__ __forceinline__ void load16_bf16(const __nv_bfloat16* src, float* v, int nv) {
  uint4 lo = __ldg(reinterpret_cast<const uint4*>(src));
  uint4 hi = make_uint4(0, 0, 0, 0);
  if (nv > 8) hi = __ldg(reinterpret_cast<const uint4*>(src + 8));
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    v[2 * j] = __low2float(h);
    v[2 * j + 1] = __high2float(h);
  }
}

template <int MT>
__global__ void __launch_bounds__(128) conv_gemm_kernel(const __grid_constant__ TmapPack tm,
                                                        const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ KEntry kp[kMaxKProg];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile decode
  int t = blockIdx.x;
  const int tile_x = t % p.tiles_x;
  t /= p.tiles_x;
  const int tile_y = t % p.tiles_y;
  t /= p.tiles_y;
  const int img = t % p.n_img;
  const int ph = t / p.n_img;
  const int n0 = blockIdx.y * p.n_tile;
  const int x0 = tile_x * kTileW;
  const int y0 = tile_y * kTileH;
  const PhaseInfo phase = p.phase[ph];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = p.sub_bytes * p.group;
  const int n_groups = (phase.k_count + p.group - 1) / p.group;

  for (int i = threadIdx.x; i < phase.k_count; i += blockDim.x) kp[i] = p.kprog[phase.k_begin + i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int g = 0; g < n_groups; ++g) {
        const int s = g % p.stages;
        const uint32_t par = (g / p.stages) & 1;
        mbar_wait(&empty_bar[s], par ^ 1);
        const int k_first = g * p.group;
        const int nsub = min(p.group, phase.k_count - k_first);
        mbar_expect_tx(&full_bar[s], nsub * (MT * p.a_bytes + p.b_bytes));
        uint8_t* sbase = smem + s * stage_bytes;
        for (int j = 0; j < nsub; ++j) {
          const KEntry e = kp[k_first + j];
          uint8_t* sub = sbase + j * p.sub_bytes;
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            tma_load_4d(sub + m * p.a_bytes, &tm.src[e.src], &full_bar[s], e.c0, x0 * p.src_sx[e.src] + e.dx,
                        y0 * p.src_sy[e.src] + e.dy, img + e.n_off + m * p.pair_off);
          }
          tma_load_2d(sub + MT * p.a_bytes, &tm.w, &full_bar[s], e.wk, phase.w_row + n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc_bf16(p.n_tile);
      const uint32_t row_bytes = p.kc * 2;
      const uint32_t sbo = 8 * row_bytes;
      const int ksteps = p.kc / 16;
      uint32_t first = 1;
      for (int g = 0; g < n_groups; ++g) {
        const int s = g % p.stages;
        const uint32_t par = (g / p.stages) & 1;
        mbar_wait(&full_bar[s], par);
        tc_fence_after();
        const int nsub = min(p.group, phase.k_count - g * p.group);
        const uint32_t sbase = smem_u32(smem + s * stage_bytes);
        for (int j = 0; j < nsub; ++j) {
          const uint32_t sub = sbase + j * p.sub_bytes;
          const uint64_t bdesc = make_kmajor_desc(sub + MT * p.a_bytes, row_bytes, sbo);
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const uint64_t adesc = make_kmajor_desc(sub + m * p.a_bytes, row_bytes, sbo);
            for (int k = 0; k < ksteps; ++k) {
              // advancing K by 16 bf16 = 32 B = 2 units of the 16-byte start-address field
              umma_bf16(tmem_base + m * p.n_tile, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
            }
          }
          first = 0;
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
    __syncwarp();
  }

  // ===== epilogue (all 4 warps; warp w owns TMEM lanes [32w, 32w+32)) =====
  mbar_wait(&accum_bar, 0);
  tc_fence_after();

  const int ty = 2 * warp + (lane >> 4);
  const int tx = lane & 15;
  const int gy = y0 + ty, gx = x0 + tx;
  const bool valid = (gy < p.hg) && (gx < p.wg);
  const int oy = gy * p.osy + phase.oy, ox = gx * p.osx + phase.ox;
  const size_t opix = (static_cast<size_t>(oy) * p.wo + ox);
  const size_t img_pix = static_cast<size_t>(p.ho) * p.wo;
  const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
    const int ch = n0 + c0;
    if (ch >= p.cout) break;
    const int nv = min(16, p.cout - ch);
    uint32_t raw[MT][16];
#pragma unroll
    for (int m = 0; m < MT; ++m) tmem_ld16(tlane + m * p.n_tile + c0, raw[m]);
    tmem_wait_ld();

    float v[MT][16];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int n = img + m * p.pair_off;
      const size_t pix = static_cast<size_t>(n) * img_pix + opix;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[m][j] = fmaf(__uint_as_float(raw[m][j]), __ldg(p.scale + ch + j), __ldg(p.shift + ch + j));
      if (p.out_raw != nullptr && valid) store16_bf16(p.out_raw + pix * p.out_raw_c + ch, v[m], nv);
      if (p.scale2 != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[m][j] = fmaf(v[m][j], __ldg(p.scale2 + ch + j), __ldg(p.shift2 + ch + j));
      }
      if (p.res != nullptr && valid) {
        float r[16];
        load16_bf16(p.res + pix * p.res_c + ch, r, nv);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[m][j] += r[j];
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[m][j] = fmaxf(v[m][j], 0.f);
      }
      if (p.out0 != nullptr && valid) store16_bf16(p.out0 + pix * p.out0_c + p.out0_coff + ch, v[m], nv);
      if (p.out_f32 != nullptr && valid && n < p.n_valid) {
        for (int j = 0; j < nv; ++j) p.out_f32[(static_cast<size_t>(n) * p.cout + ch + j) * img_pix + opix] = v[m][j];
      }
      if (p.out_pool != nullptr) {
        float q[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = fmaxf(v[m][j], __shfl_xor_sync(0xffffffffu, v[m][j], 1));
          q[j] = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, 16));
        }
        if (valid && ((lane & 17) == 0)) {
          const size_t ppix = (static_cast<size_t>(n) * (p.ho >> 1) + (oy >> 1)) * (p.wo >> 1) + (ox >> 1);
          store16_bf16(p.out_pool + ppix * p.out_pool_c + ch, q, nv);
        }
      }
    }
    if (MT == 2 && p.out_diff != nullptr && valid) {
      float d[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) d[j] = fabsf(v[0][j] - v[MT - 1][j]);
      store16_bf16(p.out_diff + (static_cast<size_t>(img) * img_pix + opix) * p.out_diff_c + ch, d, nv);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace stcd
