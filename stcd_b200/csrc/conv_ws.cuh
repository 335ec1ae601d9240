// K1: persistent, warp-specialised implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels x N channels] (TMEM, fp32) += A[128 x 16] * W[N x 16]^T   per filter tap and 16 channels
//
// Data layout.  Activations live in HBM as bf16 [img][c/8][h][w][8].  One TMA box
// {(8 ch, box_w) merged, box_h, kc/8, 1 img} therefore lands in shared memory as
// [kc/8][box_h][box_w][8 ch]: the un-swizzled K-major "core matrix" layout of tcgen05.mma
// (8 rows x 16 bytes contiguous), with
//   SBO (next 8 tile rows = next tile line)  = box_w * 16 bytes
//   LBO (next 8 channels)                    = box_h * box_w * 16 bytes
// A CTA tile is 16 x 8 output pixels (M = 128; 8 consecutive pixels of a line form a core matrix),
// so the A operand of filter tap (ty, tx) is the SAME box at byte offset (ty*box_w + tx)*16: a 3x3
// conv fetches its 18x10 halo once per K-chunk and issues 9 MMAs from it (L2->SM traffic 1.4x
// instead of 9x).  Out-of-bounds box pixels are zero-filled by TMA (= the conv's zero padding).
// Strided convs use element strides in a 5-D tensor map and one box per tap.
//
// Weights are host-packed per (N tile, phase) in consumption order as [block][kc/8][n_tile][8]
// (block = one tap of one chunk; SBO = 128 B, LBO = n_tile*16 B) and arrive by 1-D bulk copies:
// either ALL blocks once per CTA (weight-stationary: the CTA then streams pixel tiles past them)
// or through a ring when they do not fit.
//
// Warp roles (8 warps): 0 = A producer (TMA), 1 = W producer (bulk copy), 2 = MMA issuer (also
// owns TMEM alloc/dealloc), 3..6 = epilogue (TMEM -> registers -> folded BN / ReLU / residual /
// |f1-f2| / 2x2 max-pool -> global), 7 = residual producer (TMA into a small ring, when the op has one).  The NE = 8 instances
// (one CTA per SM: short-K residual layers, horizontally folded layers) have eight epilogue warps 3..10 -- two per TMEM lane
// quarter, alternating 16-column steps -- and the residual producer is warp 11.  Accumulators are double-buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1; CTAs are persistent over their tiles.  How many M sub-tiles (images) share a
// weight block per pass, the epilogue width and the folded-layer issue loop are chosen per op by the plan, which times the
// admissible variants on its own workspace (stcd_plan_finalize).
//
// Measured on B200 (tools/ubench, profiles/): a tcgen05.mma M=128 K=16 in SS mode costs ~45
// cycles for N <= 64 (the 4 KB A read from shared memory), 64 for N=128 — and a single warp
// retires one dependent SASS instruction every ~10 cycles, so the producer / issuer loops are
// table-driven (everything tile-invariant is precomputed into shared memory at CTA start), run
// by the whole warp in uniform control flow (descriptors stay in uniform registers) and only
// the tcgen05 / TMA instructions themselves are issued by one elected lane.
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
constexpr int kMaxChunks = 128;  // A-stage loads per phase
constexpr int kMaxTaps = 512;    // weight blocks per phase (split precision triples the K-program: 1024 channels x 9 taps / 64 x 3 = 432)
constexpr int kConvThreads = 256;   // 8 warps: A producer, W producer, MMA issuer, 4 epilogue warps, residual producer
constexpr int kConvThreads8 = 384;  // 12 warps: the same with 8 epilogue warps (one CTA per SM)
constexpr int kFastMma = 36;     // per-chunk MMA offsets kept in the constant bank (9 taps x 4 K steps)
constexpr int kMaxRSlots = 8;    // residual blocks in flight (TMA -> shared-memory ring)
// Horizontally folded 3x3 convs (E_XF): the tile is 8 rows x 16 columns of INPUT positions, of which the inner 14 columns
// are outputs (tiles overlap by one column on each side).
constexpr int kXfTileH = 8;
constexpr int kXfTileW = 16;
constexpr int kXfStep = 14;

// epilogue features; a kernel instance either fixes them at compile time or (E_GENERIC) reads
// them from ConvParams at run time
enum : uint32_t {
  E_RAW = 1u,     // out_raw <- acc*scale + shift
  E_AFF2 = 2u,    // v = v*scale2 + shift2
  E_RES = 4u,     // v += res
  E_RELU = 8u,    // v = max(v, 0)
  E_OUT0 = 16u,   // out0 <- v
  E_POOL = 32u,   // out_pool <- maxpool2x2(v)
  E_DIFF = 64u,   // out_diff <- |v(T1) - v(T2)|
  E_F32 = 128u,   // out_f32 (NCHW fp32, external) <- v
  E_ACTX = 256u,  // extended activation: GELU / PReLU and/or activation before the second affine (read at run time)
  E_RSM = 512u,   // with E_RES: the residual tile arrives by TMA in a shared-memory ring (ConvParams::res_slots > 0)
  // Horizontal tap folding of a stride-1 3x3 conv with few output channels.  An SS-mode tcgen05.mma M=128 K=16 costs ~45
  // cycles for any N <= 64 (the 4 KB A operand read from shared memory): a Cout-32 layer run tap by tap is capped at 36 % of
  // the tensor peak.  Folded, the three taps of one filter ROW share one MMA: A = the input tile shifted vertically only,
  // B = [W(dy,-1); W(dy,0); W(dy,+1)] (N = 3 * cs, cs = Cout rounded up to 16), so column block b of the accumulator at input
  // position x holds that tap's contribution to output x - (b - 1), and the epilogue forms
  //     out[x] = D_0[x - 1] + D_1[x] + D_2[x + 1]
  // with two warp shuffles per value (a warp owns two tile rows of 16 positions; the outer two columns of a tile only
  // feed their neighbours).  3 MMAs of N = 96 (~56 cycles) replace 9 of N = 32 (45 cycles): 2.4x fewer tensor-pipe cycles
  // for Cout 32, 3x for Cout 16, and the box has no horizontal halo (rows of 256 contiguous bytes).
  E_XF = 1024u,
  E_GENERIC = 1u << 31
};

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
  int32_t tap0;      // first entry of this phase in ConvParams::taps (= chunks[chunk_begin].tap_begin)
};

struct TmapPack {
  CUtensorMap src[kMaxSrc];
  CUtensorMap res;   // residual tensor, box = one output tile x n_tile channels (E_RSM)
};

struct ConvParams {
  int32_t hg, wg, tiles_x, tiles_y, n_img, n_tiles;  // n_img = base images; n_tiles = tiles_x*tiles_y*n_img
  int32_t m_off[4];  // image offset of sub-tile m (M tiles a CTA accumulates side by side): T2 partner and/or other images
  int32_t n_ntiles;                                            // N tiles (grid.y = n_phase * n_ntiles)
  int32_t src_sy[kMaxSrc], src_sx[kMaxSrc], src_pw[kMaxSrc], src_ph[kMaxSrc];
  int32_t src_merged[kMaxSrc];  // 1: 4-D map with (8 ch, x) merged into one dimension (stride-1 sources)
  int32_t osy, osx, ho, wo;
  int32_t n_phase, n_src;
  PhaseInfo phase[kMaxPhase];
  const Chunk* chunks;
  const Tap* taps;
  int32_t kc, n_tile, cout, mt;
  int32_t a_stages, w_stages, w_resident;
  uint32_t a_stage_bytes, a_sub_bytes, wblk_bytes, tmem_cols, acc_cols;
  const uint8_t* wpack;        // device: [n_ntiles][total blocks] blocks of wblk_bytes
  int32_t blocks_per_ntile;    // total blocks over all phases
  uint32_t tab_bytes;          // dynamic shared memory reserved for the per-MMA descriptor table
  const float* scale;
  const float* shift;
  const float* scale2;
  const float* shift2;
  int32_t relu;        // activation kind: 0 none, 1 ReLU, 2 GELU (erf), 3 PReLU (one slope act_alpha)
  int32_t act_pre;     // 1: the activation sits before the second affine (conv -> act -> BN) instead of last
  float act_alpha;
  const __nv_bfloat16* res;
  int32_t res_c8;
  // Residual ring (E_RSM).  Per-thread residual loads kept only ~2 KB per epilogue warp in flight: at DRAM latency that is
  // 2.4 TB/s for the whole chip, and the short-K layers with a residual (conv*.conv2 of the nested blocks) ran at exactly
  // that rate (262 us with the residual, 152 us without).  A dedicated warp now fetches the residual by TMA into a ring of
  // small slots -- one slot = `res_rb` channels of the MS sub-tiles the epilogue finishes together (<= 16 KB, so the ring also
  // fits beside 128-column Siamese-pair ops whose whole residual tile would take 64 KB) -- in the order the epilogue consumes
  // them, up to `res_slots` blocks ahead; the epilogue reads it with conflict-free 16-byte LDS.
  int32_t res_slots;                       // 0: per-thread global loads
  uint32_t res_slot_bytes, res_sub_bytes;  // ring slot = MS sub-tiles of [res_rb/8][tile rows][tile px][8] bf16
  int32_t res_ch;                          // residual channels this CTA needs per sub-tile (n_tile; cs for E_XF)
  int32_t res_rb;                          // channels per slot (multiple of 16)
  int32_t xf_cs;                           // E_XF: column-block stride cs (n_tile = 3 * cs); 0 otherwise
  // Split precision (the "tf32" tolerance class; generic instances only).  Every bf16 tensor holds a hi plane (channel groups
  // [0, c8/2)) and a lo plane ([c8/2, c8)): hi = bf16(v), lo = bf16(v - hi).  The K-program already reads (hi, lo, hi) against
  // (Whi, Whi, Wlo); the epilogue writes both planes of every bf16 output and reads the residual as hi + lo.
  int32_t split;
  __nv_bfloat16* out0;
  int32_t out0_c8, out0_coff;
  int32_t reverse;   // 1: walk the tiles last-to-first (consecutive layers alternate, so a layer starts on the lines its producer wrote last: L2 hits)
  int32_t out0_s2d;  // 1: out0 is stored space-to-depth ([ho/2][wo/2] pixels, 4*cout channels)
  int32_t fold_cs, fold_cout;  // > 0: output phases folded into N (column p*fold_cs + c -> phase (oy + p / osx, ox + p % osx), channel c)
  __nv_bfloat16* out_raw;
  int32_t out_raw_c8;
  __nv_bfloat16* out_pool;
  int32_t out_pool_c8;
  __nv_bfloat16* out_diff;
  int32_t out_diff_c8;
  float* out_f32;
  int32_t n_valid;
  // Fast issue path for "regular" phases (every chunk has the same tap list and box geometry, e.g.
  // any stride-1 3x3 conv): the per-MMA operand offsets live here, in the constant bank, so the
  // issue loop reads them with uniform loads and never leaves the uniform datapath.
  int32_t f_regular[kMaxPhase];      // 1: use f_off
  int32_t f_nmma[kMaxPhase];         // MMAs per chunk and M tile = taps * K steps (<= kFastMma)
  uint32_t f_a_hi[kMaxPhase];        // A descriptor high word (SBO | version)
  uint32_t f_a_lo_lbo[kMaxPhase];    // A descriptor LBO field
  uint32_t f_b_chunk16[kMaxPhase];   // weight bytes per chunk >> 4
  alignas(16) uint32_t f_off[kMaxPhase][36];  // A offset (16 B units) | B offset inside the chunk's weights << 16
  int32_t xf_fast;   // E_XF: 1 = this op's issuer runs the register-resident chunk loop (chosen per op by the plan)
  int32_t dbg;       // diagnostics (STCD_DBG): bit0 skip MMAs
  long long* trace;  // diagnostics (STCD_TRACE=1): 16 clock stamps per CTA, else nullptr
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
// lo part of the split representation of 8 values: bf16(v - float(bf16(v)))
__device__ __forceinline__ uint4 pack8_bf16_lo(const float* v) {
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
  return pack8_bf16(r);
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
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
// 2x2 max over the window whose top-left pixel this lane holds: the right neighbour is DX lanes away, the lower one DY
template <int DX, int DY>
__device__ __forceinline__ uint32_t pool1_bf16x2(uint32_t a) {
  a = max_bf16x2(a, __shfl_down_sync(0xffffffffu, a, DX));
  return max_bf16x2(a, __shfl_down_sync(0xffffffffu, a, DY));
}
template <int DX, int DY>
__device__ __forceinline__ uint4 pool4_bf16x2(uint4 q) {
  return make_uint4(pool1_bf16x2<DX, DY>(q.x), pool1_bf16x2<DX, DY>(q.y), pool1_bf16x2<DX, DY>(q.z), pool1_bf16x2<DX, DY>(q.w));
}

// per-chunk tables, built once per CTA
struct ChunkLoad {   // A producer
  int32_t xm, ym;    // tile origin multipliers: coordinate = x0 * xm + xa, y0 * ym + ya
  int32_t xa, ya;
  int32_t c8, n_off;
  uint32_t tx_bytes; // bytes of all MT boxes
  int32_t src_merged;  // src | merged << 8
};
struct ChunkMma {    // MMA issuer
  uint32_t a_hi;       // A descriptor high word: SBO | version
  uint32_t n_mma;      // MMAs (per M tile) this chunk feeds = taps * K steps
};

template <int N>
struct IntC {
  static constexpr int value = N;
};

#define STCD_HAS(flag, runtime_expr) ((EPI & E_GENERIC) ? (runtime_expr) : ((EPI & (flag)) != 0))

// MT = M tiles (images) per CTA pass sharing every weight block; MS = 2 when consecutive sub-tiles are
// (T1, T2) Siamese pairs (the |f1 - f2| epilogue needs both), else 1.
// NE = epilogue warps: 4 (one per TMEM lane quarter; two CTAs per SM) or 8 (two per quarter, alternating 16-column steps;
// 384 threads, one CTA per SM).  With one epilogue warp per scheduler every instruction of the step waits out its own latency
// (ncu: 8.9 cycles per issued instruction, the MMA warp idle 58 % of the time on the short-K residual layers); plans that run
// one CTA per SM anyway take the eight-warp instances where they exist.
template <int MT, int MS, uint32_t EPI, int NE = 4>
__global__ void __launch_bounds__(NE == 8 ? kConvThreads8 : kConvThreads, NE == 8 ? 1 : 2) conv_ws_kernel(const __grid_constant__ TmapPack tm,
                                                                                                           const __grid_constant__ ConvParams p) {
  static_assert(NE == 4 || NE == 8, "4 or 8 epilogue warps");
  static_assert(NE == 4 || ((EPI & E_RSM) != 0 || (EPI & E_RES) == 0), "eight epilogue warps: residual through the ring only");
  constexpr int kResWarp = 3 + NE;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxAStages], a_empty[kMaxAStages];
  __shared__ __align__(8) uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t r_full[kMaxRSlots], r_empty[kMaxRSlots];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) ChunkLoad s_cload[kMaxChunks];
  __shared__ __align__(8) ChunkMma s_cmma[kMaxChunks];
  __shared__ __align__(16) float s_aff[4][256];      // scale, shift, scale2, shift2 of this CTA's N tile
  __shared__ uint8_t s_blk_src[kMaxTaps];            // source index of each weight block's chunk (table set-up only)

  constexpr bool XF = (EPI & E_XF) != 0;
  constexpr int TH = XF ? kXfTileH : kTileH;      // tile rows
  constexpr int TW = XF ? kXfTileW : kTileW;      // tile columns (input positions for XF)
  constexpr int XSTEP = XF ? kXfStep : kTileW;    // output columns per tile
  constexpr int XOFF = XF ? 1 : 0;                // the tile starts one column left of its first output
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  long long* tr = p.trace ? p.trace + ((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 : nullptr;
  const long long clk0 = clock64();
#define STCD_STAMP(i) do { if (tr) tr[i] = clock64() - clk0; } while (0)
  // Trace builds (-DSTCD_TRACE_WAITS, tools/trace_build.sh): cycles a role spends inside each of its waits, summed over the CTA's
  // tiles (slots 3, 12..15 of the trace record).  Compiled out of the product library: the run-time form of these checks alone
  // cost SegCD 1-2 % (same-box A/B against the build without them, round 2).
#ifdef STCD_TRACE_WAITS
#define STCD_TWAIT(call, counter) do { if (tr) { const long long w0_ = clock64(); call; counter += clock64() - w0_; } else { call; } } while (0)
#else
#define STCD_TWAIT(call, counter) do { call; } while (0)
#endif
  if (tr && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tr[0] = static_cast<long long>(gt);
  }

  const int ph = blockIdx.z;   // output phase
  const int nt = blockIdx.y;   // N tile
  const int n0 = nt * p.n_tile;
  const PhaseInfo phase = p.phase[ph];
  const int my_tiles = (p.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  // per-MMA descriptor table {A descriptor low word relative to the stage, B descriptor low word},
  // one entry per (tap, K step) of this phase in issue order, then the W region, then the A ring
  uint2* s_mma = reinterpret_cast<uint2*>(smem);
  uint8_t* smem_w = smem + p.tab_bytes;
  const uint32_t w_region = (p.w_resident ? static_cast<uint32_t>(phase.n_blocks) : static_cast<uint32_t>(p.w_stages)) * p.wblk_bytes;
  uint8_t* smem_a = smem_w + ((w_region + 127u) & ~127u);
  uint8_t* smem_r = smem_a + static_cast<size_t>(p.a_stages) * p.a_stage_bytes;   // residual ring (a_stage_bytes is a multiple of 128)
  const uint8_t* wsrc = p.wpack + (static_cast<size_t>(nt) * p.blocks_per_ntile + phase.w_block) * p.wblk_bytes;
  const int w_per = (phase.n_blocks + kMaxWStages - 1) / kMaxWStages;  // resident mode: blocks per barrier slot

  // CTA set-up, spread over the warps so that none of it is serial (trace, round 2: 4 500-4 900 cycles = 2.4 us per CTA when one
  // thread initialised ~56 barriers after the table loop and the TMEM allocation came last; at 8 pairs every layer of the
  // FC-Siam nets is 10-28 us, so the set-up is on the critical path 25 times per step): the dependent grid is released first
  // (its own prologue does not read our outputs), warp 2 allocates TMEM, warp 3 initialises the barriers (<= 8 per lane), a
  // thread of warp 4 prefetches the tensor maps, and every thread builds its share of the tables below.
  pdl_launch_dependents();
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  } else if (warp == 3) {
    if (lane < p.a_stages) {
      mbar_init(&a_full[lane], 1);
      mbar_init(&a_empty[lane], 1);
    }
    if (lane < kMaxWStages) {
      mbar_init(&w_full[lane], 1);
      mbar_init(&w_empty[lane], 1);
    }
    if (lane < 2) {
      mbar_init(&acc_full[lane], 1);
      mbar_init(&acc_empty[lane], NE);
    }
    if (lane < kMaxRSlots) {
      mbar_init(&r_full[lane], 1);
      mbar_init(&r_empty[lane], NE);
    }
    fence_mbar_init();
  } else if (threadIdx.x == 128) {
    for (int i = 0; i < p.n_src; ++i) tma_prefetch_desc(&tm.src[i]);
    if (p.res_slots) tma_prefetch_desc(&tm.res);
  }
  // ---- tile-invariant tables.  Two passes so that no thread walks a list of dependent global loads (one thread per chunk
  // looping over its taps cost 4 000-5 800 cycles of set-up on the 3x3 layers, 2 200-2 900 on the single-tap ones: n_taps
  // serialised L2 round trips): every thread requests its chunk record AND its tap record(s) up front -- the phase's first
  // tap index comes with the kernel parameters -- pass 1 files the per-chunk tables and the block -> chunk map, pass 2 the
  // per-MMA descriptors, one thread per weight block.
  {
    const int tap0 = phase.tap0;
    const uint32_t w_base16 = (smem_u32(smem_w) & 0x3FFFF) >> 4;
    const int ksteps = p.kc >> 4;
    Tap my_tap[kMaxTaps / kConvThreads];
#pragma unroll
    for (int j = 0; j < kMaxTaps / kConvThreads; ++j) {
      const int blk = threadIdx.x + j * blockDim.x;
      my_tap[j] = blk < phase.n_blocks ? p.taps[tap0 + blk] : Tap{0, 0};
    }
    for (int c = threadIdx.x; c < phase.chunk_count; c += blockDim.x) {
      const Chunk ch = p.chunks[phase.chunk_begin + c];
      const int pw = p.src_pw[ch.src], phh = p.src_ph[ch.src];
      const int merged = p.src_merged[ch.src];
      ChunkLoad L;
      L.xm = merged ? 8 : p.src_sx[ch.src];
      L.ym = merged ? 1 : p.src_sy[ch.src];
      L.xa = merged ? ch.bx * 8 : ch.bx;
      L.ya = ch.by;
      L.c8 = ch.c0 >> 3;
      L.n_off = ch.n_off;
      L.tx_bytes = static_cast<uint32_t>(MT) * (p.kc / 8) * phh * pw * 16u;
      L.src_merged = ch.src | (merged << 8);
      s_cload[c] = L;
      ChunkMma M;
      // SBO (next 8 rows of M): the next tile line, pw * 16 B; XF tiles have no horizontal halo and 16-pixel lines, so the
      // 8-pixel groups of a tile are 128 B apart throughout
      M.a_hi = (XF ? 8u : (static_cast<uint32_t>(pw) & 0x3FFF)) | (1u << 14);   // descriptor version 1
      M.n_mma = static_cast<uint32_t>(ch.n_taps * ksteps);
      s_cmma[c] = M;
      for (int k = 0; k < ch.n_taps; ++k) s_blk_src[ch.tap_begin - tap0 + k] = static_cast<uint8_t>(ch.src);
    }
    __syncthreads();
    const uint32_t b_lo_lbo = (static_cast<uint32_t>(p.n_tile) & 0x3FFF) << 16;  // LBO = n_tile * 16 B
#pragma unroll
    for (int j = 0; j < kMaxTaps / kConvThreads; ++j) {
      const int blk = threadIdx.x + j * blockDim.x;
      if (blk < phase.n_blocks) {
        const int src = s_blk_src[blk];
        const int pw = p.src_pw[src], phh = p.src_ph[src];
        const uint32_t a_lo_lbo = (static_cast<uint32_t>(pw * phh) & 0x3FFF) << 16;   // LBO = pw * ph * 16 B
        const Tap tp = my_tap[j];
        for (int ks = 0; ks < ksteps; ++ks) {
          uint2 e;
          e.x = a_lo_lbo + static_cast<uint32_t>(tp.ty * pw + tp.tx) + static_cast<uint32_t>(ks) * 2u * (pw * phh);
          // resident weights: absolute block address; streamed weights: offset inside the ring slot
          e.y = b_lo_lbo + (p.w_resident ? w_base16 + static_cast<uint32_t>(blk) * (p.wblk_bytes >> 4) : 0u) +
                static_cast<uint32_t>(ks) * 2u * p.n_tile;
          s_mma[blk * ksteps + ks] = e;
        }
      }
    }
  }
  for (int i = threadIdx.x; i < p.n_tile; i += blockDim.x) {
    s_aff[0][i] = p.scale[n0 + i];
    s_aff[1][i] = p.shift[n0 + i];
    s_aff[2][i] = p.scale2 ? p.scale2[n0 + i] : 1.f;
    s_aff[3][i] = p.shift2 ? p.shift2[n0 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) STCD_STAMP(1);
  // Programmatic dependent launch: everything above (tables, barriers, TMEM) and the weight loads
  // below do not depend on the previous layer; only the activation loads wait for the previous kernel to finish.

  long long tw_a = 0, tw_acc = 0, tw_w = 0;     // trace mode (issuer warp): cycles waited for activations / a free accumulator / streamed weights
  // ---- Horizontally folded layers: the issuer's loop.
  // A chunk is one 16..64-channel block of one source = 3 filter rows x (1, 2 or 4) K steps.  ncu's source view
  // (profiles/r2_src_conv0_4_conv1.txt) showed the issuer as a pacing role -- never waiting for an operand, ~86 instructions per
  // 6-MMA chunk, two divergent regions and two constant-bank loads per chunk.  Here the chunk's operand offsets live in
  // registers for the whole kernel, the chunk is ONE divergent region and the weight base advances by an add: conv0_2..4.conv1
  // of SNUNet (4-6 chunks per tile) -5 .. -10 %, conv0_1.conv1 (3 chunks) +7 .. +13 % in the same A/B runs, so the plan switches
  // it on from 4 chunks per tile (STCD_XF_FAST_MIN).  (The same loop measured 11-17 % SLOWER than the table loop on the Siamese
  // 3x3 layers, so it stays with the folded instances.)
  const bool xf_fast_ok = XF && p.w_resident != 0 && p.f_regular[ph] != 0 && p.xf_fast != 0 &&
                          (p.f_nmma[ph] == 3 || p.f_nmma[ph] == 6 || p.f_nmma[ph] == 12);
  auto xf_issue = [&](auto NC) {
    constexpr int NR = decltype(NC)::value;
    const bool elected = elect_one();
    const bool leader = elected && !(p.dbg & 1);      // dbg bit 1: no MMAs (the commits still arrive)
    const uint32_t idesc = make_idesc_bf16(p.n_tile);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);   // SBO = 128 B, version 1
    const uint32_t w_base16 = (smem_u32(smem_w) & 0x3FFFF) >> 4;
    const uint32_t a_ring16 = (smem_u32(smem_a) & 0x3FFFF) >> 4;
    const uint32_t a_stage16 = p.a_stage_bytes >> 4;
    const uint32_t a_sub16 = p.a_sub_bytes >> 4;
    const uint32_t n_tile_u = static_cast<uint32_t>(p.n_tile);
    const uint32_t f_a_hi = p.f_a_hi[ph], f_a_lo_lbo = p.f_a_lo_lbo[ph], f_b_chunk16 = p.f_b_chunk16[ph];
    const uint32_t b_lo_lbo = (n_tile_u & 0x3FFF) << 16;
    for (int i = 0; i < kMaxWStages; ++i) {           // resident weights: requested before the dependency wait
      if (i * w_per >= phase.n_blocks) break;
      mbar_wait(&w_full[i], 0);
    }
    tc_fence_after();
    uint32_t off[NR];            // A offset | B offset << 16, as in the table (unpacked copies spill)
#pragma unroll
    for (int i = 0; i < NR; ++i) off[i] = p.f_off[ph][i];
    int s = 0;
    uint32_t a_par = 0;
    uint32_t a_base16 = a_ring16;
    bool stamp2 = tr != nullptr;
    for (int t = 0; t < my_tiles; ++t) {
      const int acc = t & 1;
      STCD_TWAIT(mbar_wait(&acc_empty[acc], ((t >> 1) & 1) ^ 1), tw_acc);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * p.acc_cols;
      uint32_t accum = 0;
      uint32_t b0 = b_lo_lbo + w_base16;
      for (int c = 0; c < phase.chunk_count; ++c) {
        STCD_TWAIT(mbar_wait(&a_full[s], a_par), tw_a);
        tc_fence_after();
        if (stamp2) {
          if (elected) STCD_STAMP(2);
          stamp2 = false;
        }
        const uint32_t a0 = f_a_lo_lbo + a_base16;
        if (leader) {
#pragma unroll
          for (int i = 0; i < NR; ++i) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              umma_bf16_lohi(d0 + m * n_tile_u, a0 + (off[i] & 0xFFFFu) + m * a_sub16, f_a_hi, b0 + (off[i] >> 16), b_hi, idesc, i == 0 ? accum : 1u);
          }
        }
        accum = 1;
        if (elected) umma_commit(&a_empty[s]);
        b0 += f_b_chunk16;
        a_base16 += a_stage16;
        if (++s == p.a_stages) {
          s = 0;
          a_par ^= 1;
          a_base16 = a_ring16;
        }
      }
      if (elected) {
        umma_commit(&acc_full[acc]);
        if (t == 0) STCD_STAMP(4);
        if (t == my_tiles - 1) STCD_STAMP(9);
      }
    }
  };

  if (warp == 0) {
    // ============================== A producer (TMA) ==============================
    const bool leader = elect_one();
    int s = 0;
    uint32_t par = 1;  // parity to wait for on a_empty[s]: first pass through the ring never blocks
    long long tw_prod = 0;
    pdl_wait();        // activations are written by the previous kernel(s)
    for (int t = 0; t < my_tiles; ++t) {
      int tile = blockIdx.x + t * gridDim.x;
      if (p.reverse) tile = p.n_tiles - 1 - tile;
      const int tile_x = tile % p.tiles_x;
      tile /= p.tiles_x;
      const int tile_y = tile % p.tiles_y;
      const int img = tile / p.tiles_y;
      const int x0 = tile_x * XSTEP - XOFF, y0 = tile_y * TH;
      for (int c = 0; c < phase.chunk_count; ++c) {
        const ChunkLoad L = s_cload[c];
        STCD_TWAIT(mbar_wait_relaxed(&a_empty[s], par), tw_prod);
        if (leader) {
          mbar_expect_tx(&a_full[s], L.tx_bytes);
          if (t == 0 && c == 0) STCD_STAMP(8);
          uint8_t* dst = smem_a + s * p.a_stage_bytes;
          const CUtensorMap* map = &tm.src[L.src_merged & 0xff];
          const int cx = x0 * L.xm + L.xa, cy = y0 * L.ym + L.ya, cn = img + L.n_off;
          if (L.src_merged >> 8) {
#pragma unroll
            for (int m = 0; m < MT; ++m) tma_load_4d(dst + m * p.a_sub_bytes, map, &a_full[s], cx, cy, L.c8, cn + p.m_off[m]);
          } else {
#pragma unroll
            for (int m = 0; m < MT; ++m) tma_load_5d(dst + m * p.a_sub_bytes, map, &a_full[s], 0, cx, cy, L.c8, cn + p.m_off[m]);
          }
        }
        if (++s == p.a_stages) {
          s = 0;
          par ^= 1;
        }
      }
    }
#ifdef STCD_TRACE_WAITS
    if (tr && leader) tr[15] = tw_prod;
#endif
    __syncwarp();
  } else if (warp == 1) {
    // ============================== W producer (bulk copies) ==============================
    const bool leader = elect_one();
    if (p.w_resident) {
      // everything once: barrier slot s covers blocks [s*per, (s+1)*per) so MMAs can start early
      if (leader) {
        for (int s = 0; s < kMaxWStages; ++s) {
          const int b0 = s * w_per, b1 = min(phase.n_blocks, b0 + w_per);
          if (b0 >= b1) break;
          const uint32_t bytes = static_cast<uint32_t>(b1 - b0) * p.wblk_bytes;
          mbar_expect_tx(&w_full[s], bytes);
          bulk_load(smem_w + static_cast<size_t>(b0) * p.wblk_bytes, wsrc + static_cast<size_t>(b0) * p.wblk_bytes, bytes,
                    &w_full[s]);
        }
      }
    } else {
      int s = 0;
      uint32_t par = 1;
      for (int t = 0; t < my_tiles; ++t) {
        for (int b = 0; b < phase.n_blocks; ++b) {
          mbar_wait_relaxed(&w_empty[s], par);
          if (leader) {
            mbar_expect_tx(&w_full[s], p.wblk_bytes);
            bulk_load(smem_w + static_cast<size_t>(s) * p.wblk_bytes, wsrc + static_cast<size_t>(b) * p.wblk_bytes,
                      p.wblk_bytes, &w_full[s]);
          }
          if (++s == p.w_stages) {
            s = 0;
            par ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ============================== MMA issuer ==============================
    const bool elected = elect_one();
    const bool leader = elected && !(p.dbg & 1);      // dbg bit 1: no MMAs (the commits still arrive)
    const uint32_t idesc = make_idesc_bf16(p.n_tile);
    const int ksteps = p.kc >> 4;                   // 1, 2 or 4
    const uint32_t b_hi = (128u >> 4) | (1u << 14);                                   // SBO = 128 B, version 1
    const uint32_t wblk16 = p.wblk_bytes >> 4;
    const uint32_t w_base16 = (smem_u32(smem_w) & 0x3FFFF) >> 4;
    const uint32_t a_ring16 = (smem_u32(smem_a) & 0x3FFFF) >> 4;
    const uint32_t a_stage16 = p.a_stage_bytes >> 4;
    const uint32_t a_sub16 = p.a_sub_bytes >> 4;
    const bool resident = p.w_resident != 0;
    const uint32_t n_tile_u = static_cast<uint32_t>(p.n_tile);
    const bool fast = resident && p.f_regular[ph] != 0;
    const int f_nmma = p.f_nmma[ph];
    const uint32_t f_a_hi = p.f_a_hi[ph], f_a_lo_lbo = p.f_a_lo_lbo[ph], f_b_chunk16 = p.f_b_chunk16[ph];
    const uint32_t b_lo_lbo = (n_tile_u & 0x3FFF) << 16;
    int s = 0, ws = 0;
    uint32_t a_par = 0, w_par = 0;
    uint32_t a_base16 = a_ring16;
    if (resident) {
      // weights were requested before the dependency wait: they land while the previous kernel
      // drains; take all of them before the first MMA so the issue loop never looks at them again
      for (int i = 0; i < kMaxWStages; ++i) {
        if (i * w_per >= phase.n_blocks) break;
        mbar_wait(&w_full[i], 0);
      }
      tc_fence_after();
    }
    bool done = false;
    if constexpr (XF) {
      if (xf_fast_ok) {
        if (f_nmma == 3) xf_issue(IntC<3>{});
        else if (f_nmma == 6) xf_issue(IntC<6>{});
        else xf_issue(IntC<12>{});
        done = true;
      }
    }
    for (int t = 0; t < (done ? 0 : my_tiles); ++t) {
      const int acc = t & 1;
      STCD_TWAIT(mbar_wait(&acc_empty[acc], ((t >> 1) & 1) ^ 1), tw_acc);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * p.acc_cols;
      uint32_t accum = 0;
      int mma_off = 0;
      for (int c = 0; c < phase.chunk_count; ++c) {
        const ChunkMma M = s_cmma[c];
        const int n_mma = static_cast<int>(M.n_mma);
        STCD_TWAIT(mbar_wait(&a_full[s], a_par), tw_a);
        tc_fence_after();
        if (t == 0 && c == 0 && elected) STCD_STAMP(2);
        if (fast) {
          // regular phase: operand offsets come from the constant bank through uniform loads
          const uint32_t a0 = f_a_lo_lbo + a_base16;
          const uint32_t b0 = b_lo_lbo + w_base16 + static_cast<uint32_t>(c) * f_b_chunk16;
          int i = 0;
          for (; i + 4 <= f_nmma; i += 4) {
            const uint4 e = *reinterpret_cast<const uint4*>(&p.f_off[ph][i]);   // one 16-byte constant load
            if (leader) {  // ONE divergent region per 4 (x MT) MMAs: the four operand chains overlap
#pragma unroll
              for (int m = 0; m < MT; ++m)
                umma_bf16_lohi(d0 + m * n_tile_u, a0 + (e.x & 0xFFFFu) + m * a_sub16, f_a_hi, b0 + (e.x >> 16), b_hi, idesc, accum);
#pragma unroll
              for (int m = 0; m < MT; ++m)
                umma_bf16_lohi(d0 + m * n_tile_u, a0 + (e.y & 0xFFFFu) + m * a_sub16, f_a_hi, b0 + (e.y >> 16), b_hi, idesc, 1u);
#pragma unroll
              for (int m = 0; m < MT; ++m)
                umma_bf16_lohi(d0 + m * n_tile_u, a0 + (e.z & 0xFFFFu) + m * a_sub16, f_a_hi, b0 + (e.z >> 16), b_hi, idesc, 1u);
#pragma unroll
              for (int m = 0; m < MT; ++m)
                umma_bf16_lohi(d0 + m * n_tile_u, a0 + (e.w & 0xFFFFu) + m * a_sub16, f_a_hi, b0 + (e.w >> 16), b_hi, idesc, 1u);
            }
            accum = 1;
          }
          if (i < f_nmma) {
            const uint4 e = *reinterpret_cast<const uint4*>(&p.f_off[ph][i]);   // entries past f_nmma are zero padding
            const uint32_t ev[3] = {e.x, e.y, e.z};
            if (leader) {
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                if (i + r < f_nmma) {
#pragma unroll
                  for (int m = 0; m < MT; ++m)
                    umma_bf16_lohi(d0 + m * n_tile_u, a0 + (ev[r] & 0xFFFFu) + m * a_sub16, f_a_hi, b0 + (ev[r] >> 16), b_hi, idesc,
                                   (r == 0) ? accum : 1u);
                }
              }
            }
            accum = 1;
          }
        } else if (resident) {
          // Lane i holds the descriptors of MMA i of this chunk; the issue loop only shuffles them
          // out, so its instructions are independent and pipeline instead of forming one chain.
          for (int base = 0; base < n_mma; base += 32) {
            const int cnt = min(32, n_mma - base);
            uint2 my = make_uint2(0u, 0u);
            if (lane < cnt) my = s_mma[mma_off + base + lane];
            my.x += a_base16;
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
              const uint32_t a_lo = __shfl_sync(0xffffffffu, my.x, j);
              const uint32_t b_lo = __shfl_sync(0xffffffffu, my.y, j);
              if (leader) {
#pragma unroll
                for (int m = 0; m < MT; ++m) umma_bf16_lohi(d0 + m * n_tile_u, a_lo + m * a_sub16, M.a_hi, b_lo, b_hi, idesc, accum);
              }
              accum = 1;
            }
          }
        } else {
          for (int i = 0; i < n_mma; i += ksteps) {  // one weight block (tap) per ring slot
            STCD_TWAIT(mbar_wait(&w_full[ws], w_par), tw_w);
            tc_fence_after();
            const uint32_t b_slot16 = w_base16 + static_cast<uint32_t>(ws) * wblk16;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint2 e = s_mma[mma_off + i + ks];
              if (leader) {
#pragma unroll
                for (int m = 0; m < MT; ++m)
                  umma_bf16_lohi(d0 + m * n_tile_u, e.x + a_base16 + m * a_sub16, M.a_hi, e.y + b_slot16, b_hi, idesc, accum);
              }
              accum = 1;
            }
            if (elected) umma_commit(&w_empty[ws]);
            if (++ws == p.w_stages) {
              ws = 0;
              w_par ^= 1;
            }
          }
        }
        mma_off += n_mma;
        if (elected) umma_commit(&a_empty[s]);
        a_base16 += a_stage16;
        if (++s == p.a_stages) {
          s = 0;
          a_par ^= 1;
          a_base16 = a_ring16;
        }
      }
      if (elected) {
        umma_commit(&acc_full[acc]);
        if (t == 0) STCD_STAMP(4);
        if (t == my_tiles - 1) STCD_STAMP(9);
      }
    }
#ifdef STCD_TRACE_WAITS
    if (tr && elected) {
      tr[12] = tw_a;
      tr[13] = tw_acc;
      tr[3] = tw_w;
    }
#endif
    __syncwarp();
  } else if (warp == kResWarp) {
    // ============================== residual producer (TMA) ==============================
    if (p.res_slots) {
      const bool leader = elect_one();
      int rs = 0;
      uint32_t rpar = 1;     // first pass through the ring never blocks
      const int n_ch = min(XF ? p.xf_cs : p.n_tile, p.cout - n0);          // channels the epilogue walks (in steps of 16)
      const int n_blocks = (n_ch + p.res_rb - 1) / p.res_rb;
      const uint32_t tx = static_cast<uint32_t>(MS) * static_cast<uint32_t>(p.res_rb) * (TH * TW * 2u);
      pdl_wait();            // the residual is written by an earlier kernel
      for (int t = 0; t < my_tiles; ++t) {
        int tile = blockIdx.x + t * gridDim.x;
        if (p.reverse) tile = p.n_tiles - 1 - tile;
        const int tile_x = tile % p.tiles_x;
        tile /= p.tiles_x;
        const int tile_y = tile % p.tiles_y;
        const int img = tile / p.tiles_y;
        const int x0 = tile_x * XSTEP - XOFF, y0 = tile_y * TH;
        for (int mb = 0; mb < MT; mb += MS) {
          for (int cb = 0; cb < n_blocks; ++cb) {     // the order the epilogue consumes: (sub-tile group, channel block)
            mbar_wait_relaxed(&r_empty[rs], rpar);
            if (leader) {
              mbar_expect_tx(&r_full[rs], tx);
              uint8_t* dst = smem_r + static_cast<size_t>(rs) * p.res_slot_bytes;
#pragma unroll
              for (int m = 0; m < MS; ++m)
                tma_load_4d(dst + m * p.res_sub_bytes, &tm.res, &r_full[rs], x0 * 8, y0, (n0 + cb * p.res_rb) >> 3, img + p.m_off[mb + m]);
            }
            if (++rs == p.res_slots) {
              rs = 0;
              rpar ^= 1;
            }
          }
        }
      }
      __syncwarp();
    }
  } else if (warp < kResWarp) {
    // ============================== epilogue (warps 3..6) ==============================
    const bool has_raw = STCD_HAS(E_RAW, p.out_raw != nullptr);
    const bool has_aff2 = STCD_HAS(E_AFF2, p.scale2 != nullptr);
    const bool has_res = STCD_HAS(E_RES, p.res != nullptr);
    const bool split = (EPI & E_GENERIC) != 0 && p.split != 0;         // split precision: generic instances only
    const bool res_sm = has_res && STCD_HAS(E_RSM, p.res_slots > 0);   // residual tile in the shared-memory ring
    const bool res_rg = has_res && !res_sm && !split;                  // residual through prefetched per-thread global loads
    const int rb_mask = p.res_rb - 1;                                  // res_rb is a power of two (16 / 32 / 64)
    // 16 values -> the hi plane at `o` (two 8-channel groups `plane` elements apart) and, in split precision, the lo plane
    // `lo_off` elements further
    auto store16 = [&](__nv_bfloat16* o, size_t plane, size_t lo_off, const float* x, bool both) {
      *reinterpret_cast<uint4*>(o) = pack8_bf16(x);
      if (both) *reinterpret_cast<uint4*>(o + plane) = pack8_bf16(x + 8);
      if (split) {
        *reinterpret_cast<uint4*>(o + lo_off) = pack8_bf16_lo(x);
        if (both) *reinterpret_cast<uint4*>(o + lo_off + plane) = pack8_bf16_lo(x + 8);
      }
    };
    const bool has_relu = STCD_HAS(E_RELU, p.relu != 0);
    const bool has_out0 = STCD_HAS(E_OUT0, p.out0 != nullptr);
    const bool has_pool = STCD_HAS(E_POOL, p.out_pool != nullptr);
    const bool has_diff = (MS == 2) && STCD_HAS(E_DIFF, p.out_diff != nullptr);
    const bool has_f32 = STCD_HAS(E_F32, p.out_f32 != nullptr);
    // ReLU-only instances keep the lean path; GELU / PReLU / act-before-BN live behind E_ACTX so that their code
    // (erff, the extra selects) costs the ReLU nets neither registers nor instructions
    const bool has_actx = STCD_HAS(E_ACTX, p.relu >= 2 || p.act_pre != 0);
    const bool act_pre = has_actx && p.act_pre != 0;
    auto apply_act = [&](float (&x)[16]) {   // p.relu is warp-uniform
      if (!has_actx || p.relu == 1) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
      } else if (p.relu == 2) {              // nn.GELU(): 0.5 x (1 + erf(x / sqrt 2))
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float2 g2 = gelu_fast2(make_float2(x[j], x[j + 1]));
          x[j] = g2.x, x[j + 1] = g2.y;
        }
      } else if (p.relu == 3) {              // nn.PReLU() with one slope
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = x[j] >= 0.f ? x[j] : x[j] * p.act_alpha;
      }
    };

    // the residual is prefetched ahead of the accumulator wait, so this role reads global memory
    // written by the previous kernel without going through the A producer's dependency wait
    if (res_rg || (has_res && split)) pdl_wait();
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int my_half = (warp - 3) >> 2;   // NE == 8: the two warps of a quarter take alternate 16-column steps
    // pixel of the tile this thread (= TMEM lane 32 * wq + lane) owns: 4 lines of 8 per warp, or (XF) 2 lines of 16
    const int ty = XF ? 2 * wq + (lane >> 4) : 4 * wq + (lane >> 3);
    const int tx = XF ? (lane & 15) : (lane & 7);
    const bool lane_out = !XF || (tx >= 1 && tx <= kXfStep);   // XF: the outer two columns only feed their neighbours
    const int c_lim = XF ? p.xf_cs : p.n_tile;                 // output channels this CTA finishes
    const int n_ch_epi = min(c_lim, p.cout - n0);              // ... of which real ones
    const uint32_t hw = static_cast<uint32_t>(p.ho) * p.wo;  // pixels per image plane
    const uint32_t hw_pool = hw >> 2;
    const int cg8 = n0 >> 3;  // first 8-channel group of this N tile
    // The residual does not depend on the accumulator.  With short K the epilogue is the pacing role (ncu: tensor pipe 50 %
    // busy, epilogue warps on the long scoreboard), so its global-load latency must never be exposed: the first PF
    // 16-channel steps of every (tile, sub-tile group) are fetched ONE WHOLE GROUP ahead -- issued before the wait for the
    // previous group's accumulator -- and the later steps one step ahead.
    constexpr int PF = (MS == 1) ? 2 : 1;   // deeper prefetch spills: the kernel sits at its 128-register budget
    uint4 r_pre[PF][MS][2];
    auto prefetch_unit = [&](int t2, int mb2) {
      int tile2 = blockIdx.x + t2 * gridDim.x;
      if (p.reverse) tile2 = p.n_tiles - 1 - tile2;
      const int tile_x2 = tile2 % p.tiles_x;
      tile2 /= p.tiles_x;
      const int tile_y2 = tile2 % p.tiles_y;
      const int img2 = tile2 / p.tiles_y;
      const int gy2 = tile_y2 * TH + ty, gx2 = tile_x2 * XSTEP + tx - XOFF;
      if (gy2 < p.hg && gx2 < p.wg && lane_out) {
        const uint32_t pix2 = static_cast<uint32_t>(gy2 * p.osy + phase.oy) * p.wo + (gx2 * p.osx + phase.ox);
#pragma unroll
        for (int sidx = 0; sidx < PF; ++sidx) {
          const int c0 = 16 * sidx;
          if (c0 < c_lim && n0 + c0 < p.cout) {
#pragma unroll
            for (int m = 0; m < MS; ++m) {
              const __nv_bfloat16* r = p.res + ((static_cast<size_t>(img2 + p.m_off[mb2 + m]) * p.res_c8 + cg8 + 2 * sidx) * hw + pix2) * 8;
              r_pre[sidx][m][0] = __ldg(reinterpret_cast<const uint4*>(r));
              r_pre[sidx][m][1] = (p.cout - (n0 + c0) > 8) ? __ldg(reinterpret_cast<const uint4*>(r + static_cast<size_t>(hw) * 8))
                                                          : make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
      }
    };
    if (res_rg && my_tiles > 0) prefetch_unit(0, 0);
    int ers = 0;
    uint32_t erpar = 0;
    long long tw_epi = 0;
    for (int t = 0; t < my_tiles; ++t) {
      int tile = blockIdx.x + t * gridDim.x;
      if (p.reverse) tile = p.n_tiles - 1 - tile;
      const int tile_x = tile % p.tiles_x;
      tile /= p.tiles_x;
      const int tile_y = tile % p.tiles_y;
      const int img = tile / p.tiles_y;
      const int gy = tile_y * TH + ty, gx = tile_x * XSTEP + tx - XOFF;
      const bool valid = (gy < p.hg) && (gx < p.wg) && lane_out && !(p.dbg & 2);   // dbg bit 2: no epilogue stores (and no residual adds)
      const int oy = gy * p.osy + phase.oy, ox = gx * p.osx + phase.ox;
      const uint32_t pix = static_cast<uint32_t>(oy) * p.wo + ox;
      const uint32_t pix_pool = static_cast<uint32_t>(oy >> 1) * (p.wo >> 1) + (ox >> 1);
      // element pointers of (image 0, first channel group of this N tile, pixel)
      // out0 may be stored space-to-depth: plane of hw/4 pixels, parity class (oy&1, ox&1) selects the channel block
      const uint32_t hw0 = p.out0_s2d ? (hw >> 2) : hw;
      const uint32_t pix0 = p.out0_s2d ? pix_pool : pix;
      const uint32_t cg0 = p.out0_s2d ? static_cast<uint32_t>(((oy & 1) * 2 + (ox & 1)) * (p.cout >> 3)) : static_cast<uint32_t>(p.out0_coff >> 3);
      __nv_bfloat16* q_out0 = has_out0 ? p.out0 + ((static_cast<size_t>(cg0) + cg8) * hw0 + pix0) * 8 : nullptr;
      __nv_bfloat16* q_raw = has_raw ? p.out_raw + (static_cast<size_t>(cg8) * hw + pix) * 8 : nullptr;
      const __nv_bfloat16* q_res = res_rg ? p.res + (static_cast<size_t>(cg8) * hw + pix) * 8 : nullptr;
      __nv_bfloat16* q_pool = has_pool ? p.out_pool + (static_cast<size_t>(cg8) * hw_pool + pix_pool) * 8 : nullptr;
      __nv_bfloat16* q_diff = has_diff ? p.out_diff + (static_cast<size_t>(cg8) * hw + pix) * 8 : nullptr;
      const int acc = t & 1;
      const uint32_t tlane = tmem_base + acc * p.acc_cols + (static_cast<uint32_t>(wq * 32) << 16);
#pragma unroll
      for (int mb = 0; mb < MT; mb += MS) {  // sub-tiles one at a time, Siamese pairs two at a time
        size_t im[MS];                       // image index of each sub-tile of this group
#pragma unroll
        for (int m = 0; m < MS; ++m) im[m] = static_cast<size_t>(img + p.m_off[mb + m]);
        // Per-(sub-tile) base pointers and per-step offsets that only ADD inside the channel loop: the 64-bit index products
        // were recomputed every 16-column step (75 IMADs per step in the SASS of the residual instance -- dependent chains that
        // one epilogue warp per scheduler cannot hide).
        __nv_bfloat16* b_out0[MS];
        __nv_bfloat16* b_raw[MS];
        __nv_bfloat16* b_pool[MS];
#pragma unroll
        for (int m = 0; m < MS; ++m) {
          b_out0[m] = has_out0 ? q_out0 + im[m] * p.out0_c8 * hw0 * 8 : nullptr;
          b_raw[m] = has_raw ? q_raw + im[m] * p.out_raw_c8 * hw * 8 : nullptr;
          b_pool[m] = has_pool ? q_pool + im[m] * p.out_pool_c8 * hw_pool * 8 : nullptr;
        }
        __nv_bfloat16* const b_diff = has_diff ? q_diff + im[0] * p.out_diff_c8 * hw * 8 : nullptr;
        const size_t st_hw = static_cast<size_t>(hw) * 16, st_hw0 = static_cast<size_t>(hw0) * 16, st_pool = static_cast<size_t>(hw_pool) * 16;
        size_t off_hw = 0, off_hw0 = 0, off_pool = 0;      // element offset of the step's first 8-channel group in planes of hw / hw0 / hw_pool pixels
        uint4 r_cur[MS][2], r_nxt[MS][2], r_pf[PF > 1 ? PF - 1 : 1][MS][2];
        auto load_res = [&](int c0, uint4 (&dst)[MS][2]) {
#pragma unroll
          for (int m = 0; m < MS; ++m) {
            const __nv_bfloat16* r = q_res + (im[m] * p.res_c8 + (c0 >> 3)) * hw * 8;
            dst[m][0] = __ldg(reinterpret_cast<const uint4*>(r));
            dst[m][1] = (p.cout - (n0 + c0) > 8) ? __ldg(reinterpret_cast<const uint4*>(r + static_cast<size_t>(hw) * 8))
                                                  : make_uint4(0u, 0u, 0u, 0u);
          }
        };
        if (res_rg) {
#pragma unroll
          for (int m = 0; m < MS; ++m) {
            r_cur[m][0] = r_pre[0][m][0];
            r_cur[m][1] = r_pre[0][m][1];
#pragma unroll
            for (int q = 1; q < PF; ++q) {
              r_pf[q - 1][m][0] = r_pre[q][m][0];
              r_pf[q - 1][m][1] = r_pre[q][m][1];
            }
          }
          if (mb + MS < MT)                 // the next group's first steps, in flight while this group is processed
            prefetch_unit(t, mb + MS);
          else if (t + 1 < my_tiles)
            prefetch_unit(t + 1, 0);
        }
        if (mb == 0) {
          if (res_sm) mbar_wait_relaxed(&r_full[ers], erpar);   // the tile's first residual block: off the critical path here
          STCD_TWAIT(mbar_wait_relaxed(&acc_full[acc], (t >> 1) & 1), tw_epi);
          tc_fence_after();
          if (t == 0 && threadIdx.x == 96) STCD_STAMP(5);
        }
        const uint32_t tgrp = tlane + mb * p.n_tile;

        // Up-sampling ops with the horizontal output phases folded into N (osx = 2): column blocks 2q and 2q + 1 are the
        // two horizontally adjacent output pixels (row phase q) of this thread's input pixel -- 32 contiguous bytes per
        // channel group, written with ONE 256-bit store.  Two 16-byte stores (or two phases in different CTAs) write each
        // 32-byte sector in halves, and partially written sectors cost the memory system a fill: SNUNet's Up1_x ran at
        // 263 us with its stores and 85 us without them.
        constexpr bool kMayFold2 = ((EPI & E_GENERIC) != 0 || (EPI & (E_RAW | E_AFF2 | E_RES | E_POOL | E_DIFF | E_F32)) == 0) && !XF && MS == 1;
        if (kMayFold2 && p.fold_cs && p.osx == 2) {
          const int n_rows = (p.cout / p.fold_cs) >> 1;      // row phases in this GEMM phase: 2 (all phases folded) or 1
          const size_t imf = static_cast<size_t>(img + p.m_off[mb]);
          for (int q = 0; q < n_rows; ++q) {
            const uint32_t pixf = static_cast<uint32_t>(gy * p.osy + phase.oy + q) * p.wo + (gx * 2 + phase.ox);
            for (int c0 = 0; c0 < p.fold_cs; c0 += 16) {
              if (c0 >= p.fold_cout) break;
              const int col_a = 2 * q * p.fold_cs + c0, col_b = col_a + p.fold_cs;
              uint32_t ra[16], rb[16];
              tmem_ld16(tgrp + col_a, ra);
              tmem_ld16(tgrp + col_b, rb);
              tmem_wait_ld();
              float va[16], vb[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                va[j] = fmaf(__uint_as_float(ra[j]), s_aff[0][col_a + j], s_aff[1][col_a + j]);
                vb[j] = fmaf(__uint_as_float(rb[j]), s_aff[0][col_b + j], s_aff[1][col_b + j]);
              }
              if (has_relu) {
                apply_act(va);
                apply_act(vb);
              }
              if (valid) {
                __nv_bfloat16* o = p.out0 + ((imf * p.out0_c8 + ((p.out0_coff + c0) >> 3)) * hw + pixf) * 8;
                st_global_256(o, pack8_bf16(va), pack8_bf16(vb));
                if (p.fold_cout - c0 > 8) st_global_256(o + static_cast<size_t>(hw) * 8, pack8_bf16(va + 8), pack8_bf16(vb + 8));
              }
            }
          }
          continue;
        }

        for (int c0 = 0; c0 < c_lim; c0 += 16) {
          const int ch = n0 + c0;
          if (ch >= p.cout) break;
          const bool two = (p.cout - ch) > 8;  // both 8-channel groups of this 16-column step are real
          const bool mine = (NE == 4) || (((c0 >> 4) & 1) == my_half);   // NE == 8: the other warp of this quarter does the odd / even steps
          if (mine) {
          if (res_rg && valid && c0 + 16 >= 16 * PF && c0 + 16 < c_lim && ch + 16 < p.cout) load_res(c0 + 16, r_nxt);
          uint32_t raw[MS][16];
          // Everything that comes from shared memory -- the column affines and, with the residual ring, the residual -- is
          // requested while the accumulator load is in flight: one epilogue warp per scheduler has nothing else to hide their
          // latency with (ncu source view, round 2: the step's FFMAs waited on the short scoreboard for the affine LDS, the
          // unpack on the long one for the residual read through a generic pointer).
          if (!XF) {
#pragma unroll
            for (int m = 0; m < MS; ++m) tmem_ld16(tgrp + m * p.n_tile + c0, raw[m]);
          }
          float4 a_sc[4], a_sh[4], a_sc2[4], a_sh2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            a_sc[j] = *reinterpret_cast<const float4*>(&s_aff[0][c0 + 4 * j]);
            a_sh[j] = *reinterpret_cast<const float4*>(&s_aff[1][c0 + 4 * j]);
            if (has_aff2) {
              a_sc2[j] = *reinterpret_cast<const float4*>(&s_aff[2][c0 + 4 * j]);
              a_sh2[j] = *reinterpret_cast<const float4*>(&s_aff[3][c0 + 4 * j]);
            }
          }
          if (res_sm) {
            // slot in shared memory: [MS sub-tiles][res_rb / 8][tile rows][tile px] x 16 B -- consecutive lanes, consecutive 16 B
            if ((c0 & rb_mask) == 0 || (NE == 8 && (c0 & rb_mask) == 16)) mbar_wait_relaxed(&r_full[ers], erpar);     // a new block of this group (NE == 8: this warp's first step of it)
            const uint32_t rt = smem_u32(smem_r) + static_cast<uint32_t>(ers) * p.res_slot_bytes +
                                (static_cast<uint32_t>((c0 & rb_mask) >> 3) * (TH * TW) + ty * TW + tx) * 16u;
#pragma unroll
            for (int m = 0; m < MS; ++m) {
              r_cur[m][0] = lds128(rt + m * p.res_sub_bytes);
              r_cur[m][1] = lds128(rt + m * p.res_sub_bytes + TH * TW * 16);
            }
          }
          if (!XF) {
            tmem_wait_ld();
          } else {
            // out[x] = D_0[x - 1] + D_1[x] + D_2[x + 1]: column block b sits at column b * cs; the left / right neighbours
            // are the adjacent lanes (a warp holds two tile lines of 16 positions; lanes 0 and 15 of a line never output)
#pragma unroll
            for (int m = 0; m < MS; ++m) {
              uint32_t dl[16], dr[16];
              tmem_ld16(tgrp + m * p.n_tile + c0, dl);
              tmem_ld16(tgrp + m * p.n_tile + p.xf_cs + c0, raw[m]);
              tmem_ld16(tgrp + m * p.n_tile + 2 * p.xf_cs + c0, dr);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float l = __shfl_up_sync(0xffffffffu, __uint_as_float(dl[j]), 1);
                const float r = __shfl_down_sync(0xffffffffu, __uint_as_float(dr[j]), 1);
                raw[m][j] = __float_as_uint((l + __uint_as_float(raw[m][j])) + r);
              }
            }
          }
          float v[MS][16];
          const size_t g8 = static_cast<size_t>(c0 >> 3);  // channel-group offset inside the N tile
#pragma unroll
          for (int m = 0; m < MS; ++m) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              v[m][4 * j + 0] = fmaf(__uint_as_float(raw[m][4 * j + 0]), a_sc[j].x, a_sh[j].x);
              v[m][4 * j + 1] = fmaf(__uint_as_float(raw[m][4 * j + 1]), a_sc[j].y, a_sh[j].y);
              v[m][4 * j + 2] = fmaf(__uint_as_float(raw[m][4 * j + 2]), a_sc[j].z, a_sh[j].z);
              v[m][4 * j + 3] = fmaf(__uint_as_float(raw[m][4 * j + 3]), a_sc[j].w, a_sh[j].w);
            }
            if (has_raw && valid) {
              store16(b_raw[m] + off_hw, static_cast<size_t>(hw) * 8, static_cast<size_t>(p.out_raw_c8 >> 1) * hw * 8, v[m], two);
            }
            if (has_aff2) {
              if (act_pre) apply_act(v[m]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[m][4 * j + 0] = fmaf(v[m][4 * j + 0], a_sc2[j].x, a_sh2[j].x);
                v[m][4 * j + 1] = fmaf(v[m][4 * j + 1], a_sc2[j].y, a_sh2[j].y);
                v[m][4 * j + 2] = fmaf(v[m][4 * j + 2], a_sc2[j].z, a_sh2[j].z);
                v[m][4 * j + 3] = fmaf(v[m][4 * j + 3], a_sc2[j].w, a_sh2[j].w);
              }
            }
            if (has_res && valid && split) {
              // residual = hi + lo, read where it is needed (the precision path is not the throughput path)
              const __nv_bfloat16* r = p.res + ((im[m] * p.res_c8 + cg8 + g8) * hw + pix) * 8;
              const size_t lo_off = static_cast<size_t>(p.res_c8 >> 1) * hw * 8;
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                if (half == 0 || two) {
                  float a[8], b[8];
                  unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(r + half * static_cast<size_t>(hw) * 8)), a);
                  unpack8_bf16(__ldg(reinterpret_cast<const uint4*>(r + half * static_cast<size_t>(hw) * 8 + lo_off)), b);
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[m][8 * half + j] += a[j] + b[j];
                }
              }
            } else if (has_res && valid) {
              float rv[16];
              unpack8_bf16(r_cur[m][0], rv);
              unpack8_bf16(r_cur[m][1], rv + 8);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[m][j] += rv[j];
            }
            if (has_relu && !act_pre) apply_act(v[m]);
            if (has_f32) {
              if (valid && im[m] < static_cast<size_t>(p.n_valid)) {
                const int nv = min(16, p.cout - ch);
                float* o = p.out_f32 + (im[m] * p.cout + ch) * hw + pix;
                for (int j = 0; j < nv; ++j) o[static_cast<size_t>(j) * hw] = v[m][j];
              }
            }
            if ((has_out0 || has_pool) && split) {
              if (has_out0 && valid) {
                store16(b_out0[m] + off_hw0, static_cast<size_t>(hw0) * 8, static_cast<size_t>(p.out0_c8 >> 1) * hw0 * 8, v[m], two);
              }
              if (has_pool) {
                float vp[16];      // the pooled value must be formed in fp32: max does not commute with the (hi, lo) split
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float a = fmaxf(v[m][j], __shfl_down_sync(0xffffffffu, v[m][j], 1));
                  vp[j] = fmaxf(a, __shfl_down_sync(0xffffffffu, a, XF ? 16 : 8));
                }
                if (valid && (XF ? ((lane & 17) == 1) : ((lane & 9) == 0))) {
                  store16(b_pool[m] + off_pool, static_cast<size_t>(hw_pool) * 8, static_cast<size_t>(p.out_pool_c8 >> 1) * hw_pool * 8, vp, two);
                }
              }
            } else if (has_out0 || has_pool) {
              const uint4 lo = pack8_bf16(v[m]), hi = pack8_bf16(v[m] + 8);
              if (has_out0 && valid) {
                if (p.fold_cs) {
                  // phases folded into N: this 16-column step belongs to phase fp, channels [chn, chn + 16)
                  const int fp = ch / p.fold_cs, chn = ch - fp * p.fold_cs;
                  if (chn < p.fold_cout) {
                    const int fy = fp / p.osx;   // all phases folded: GEMM phase (0, 0); horizontal phases only: (oy, 0), fp < osx
                    const uint32_t pixf = static_cast<uint32_t>(gy * p.osy + phase.oy + fy) * p.wo + (gx * p.osx + phase.ox + fp - fy * p.osx);
                    __nv_bfloat16* o = p.out0 + ((im[m] * p.out0_c8 + ((p.out0_coff + chn) >> 3)) * hw + pixf) * 8;
                    *reinterpret_cast<uint4*>(o) = lo;
                    if (p.fold_cout - chn > 8) *reinterpret_cast<uint4*>(o + static_cast<size_t>(hw) * 8) = hi;
                  }
                } else {
                  __nv_bfloat16* o = b_out0[m] + off_hw0;
                  *reinterpret_cast<uint4*>(o) = lo;
                  if (two) *reinterpret_cast<uint4*>(o + static_cast<size_t>(hw0) * 8) = hi;
                }
              }
              if (has_pool) {
                // max over the 2x2 window on the packed bf16 pairs (rounding is monotonic, so
                // max(bf16(a), bf16(b)) == bf16(max(a, b))): lanes +1 (x) and +8 (y); XF tiles: +1 and +16, and the first
                // output column of a tile is lane 1 (tile origins are even, so even output columns sit on odd lanes)
                const uint4 plo = pool4_bf16x2<1, XF ? 16 : 8>(lo), phi = pool4_bf16x2<1, XF ? 16 : 8>(hi);
                if (valid && (XF ? ((lane & 17) == 1) : ((lane & 9) == 0))) {
                  __nv_bfloat16* o = b_pool[m] + off_pool;
                  *reinterpret_cast<uint4*>(o) = plo;
                  if (two) *reinterpret_cast<uint4*>(o + static_cast<size_t>(hw_pool) * 8) = phi;
                }
              }
            }
          }
          if (has_diff && valid) {
            float d[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) d[j] = fabsf(v[0][j] - v[MS - 1][j]);
            store16(b_diff + off_hw, static_cast<size_t>(hw) * 8, static_cast<size_t>(p.out_diff_c8 >> 1) * hw * 8, d, two);
          }
          if (res_rg) {
            const int nstep = (c0 >> 4) + 1;               // steps 1 .. PF-1 were prefetched with step 0
#pragma unroll
            for (int m = 0; m < MS; ++m) {
              uint4 a = r_nxt[m][0], b = r_nxt[m][1];
#pragma unroll
              for (int q = 1; q < PF; ++q)
                if (nstep == q) {
                  a = r_pf[q - 1][m][0];
                  b = r_pf[q - 1][m][1];
                }
              r_cur[m][0] = a;
              r_cur[m][1] = b;
            }
          }
          }   // mine
          off_hw += st_hw;
          off_hw0 += st_hw0;
          off_pool += st_pool;
          if (res_sm && (((c0 + 16) & rb_mask) == 0 || c0 + 16 >= n_ch_epi)) {
            // the last step of this residual block: hand the slot back to the residual producer
            __syncwarp();
            if (lane == 0) mbar_arrive(&r_empty[ers]);
            if (++ers == p.res_slots) {
              ers = 0;
              erpar ^= 1;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (threadIdx.x == 96) {
        if (t == 0) STCD_STAMP(6);
        if (t == my_tiles - 1) STCD_STAMP(10);
      }
    }
#ifdef STCD_TRACE_WAITS
    if (tr && threadIdx.x == 96) tr[14] = tw_epi;
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    STCD_STAMP(7);
    if (tr) tr[11] = my_tiles;
  }
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

#undef STCD_HAS

using ConvKernelFn = void (*)(const TmapPack, const ConvParams);

struct ConvKernelEntry {
  int mt, ms;
  uint32_t epi;
  ConvKernelFn fn;
  int ne;        // epilogue warps (4 or 8)
};

// Specialised instances for the (M tiles, pairing, epilogue) combinations the lowered nets use;
// anything else runs the generic instance (same code, epilogue features read from ConvParams at
// run time).  The instances are compiled in six translation units (conv_inst_{a..f}.cu, built in
// parallel); conv_kernel_table() in stcd_b200.cu concatenates their tables.  X(MT, MS, EPI)
#define STCD_CONV_INSTANCES_A(X)                       \
  /* FC-Siam encoder (Siamese pairs) */                \
  X(2, 2, E_RELU | E_OUT0)                             \
  X(2, 2, E_RELU | E_POOL | E_DIFF)                    \
  X(2, 2, E_RELU | E_OUT0 | E_POOL)                    \
  X(4, 2, E_RELU | E_OUT0)                             \
  X(4, 2, E_RELU | E_POOL | E_DIFF)                    \
  X(4, 2, E_RELU | E_OUT0 | E_POOL)                    \
  /* decoders, up-convs, logits */                     \
  X(1, 1, E_OUT0)                                      \
  X(2, 1, E_OUT0)                                      \
  X(1, 1, E_RELU | E_OUT0)                             \
  X(2, 1, E_RELU | E_OUT0)                             \
  X(4, 1, E_RELU | E_OUT0)                             \
  X(1, 1, E_F32)                                       \
  X(2, 1, E_F32)                                       \
  X(2, 2, E_OUT0)                                      \
  X(4, 2, E_OUT0)                                      \
  X(4, 1, E_OUT0)

#define STCD_CONV_INSTANCES_B(X)                       \
  /* SNUNet nested blocks: conv1 (raw identity + BN/ReLU), conv2 (+ identity; residual from the smem ring or by loads) */ \
  X(2, 2, E_RAW | E_AFF2 | E_RELU | E_OUT0)            \
  X(4, 2, E_RAW | E_AFF2 | E_RELU | E_OUT0)            \
  X(1, 1, E_RAW | E_AFF2 | E_RELU | E_OUT0)            \
  X(2, 1, E_RAW | E_AFF2 | E_RELU | E_OUT0)            \
  X(4, 1, E_RAW | E_AFF2 | E_RELU | E_OUT0)            \
  X(2, 2, E_RES | E_RELU | E_OUT0 | E_POOL)            \
  X(4, 2, E_RES | E_RELU | E_OUT0 | E_POOL)            \
  X(2, 2, E_RES | E_RSM | E_RELU | E_OUT0 | E_POOL)    \
  X(4, 2, E_RES | E_RSM | E_RELU | E_OUT0 | E_POOL)    \
  X(1, 1, E_RES | E_RELU | E_OUT0)                     \
  X(2, 1, E_RES | E_RELU | E_OUT0)                     \
  X(4, 1, E_RES | E_RELU | E_OUT0)                     \
  X(1, 1, E_RES | E_RSM | E_RELU | E_OUT0)             \
  X(2, 1, E_RES | E_RSM | E_RELU | E_OUT0)             \
  X(4, 1, E_RES | E_RSM | E_RELU | E_OUT0)

#define STCD_CONV_INSTANCES_C(X)                       \
  /* SegCD: BasicBlock conv2 (+identity, ReLU), Siamese pairs */ \
  X(2, 2, E_RES | E_RELU | E_OUT0)                     \
  X(4, 2, E_RES | E_RELU | E_OUT0)                     \
  X(2, 2, E_RES | E_RSM | E_RELU | E_OUT0)             \
  X(4, 2, E_RES | E_RSM | E_RELU | E_OUT0)             \
  /* ChangeGNN / ChangeFormer: 1x1 conv + residual without activation */ \
  X(2, 2, E_RES | E_OUT0)                              \
  X(4, 2, E_RES | E_OUT0)                              \
  X(1, 1, E_RES | E_OUT0)                              \
  X(2, 1, E_RES | E_OUT0)                              \
  X(4, 1, E_RES | E_OUT0)                              \
  X(2, 2, E_RES | E_RSM | E_OUT0)                      \
  X(4, 2, E_RES | E_RSM | E_OUT0)                      \
  X(1, 1, E_RES | E_RSM | E_OUT0)                      \
  X(2, 1, E_RES | E_RSM | E_OUT0)                      \
  X(4, 1, E_RES | E_RSM | E_OUT0)

#define STCD_CONV_INSTANCES_D(X)                       \
  /* GELU convs; conv -> act -> BN (+ residual) decoders */ \
  X(2, 2, E_RELU | E_ACTX | E_OUT0)                    \
  X(4, 2, E_RELU | E_ACTX | E_OUT0)                    \
  X(1, 1, E_AFF2 | E_RELU | E_ACTX | E_OUT0)           \
  X(2, 1, E_AFF2 | E_RELU | E_ACTX | E_OUT0)           \
  X(4, 1, E_AFF2 | E_RELU | E_ACTX | E_OUT0)           \
  X(1, 1, E_AFF2 | E_RES | E_RELU | E_ACTX | E_OUT0)   \
  X(2, 1, E_AFF2 | E_RES | E_RELU | E_ACTX | E_OUT0)   \
  X(4, 1, E_AFF2 | E_RES | E_RELU | E_ACTX | E_OUT0)   \
  X(1, 1, E_AFF2 | E_RES | E_RSM | E_RELU | E_ACTX | E_OUT0)   \
  X(2, 1, E_AFF2 | E_RES | E_RSM | E_RELU | E_ACTX | E_OUT0)   \
  X(4, 1, E_AFF2 | E_RES | E_RSM | E_RELU | E_ACTX | E_OUT0)   \
  /* generic: epilogue features read from ConvParams at run time */ \
  X(1, 1, E_GENERIC)                                   \
  X(2, 1, E_GENERIC)                                   \
  X(2, 2, E_GENERIC)                                   \
  X(4, 1, E_GENERIC)                                   \
  X(4, 2, E_GENERIC)

#define STCD_CONV_INSTANCES_E(X)                       \
  /* horizontally folded 3x3 convs (E_XF), one stream: decoders, nested-block nodes, logits.  The lowering never folds      \
     Siamese-pair ops (twice the accumulator columns: one CTA per SM, measured 2x slower), so there are no MS = 2 instances */ \
  X(1, 1, E_XF | E_RELU | E_OUT0)                      \
  X(2, 1, E_XF | E_RELU | E_OUT0)                      \
  X(4, 1, E_XF | E_RELU | E_OUT0)                      \
  X(1, 1, E_XF | E_RAW | E_AFF2 | E_RELU | E_OUT0)     \
  X(2, 1, E_XF | E_RAW | E_AFF2 | E_RELU | E_OUT0)     \
  X(4, 1, E_XF | E_RAW | E_AFF2 | E_RELU | E_OUT0)     \
  X(1, 1, E_XF | E_RES | E_RSM | E_RELU | E_OUT0)      \
  X(2, 1, E_XF | E_RES | E_RSM | E_RELU | E_OUT0)      \
  X(4, 1, E_XF | E_RES | E_RSM | E_RELU | E_OUT0)      \
  X(1, 1, E_XF | E_F32)                                \
  X(2, 1, E_XF | E_F32)                                \
  X(1, 1, E_XF | E_GENERIC)                            \
  X(2, 1, E_XF | E_GENERIC)                            \
  X(4, 1, E_XF | E_GENERIC)

#define STCD_CONV_INSTANCES_F(X)                       \
  /* eight epilogue warps (one CTA per SM): the short-K residual layers of the nested blocks and of the ResNet encoders */ \
  X(1, 1, E_RES | E_RSM | E_RELU | E_OUT0)             \
  X(2, 1, E_RES | E_RSM | E_RELU | E_OUT0)             \
  X(4, 1, E_RES | E_RSM | E_RELU | E_OUT0)             \
  X(2, 2, E_RES | E_RSM | E_RELU | E_OUT0 | E_POOL)    \
  X(4, 2, E_RES | E_RSM | E_RELU | E_OUT0 | E_POOL)    \
  X(2, 2, E_RES | E_RSM | E_RELU | E_OUT0)             \
  X(4, 2, E_RES | E_RSM | E_RELU | E_OUT0)             \
  /* horizontally folded layers at one CTA per SM: the 3x accumulator read + shuffles make their epilogue a pacing role */ \
  X(1, 1, E_XF | E_RELU | E_OUT0)                      \
  X(2, 1, E_XF | E_RELU | E_OUT0)                      \
  X(1, 1, E_XF | E_RAW | E_AFF2 | E_RELU | E_OUT0)     \
  X(2, 1, E_XF | E_RAW | E_AFF2 | E_RELU | E_OUT0)

// one table per translation unit
const ConvKernelEntry* conv_kernel_table_f(int* n);
const ConvKernelEntry* conv_kernel_table_e(int* n);
const ConvKernelEntry* conv_kernel_table_a(int* n);
const ConvKernelEntry* conv_kernel_table_b(int* n);
const ConvKernelEntry* conv_kernel_table_c(int* n);
const ConvKernelEntry* conv_kernel_table_d(int* n);

#define STCD_DEFINE_CONV_TABLE_WITH(NAME, LIST, ENTRY)                                       \
  const ConvKernelEntry* NAME(int* n) {                                                      \
    static const ConvKernelEntry table[] = {LIST(ENTRY)};                                    \
    *n = static_cast<int>(sizeof(table) / sizeof(table[0]));                                 \
    return table;                                                                            \
  }
#define STCD_DEFINE_CONV_TABLE(NAME, LIST) STCD_DEFINE_CONV_TABLE_WITH(NAME, LIST, STCD_CONV_TABLE_ENTRY)
#define STCD_CONV_TABLE_ENTRY(MT_, MS_, EPI_) {MT_, MS_, (EPI_), conv_ws_kernel<MT_, MS_, (EPI_)>, 4},
#define STCD_CONV_TABLE_ENTRY8(MT_, MS_, EPI_) {MT_, MS_, (EPI_), conv_ws_kernel<MT_, MS_, (EPI_), 8>, 8},

}  // namespace stcd
