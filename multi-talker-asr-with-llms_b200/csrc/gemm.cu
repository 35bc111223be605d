// tcgen05 / TMEM / TMA GEMM for sm_100a -- the single dense-contraction kernel of the hot path.
//
//   C[b][m][n] = epilogue(alpha * sum_k A[b][m][k] * B[b][n][k])      bf16 operands, fp32 accumulation in TMEM
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer  (one elected lane): global -> 128B-swizzled smem ring, 4 stages x (16 KB A + 32 KB B)
//   warp 1      MMA issuer    (one elected lane): tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x block_n x 16
//   warp 2      TMEM allocator (512 columns = 2 accumulator stages of 256 fp32 columns)
//   warps 4..7  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue); the second TMEM
// stage lets tile i+1's MMAs run under tile i's epilogue.
//
// Both operands may be K-major or MN-major (UMMA descriptor major bits), so forward (X W^T), input-gradient
// (dY W) and weight-gradient (dY^T X) contractions all run here without a transpose pass.  K-major A also takes
// an implicit-im2col addressing (tap, phase, row) so conv1d over a channels-last tensor is the same kernel.
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mtasr {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int A_STAGE_BYTES = BM * BK * 2;    // 16 KB; the B stage is block_n * BK * 2 bytes
static constexpr int CTRL_WARPS = 4;                 // TMA producer, MMA issuer, TMEM allocator, spare
static constexpr int EPI_WARPS = 8;                  // default: 2 warps per TMEM lane quadrant, alternating column groups
static constexpr int EPI_WARPS_MAX = 16;             // EW = 16 variant: 4 warps per quadrant (see epilogue_group16)
static constexpr int GEMM_THREADS = (CTRL_WARPS + EPI_WARPS) * 32;
static constexpr int STG_BYTES = 4096;               // one 32-row x 128-byte swizzled staging box per warp and tensor
static constexpr int MAX_STAGES = 8;
static constexpr int BAR_BYTES = 512;                // mbarriers + TMEM slot
static constexpr int BIAS_BYTES = EPI_WARPS * 64 * 4; // per-warp bias slot of one column group (EW warps: EW * 256 bytes)
static constexpr int SMEM_LIMIT = 232448;            // 227 KB of dynamic shared memory per CTA on sm_100
static constexpr int TMEM_COLS = 512;

struct GemmKP {
  int M, N, K, batch0, batch1;
  int a_major, b_major, block_n;
  int a_inner, a_phase;
  int a_use0, a_use1, b_use0, b_use1;  // 0 => operand broadcast over that batch dim (coordinate forced to 0)
  int m_tiles, n_tiles, num_tiles, num_kb;
  int n_fast;                          // tile rasterisation: 1 = consecutive tiles walk the N tiles of one M tile first
  int tail_start;                      // tiles >= tail_start are HALF tiles (block_n / 2 columns): see decode_tile; = num_tiles if unused
  int splits, kb_per_split;            // split-K (fp32 TMA reduce-add into a pre-zeroed C) for launches with few tiles
  int stages, stage_bytes;             // smem ring geometry (host-chosen to fit the staging buffers)
  int tma_epi;                         // 1: outputs leave through swizzled smem staging + TMA tiled stores
  int gw;                              // column-group width of the epilogue: 32 or 64
  int stg_aux, stg_res;                // staging buffer index of aux / residual (0 = C), -1 if absent
  int n_stg;
  void* c;
  int c_dtype;
  long long c_ld, c_sb0, c_sb1;
  __nv_bfloat16* aux;
  const float* bias;
  long long bias_sb0;
  const void* residual;
  int res_dtype;
  long long r_ld, r_sb0, r_sb1;
  int act;
  float alpha;
  int accumulate;
  int mode;
  const float* row_vec;
  const float* row_scale;
  float4* lse_part;
  DropP drop;                          // fused dropout (seed == nullptr: off)
  int dbg;                             // timing experiments only (MTASR_GEMM_DBG, WRONG results): 1 skip the staging store-read wait, 2 skip the bias load, 4 skip fence + TMA stores
};

__device__ __forceinline__ void load8(const void* base, int dtype, long long idx, int nv, float (&o)[8]) {
  if (dtype == MTASR_DT_BF16) {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + idx;
    if (nv == 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const uint4 u = *reinterpret_cast<const uint4*>(p);
      float2 t;
      t = unpack_bf16x2(u.x); o[0] = t.x; o[1] = t.y;
      t = unpack_bf16x2(u.y); o[2] = t.x; o[3] = t.y;
      t = unpack_bf16x2(u.z); o[4] = t.x; o[5] = t.y;
      t = unpack_bf16x2(u.w); o[6] = t.x; o[7] = t.y;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = j < nv ? bf2f(p[j]) : 0.f;
    }
  } else {
    const float* p = reinterpret_cast<const float*>(base) + idx;
    if (nv == 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const float4 u0 = *reinterpret_cast<const float4*>(p);
      const float4 u1 = *reinterpret_cast<const float4*>(p + 4);
      o[0] = u0.x; o[1] = u0.y; o[2] = u0.z; o[3] = u0.w;
      o[4] = u1.x; o[5] = u1.y; o[6] = u1.z; o[7] = u1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = j < nv ? p[j] : 0.f;
    }
  }
}

__device__ __forceinline__ void store8(void* base, int dtype, long long idx, int nv, const float (&v)[8]) {
  if (dtype == MTASR_DT_BF16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + idx;
    if (nv == 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      uint4 u;
      u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(p) = u;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nv) p[j] = f2bf(v[j]);
    }
  } else {
    float* p = reinterpret_cast<float*>(base) + idx;
    if (nv == 8 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nv) p[j] = v[j];
    }
  }
}

// SWIZZLE_128B address transform of a tiled TMA box in shared memory (buffer 1024-byte aligned):
// the 16-byte chunk index (address bits 4..6) is XORed with address bits 7..9.
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

// Boxes whose rows are 128 bytes (64 bf16 / 32 fp32 columns) use SWIZZLE_128B; the narrow 64-byte rows (bf16 tensors in
// a 32-column group, i.e. a bf16 tensor next to an fp32 one) are plain row-major (SWIZZLE_NONE).
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void stg_store8(uint8_t* buf, int dtype, int row, int gw, int c8, const float (&v)[8]) {
  if (dtype == MTASR_DT_F16) {
    uint4 u;
    u.x = pack_f16x2(v[0], v[1]); u.y = pack_f16x2(v[2], v[3]);
    u.z = pack_f16x2(v[4], v[5]); u.w = pack_f16x2(v[6], v[7]);
    const uint32_t off = static_cast<uint32_t>(row * gw * 2 + c8 * 16);
    *reinterpret_cast<uint4*>(buf + (gw == 64 ? swz(off) : off)) = u;
  } else if (dtype == MTASR_DT_BF16) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    const uint32_t off = static_cast<uint32_t>(row * gw * 2 + c8 * 16);
    *reinterpret_cast<uint4*>(buf + (gw == 64 ? swz(off) : off)) = u;
  } else {
    const uint32_t off = static_cast<uint32_t>(row * gw * 4 + c8 * 32);
    *reinterpret_cast<float4*>(buf + swz(off)) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(buf + swz(off + 16)) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

__device__ __forceinline__ void stg_load8(const uint8_t* buf, int dtype, int row, int gw, int c8, float (&o)[8]) {
  if (dtype == MTASR_DT_BF16) {
    const uint32_t boff = static_cast<uint32_t>(row * gw * 2 + c8 * 16);
    const uint4 u = *reinterpret_cast<const uint4*>(buf + (gw == 64 ? swz(boff) : boff));
    float2 t;
    t = unpack_bf16x2(u.x); o[0] = t.x; o[1] = t.y;
    t = unpack_bf16x2(u.y); o[2] = t.x; o[3] = t.y;
    t = unpack_bf16x2(u.z); o[4] = t.x; o[5] = t.y;
    t = unpack_bf16x2(u.w); o[6] = t.x; o[7] = t.y;
  } else {
    const uint32_t off = static_cast<uint32_t>(row * gw * 4 + c8 * 32);
    const float4 u0 = *reinterpret_cast<const float4*>(buf + swz(off));
    const float4 u1 = *reinterpret_cast<const float4*>(buf + swz(off + 16));
    o[0] = u0.x; o[1] = u0.y; o[2] = u0.z; o[3] = u0.w;
    o[4] = u1.x; o[5] = u1.y; o[6] = u1.z; o[7] = u1.w;
  }
}

struct TileCoord {
  int m_tile, n_tile, b0, b1, kb0, kb1;
  int n0, bn;   // first column and width of this tile (bn = block_n, or block_n / 2 for a tail half tile)
};
__device__ __forceinline__ TileCoord decode_tile(const GemmKP& p, int tile) {
  // Tiles that run at the same time (consecutive indices, one per SM or SM pair) should share the LARGER operand slab:
  // the fast index walks the dimension with fewer tiles, so a wave covers all of it and every slab of the other operand is
  // fetched from HBM once and served to its siblings from L2.  (With the M tiles always fastest, the vocabulary-sized
  // weight-gradient GEMMs -- 501 x 4 tiles -- streamed their 4.1 GB A operand once per N tile: 17 GB of DRAM reads for
  // 4.1 GB of data, profiles/gemm_traffic_r1_v15.csv.)
  // Tail splitting: the statically scheduled tiles run in rounds of one tile per SM (pair).  When the last round is less
  // than half full -- 252 tiles of 256 x 256 on 74 SM pairs are 3.4 rounds, paid as 4 -- its tiles are cut in two along N
  // (UMMA N = block_n / 2, same TMA boxes: the surplus B rows of the box are simply not read by the MMA), so the last round
  // costs half a tile time: 3.5 instead of 4 for the N = 1024 projections (out-proj, FFN2, every dgrad into D = 1024).
  TileCoord t;
  t.bn = p.block_n;
  int half = 0;
  if (tile >= p.tail_start) {
    const int h = tile - p.tail_start;
    tile = p.tail_start + (h >> 1);
    half = h & 1;
    t.bn = p.block_n >> 1;
  }
  int batch;
  if (p.n_fast) {
    t.n_tile = tile % p.n_tiles;
    const int rest = tile / p.n_tiles;
    t.m_tile = rest % p.m_tiles;
    batch = rest / p.m_tiles;
  } else {
    t.m_tile = tile % p.m_tiles;
    const int rest = tile / p.m_tiles;
    t.n_tile = rest % p.n_tiles;
    batch = rest / p.n_tiles;
  }
  if (p.splits > 1) {   // un-batched launch: the outermost index is the K split
    t.b0 = t.b1 = 0;
    t.kb0 = batch * p.kb_per_split;
    t.kb1 = min(p.num_kb, t.kb0 + p.kb_per_split);
  } else {
    t.b0 = batch % p.batch0;
    t.b1 = batch / p.batch0;
    t.kb0 = 0;
    t.kb1 = p.num_kb;
  }
  t.n0 = t.n_tile * p.block_n + half * t.bn;
  return t;
}

// ---- epilogue configuration, compile-time for the hot combinations (smaller, branch-free code: the fully generic
// kernel is ~17k SASS instructions and stalls on instruction fetch), run-time for everything else ----------------------
//   CFG = mode | act << 2 | aux << 5 | res << 6 | generic << 7
static constexpr int CFG_GENERIC = 128;
constexpr int make_cfg(int mode, int act, bool aux, bool res) { return mode | (act << 2) | (aux ? 32 : 0) | (res ? 64 : 0); }

template <int CFG>
struct Epi {
  static constexpr bool GEN = (CFG & CFG_GENERIC) != 0;
  __device__ __forceinline__ static int mode(const GemmKP& p) { return GEN ? p.mode : (CFG & 3); }
  __device__ __forceinline__ static int act(const GemmKP& p) { return GEN ? p.act : ((CFG >> 2) & 7); }
  __device__ __forceinline__ static bool aux(const GemmKP& p) { return GEN ? (p.aux != nullptr) : ((CFG & 32) != 0 && (CFG & 3) != 1); }
  // mode 1 + the aux bit: the logits tile is ALSO written (fp16) next to the LSE partials (CTC head forward for training)
  __device__ __forceinline__ static bool store1(const GemmKP& p) { return GEN ? (p.mode == 1 && p.c != nullptr) : ((CFG & 3) == 1 && (CFG & 32) != 0); }
  __device__ __forceinline__ static bool res(const GemmKP& p) { return GEN ? (p.residual != nullptr) : ((CFG & 64) != 0); }
  __device__ __forceinline__ static bool tma(const GemmKP& p) { return GEN ? (p.tma_epi != 0) : true; }
  __device__ __forceinline__ static bool res_tma(const GemmKP& p) { return GEN ? (p.tma_epi && p.stg_res >= 0) : ((CFG & 64) != 0); }
};

// One column group (GW = 32 or 64 accumulator columns) of one tile, for the 32 rows (TMEM lanes) of this warp.
template <int GW, int CFG>
__device__ __forceinline__ void epilogue_group(const GemmKP& p, const CUtensorMap* tmap_c, const CUtensorMap* tmap_aux,
                                               const CUtensorMap* tmap_r, uint8_t* stg, float* bias_s, uint64_t* res_bar,
                                               uint32_t& res_phase, uint32_t taddr, const TileCoord& t, int q, int lane,
                                               int g, const float* bias, float rvec, float rscale, bool row_ok,
                                               long long c_off, long long r_off, float& run_max, float& run_sum,
                                               int& run_idx) {
  using E = Epi<CFG>;
  const int mode = E::mode(p), act = E::act(p);
  const bool has_aux = E::aux(p), has_res = E::res(p), tma = E::tma(p), res_tma = E::res_tma(p);
  const bool store1 = E::store1(p);
  const bool tma_out = tma && (mode != 1 || store1);
  const int col0 = t.n0 + g * GW;
  const int row0 = t.m_tile * BM + q * 32;
  uint8_t* stg_c = stg;
  uint8_t* stg_aux = stg + (p.stg_aux > 0 ? p.stg_aux : 0) * STG_BYTES;
  uint8_t* stg_r = stg + (p.stg_res > 0 ? p.stg_res : 0) * STG_BYTES;
  if (tma_out) {
    // the previous group's bulk stores must have finished READING the staging buffers before they are rewritten
    if (lane == 0 && !(p.dbg & 1)) bulk_wait_read<0>();
    if (res_tma) fence_proxy_async();   // this warp's generic reads of the residual box precede its async overwrite
  }
  __syncwarp();
  if (res_tma && lane == 0) {
    mbar_arrive_expect_tx(res_bar, GW * 32 * (p.res_dtype == MTASR_DT_BF16 ? 2 : 4));
    tma_load_4d(stg_r, tmap_r, res_bar, col0, row0, t.b0, t.b1);
  }
  uint32_t r[GW];
#pragma unroll
  for (int j = 0; j < GW / 32; ++j) tmem_ld32(taddr + g * GW + j * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[j * 32]));
  // bias of this group's columns: one coalesced load per 32 columns into the warp's smem slot, overlapped with the
  // TMEM load; read back below as broadcast LDS.128 (no global-load latency inside the per-element loop)
  if (bias) {
#pragma unroll
    for (int j = 0; j < GW / 32; ++j) {
      const int col = col0 + j * 32 + lane;
      bias_s[j * 32 + lane] = (col < p.N && !(p.dbg & 2)) ? __ldg(bias + col) : 0.f;
    }
  }
  tmem_ld_wait();
  __syncwarp();
  if (res_tma) {
    mbar_wait(res_bar, res_phase);
    res_phase ^= 1;
  }
  const bool scale = p.alpha != 1.f;

  if (mode == 1) {
    // running (max, sum exp, first argmax) over this warp's columns of the row
#pragma unroll
    for (int c8 = 0; c8 < GW / 8; ++c8) {
      const int col = col0 + c8 * 8;
      if (col >= p.N) break;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c8 * 8 + j]);
      if (scale) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= p.alpha;
      }
      if (bias) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c8 * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c8 * 8 + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (store1) stg_store8(stg_c, MTASR_DT_F16, lane, GW, c8, v);   // columns >= N are clipped by the TMA store
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (col + j >= p.N) v[j] = -INFINITY;
        cm = fmaxf(cm, v[j]);
      }
      if (cm > run_max) {
        run_sum *= __expf(run_max - cm);
#pragma unroll
        for (int j = 7; j >= 0; --j)
          if (v[j] == cm) run_idx = col + j;
        run_max = cm;
      }
      const float nm = -run_max * 1.4426950408889634f;
#pragma unroll
      for (int j = 0; j < 8; ++j) run_sum += ex2_approx(fmaf(v[j], 1.4426950408889634f, nm));
    }
    if (!store1) return;
  }

  const float nrv = -rvec * 1.4426950408889634f;
  uint32_t drop_s0 = 0, drop_s1 = 0;
  if (p.drop.seed != nullptr) {
    drop_s0 = __ldg(p.drop.seed);
    drop_s1 = __ldg(p.drop.seed + 1);
  }
#pragma unroll
  for (int c8 = 0; c8 < GW / 8; ++c8) {
    if (mode == 1) break;
    const int col = col0 + c8 * 8;
    if (col >= p.N) break;   // warp-uniform
    const int nv = min(8, p.N - col);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c8 * 8 + j]);
    if (scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= p.alpha;
    }
    if (bias) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c8 * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c8 * 8 + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (mode == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ex2_approx(fmaf(v[j], 1.4426950408889634f, nrv)) * rscale;
    } else {
      if (has_aux) {
        if (tma) stg_store8(stg_aux, MTASR_DT_BF16, lane, GW, c8, v);
        else if (row_ok) store8(p.aux, MTASR_DT_BF16, c_off + col, nv, v);
      }
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_fast_f(v[j]);
      } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.drop.seed != nullptr) {
        // dropout of this GEMM's output (after bias / activation, before the residual add); for the backward forms (act 3 /
        // 4) the same factor multiplies the gradient next to the activation derivative
        float dm[8];
        const unsigned long long grow = static_cast<unsigned long long>(t.b1) * p.batch0 + t.b0;
        drop_mult8(drop_s0, drop_s1, p.drop, (grow * p.M + row0 + lane) * p.drop.ld + col, dm);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= dm[j];
      }
      if (has_res) {
        float rr[8];
        if (res_tma) stg_load8(stg_r, p.res_dtype, lane, GW, c8, rr);
        else if (row_ok) load8(p.residual, p.res_dtype, r_off + col, nv, rr);
        else {
#pragma unroll
          for (int j = 0; j < 8; ++j) rr[j] = 0.f;
        }
        if (act == 3) {          // backward through GELU: residual holds the saved pre-activation
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= gelu_grad_fast_f(rr[j]);
        } else if (act == 4) {   // backward through ReLU: residual holds the saved activation output
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = rr[j] > 0.f ? v[j] : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += rr[j];
        }
      }
      if (E::GEN && p.accumulate && row_ok) {
        float cc[8];
        load8(p.c, p.c_dtype, c_off + col, nv, cc);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += cc[j];
      }
    }
    if (tma) stg_store8(stg_c, p.c_dtype, lane, GW, c8, v);
    else if (row_ok) store8(p.c, p.c_dtype, c_off + col, nv, v);
  }
  if (tma_out && !(p.dbg & 4)) {
    fence_proxy_async();   // generic-proxy smem writes -> visible to the TMA (async proxy)
    __syncwarp();
    if (lane == 0) {
      if (p.splits > 1) tma_reduce_add_4d(tmap_c, stg_c, col0, row0, t.b0, t.b1);
      else tma_store_4d(tmap_c, stg_c, col0, row0, t.b0, t.b1);
      if (has_aux) tma_store_4d(tmap_aux, stg_aux, col0, row0, t.b0, t.b1);
      bulk_commit();
    }
  }
}

// ---- EW = 16 epilogue ------------------------------------------------------------------------------------------------
// The 8-warp epilogue is LATENCY-bound, not issue-bound: ncu on the FFN1 shape (GELU + pre-activation tap, K = 1024) shows
// 38 % issue-slot utilisation with the two epilogue warps of each scheduler at ~0.2 IPC while the tensor pipe idles 51 % of
// the time waiting for an accumulator stage (profiles/prof_gemm_gelu_aux_r2_lines.txt).  With 16 epilogue warps every
// scheduler holds four of them and each warp owns ONE 64-column group (bf16 outputs) of a 256-wide tile instead of two.
// 640 threads leave 102 registers per thread and 16 x 4 KB of staging must not cost pipeline stages, so this variant
//   * reads the accumulator in 32-column halves (32 live accumulator registers instead of 64),
//   * keeps ONE staging box per warp: a TMA-loaded residual / activation-backward source of the SAME dtype as the output is
//     combined in place (thread-private 16-byte pieces), and the pre-activation tap is written and stored first, the
//     activated output second, through the same box (one extra store-read wait per group, hidden by the other warps).
// Mode 0 only; launches whose residual has another dtype than the output, or that carry tap AND residual, use EW = 8.
template <int GW, int CFG>
__device__ __forceinline__ void epilogue_group16(const GemmKP& p, const CUtensorMap* tmap_c, const CUtensorMap* tmap_aux,
                                                 const CUtensorMap* tmap_r, uint8_t* stg, float* bias_s, uint64_t* res_bar,
                                                 uint32_t& res_phase, uint32_t taddr, const TileCoord& t, int q, int lane, int g,
                                                 const float* bias) {
  using E = Epi<CFG>;
  const int act = E::act(p);
  const bool has_aux = E::aux(p), has_res = E::res(p);
  const int col0 = t.n0 + g * GW;
  const int row0 = t.m_tile * BM + q * 32;
  const bool scale = p.alpha != 1.f;
  uint32_t drop_s0 = 0, drop_s1 = 0;
  if (p.drop.seed != nullptr) {
    drop_s0 = __ldg(p.drop.seed);
    drop_s1 = __ldg(p.drop.seed + 1);
  }
  // the previous group's bulk store must have finished READING the box before anything overwrites it
  if (lane == 0) bulk_wait_read<0>();
  if (has_res) fence_proxy_async();
  __syncwarp();
  if (bias) {
#pragma unroll
    for (int j = 0; j < GW / 32; ++j) {
      const int col = col0 + j * 32 + lane;
      bias_s[j * 32 + lane] = col < p.N ? __ldg(bias + col) : 0.f;
    }
  }
  __syncwarp();
  if (has_aux) {   // pass A: the pre-activation tap (bf16) leaves through the box first
#pragma unroll
    for (int hh = 0; hh < GW / 32; ++hh) {
      uint32_t r[32];
      tmem_ld32(taddr + g * GW + hh * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const int col = col0 + hh * 32 + c8 * 8;
        if (col >= p.N) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c8 * 8 + j]);
        if (scale) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= p.alpha;
        }
        if (bias) {
          const float4 b0 = *reinterpret_cast<const float4*>(bias_s + hh * 32 + c8 * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(bias_s + hh * 32 + c8 * 8 + 4);
          v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
          v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        stg_store8(stg, MTASR_DT_BF16, lane, GW, hh * 4 + c8, v);
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(tmap_aux, stg, col0, row0, t.b0, t.b1);
      bulk_commit();
      bulk_wait_read<0>();
    }
    __syncwarp();
  }
  if (has_res && lane == 0) {   // residual / activation-backward source -> the SAME box (same dtype as the output)
    mbar_arrive_expect_tx(res_bar, GW * 32 * (p.res_dtype == MTASR_DT_BF16 ? 2 : 4));
    tma_load_4d(stg, tmap_r, res_bar, col0, row0, t.b0, t.b1);
  }
#pragma unroll
  for (int hh = 0; hh < GW / 32; ++hh) {
    uint32_t r[32];
    tmem_ld32(taddr + g * GW + hh * 32, r);
    tmem_ld_wait();
    if (has_res && hh == 0) {
      mbar_wait(res_bar, res_phase);
      res_phase ^= 1;
    }
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      const int col = col0 + hh * 32 + c8 * 8;
      if (col >= p.N) break;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[c8 * 8 + j]);
      if (scale) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= p.alpha;
      }
      if (bias) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + hh * 32 + c8 * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + hh * 32 + c8 * 8 + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = gelu_fast_f(v[j]);
      } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.drop.seed != nullptr) {
        float dm[8];
        const unsigned long long grow = static_cast<unsigned long long>(t.b1) * p.batch0 + t.b0;
        drop_mult8(drop_s0, drop_s1, p.drop, (grow * p.M + row0 + lane) * p.drop.ld + col, dm);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= dm[j];
      }
      if (has_res) {
        float rr[8];
        stg_load8(stg, p.res_dtype, lane, GW, hh * 4 + c8, rr);   // this thread's own piece: overwritten in place below
        if (act == 3) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= gelu_grad_fast_f(rr[j]);
        } else if (act == 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = rr[j] > 0.f ? v[j] : 0.f;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += rr[j];
        }
      }
      stg_store8(stg, p.c_dtype, lane, GW, hh * 4 + c8, v);
    }
  }
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    if (p.splits > 1) tma_reduce_add_4d(tmap_c, stg, col0, row0, t.b0, t.b1);
    else tma_store_4d(tmap_c, stg, col0, row0, t.b0, t.b1);
    bulk_commit();
  }
}

// NCTA = 1: one CTA per 128 x block_n tile (tcgen05 cta_group::1).
// NCTA = 2: a CTA PAIR (cluster of 2, same TPC) per 256 x block_n tile (cta_group::2, UMMA M = 256): each CTA loads its own
// 128 rows of A and HALF of the B tile, the leader's MMA thread issues for both and each CTA's TMEM receives its 128 x
// block_n accumulator -- per-CTA shared-memory operand traffic drops from 48 KB to 32 KB per 64-wide k-block, which is
// what kept the single-CTA kernel (96 B/clk of operand reads next to the epilogue staging traffic) off the tensor peak.
template <int CFG, int NCTA, int EW>
__global__ void __launch_bounds__((CTRL_WARPS + EW) * 32, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_aux,
                 const __grid_constant__ CUtensorMap tmap_r, const GemmKP p) {
  using E = Epi<CFG>;
  // SWIZZLE_128B atoms need 1024-byte alignment.  The kernel has no static shared memory, so the dynamic window starts
  // 1024-aligned; this is checked (trap) rather than paid for with a 1 KB slack that would cost a pipeline stage.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* stg_base = smem + p.stages * p.stage_bytes;                       // 1024-aligned (stage_bytes % 1024 == 0)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + EW * p.n_stg * STG_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + EPI_WARPS_MAX);
  float* bias_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + BAR_BYTES);   // EPI_WARPS x 64 floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = NCTA == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = cta_rank == 0;
  const int first_tile = blockIdx.x / NCTA, tile_step = gridDim.x / NCTA;   // tiles are per CTA (pair)
  const int b_rows = p.block_n / NCTA;                                       // B-tile rows (N columns) held by this CTA

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&tmap_a);
      prefetch_tmap(&tmap_b);
      if (E::tma(p)) {
        prefetch_tmap(&tmap_c);
        if (E::aux(p)) prefetch_tmap(&tmap_aux);
        if (E::res_tma(p)) prefetch_tmap(&tmap_r);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&full_bar[i], NCTA);         // one producer arrival per CTA of the pair (on the leader's copy)
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], EW * NCTA);  // one arrive per epilogue warp of every CTA of the pair
      }
      for (int i = 0; i < EW; ++i) mbar_init(&res_bar[i], 1);
      fence_barrier_init();
    }
  } else if (warp == 2) {
    if (NCTA == 2) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive / TMA signal
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above is independent of the predecessor kernel and overlaps its tail (PDL, common.cuh); nothing below may
  // run before the predecessor grid has completed: operands, bias, residual and the pre-zeroed split-K output are its results
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = NCTA * (A_STAGE_BYTES + b_rows * BK * 2);   // bytes of BOTH CTAs land on the leader's barrier
      for (int tile = first_tile; tile < p.num_tiles; tile += tile_step) {
        TileCoord t = decode_tile(p, tile);
        t.m_tile = t.m_tile * NCTA + cta_rank;
        const int ab0 = p.a_use0 ? t.b0 : 0, ab1 = p.a_use1 ? t.b1 : 0;
        const int bb0 = p.b_use0 ? t.b0 : 0, bb1 = p.b_use1 ? t.b1 : 0;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          uint64_t* bar = &full_bar[stage];
          if (leader) mbar_arrive_expect_tx(bar, stage_tx);
          const int n0 = t.n0 + cta_rank * (t.bn / NCTA);   // half tiles: the box still loads b_rows rows, the MMA reads t.bn / NCTA
          if (NCTA == 2) {
            if (p.a_major == 0) {
              const int kk = kb * BK;
              const int tap = kk / p.a_inner;
              const int c0 = kk - tap * p.a_inner;
              const int ph = tap % p.a_phase;
              const int dr = tap / p.a_phase;
              tma_load_5d_pair(sa, &tmap_a, bar, c0, ph, t.m_tile * BM + dr, ab0, ab1);
            } else {
              tma_load_4d_pair(sa, &tmap_a, bar, t.m_tile * BM, kb * BK, ab0, ab1);
              tma_load_4d_pair(sa + 8192, &tmap_a, bar, t.m_tile * BM + 64, kb * BK, ab0, ab1);
            }
            if (p.b_major == 0) {
              tma_load_4d_pair(sb, &tmap_b, bar, kb * BK, n0, bb0, bb1);
            } else {
              for (int j = 0; j < b_rows / 64; ++j) tma_load_4d_pair(sb + j * 8192, &tmap_b, bar, n0 + j * 64, kb * BK, bb0, bb1);
            }
            if (!leader) mbar_arrive_leader(bar);
          } else {
            if (p.a_major == 0) {
              const int kk = kb * BK;
              const int tap = kk / p.a_inner;
              const int c0 = kk - tap * p.a_inner;
              const int ph = tap % p.a_phase;
              const int dr = tap / p.a_phase;
              tma_load_5d(sa, &tmap_a, bar, c0, ph, t.m_tile * BM + dr, ab0, ab1);
            } else {
              tma_load_4d(sa, &tmap_a, bar, t.m_tile * BM, kb * BK, ab0, ab1);
              tma_load_4d(sa + 8192, &tmap_a, bar, t.m_tile * BM + 64, kb * BK, ab0, ab1);
            }
            if (p.b_major == 0) {
              tma_load_4d(sb, &tmap_b, bar, kb * BK, n0, bb0, bb1);
            } else {
              for (int j = 0; j < b_rows / 64; ++j) tma_load_4d(sb + j * 8192, &tmap_b, bar, n0 + j * 64, kb * BK, bb0, bb1);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only for a pair)
    if (leader && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc_full = make_idesc_bf16(BM * NCTA, p.block_n, p.a_major, p.b_major);
      const uint32_t idesc_half = make_idesc_bf16(BM * NCTA, p.block_n / 2, p.a_major, p.b_major);
      for (int tile = first_tile; tile < p.num_tiles; tile += tile_step) {
        const TileCoord t = decode_tile(p, tile);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const uint32_t idesc = t.bn == p.block_n ? idesc_full : idesc_half;
        for (int kb = t.kb0; kb < t.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major SW128: +32 B per UMMA_K inside the 128 B swizzle row; 8-row groups 1024 B apart (SBO).
            // MN-major SW128: 64(mn) x 8(k) atoms of 1024 B; next 8 k = +1024 B (SBO), next 64 mn = +8192 B (LBO).
            const uint64_t da = p.a_major == 0 ? make_smem_desc(sa + k * 32, 16, 1024)
                                               : make_smem_desc(sa + k * 2048, 8192, 1024);
            const uint64_t db = p.b_major == 0 ? make_smem_desc(sb + k * 32, 16, 1024)
                                               : make_smem_desc(sb + k * 2048, 8192, 1024);
            if (NCTA == 2) umma_f16_pair(d_tmem, da, db, idesc, (kb != t.kb0 || k != 0) ? 1u : 0u);
            else umma_f16(d_tmem, da, db, idesc, (kb != t.kb0 || k != 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (NCTA == 2) umma_commit_pair(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (NCTA == 2) umma_commit_pair(&tmem_full[acc]);
        else umma_commit(&tmem_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= CTRL_WARPS) {
    // ------------------------------------------------------------------ epilogue: 8 warps, warp (q, half) owns TMEM lanes
    // 32q..32q+31 and the column groups g with (g & 1) == half
    const int e = warp - CTRL_WARPS;
    const int q = e & 3;
    const int half = e >> 2;               // column-group slot of this warp: 0..EW/4-1
    uint8_t* stg = stg_base + e * p.n_stg * STG_BYTES;
    float* bias_s = bias_smem + e * 64;
    const int mode = E::mode(p);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;
    for (int tile = first_tile; tile < p.num_tiles; tile += tile_step) {
      TileCoord t = decode_tile(p, tile);
      t.m_tile = t.m_tile * NCTA + cta_rank;
      const long long batch = static_cast<long long>(t.b1) * p.batch0 + t.b0;
      const int row = t.m_tile * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const int col_base = t.n0;
      const int ncols = min(t.bn, p.N - col_base);
      const int n_groups = (ncols + p.gw - 1) / p.gw;
      const long long c_off = t.b0 * p.c_sb0 + t.b1 * p.c_sb1 + static_cast<long long>(row) * p.c_ld;
      const long long r_off = t.b0 * p.r_sb0 + t.b1 * p.r_sb1 + static_cast<long long>(row) * p.r_ld;
      const float* bias = p.bias ? p.bias + t.b0 * p.bias_sb0 : nullptr;
      float rvec = 0.f, rscale = 1.f;
      if (mode == 2 && row_ok) {
        rvec = p.row_vec[batch * p.M + row];
        rscale = p.row_scale ? p.row_scale[batch * p.M + row] : 1.f;
      }
      float run_max = -INFINITY, run_sum = 0.f;
      int run_idx = 0x7fffffff;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      if constexpr (EW == 16) {
        for (int g = half; g < n_groups; g += 4) {
          if (p.gw == 64)
            epilogue_group16<64, CFG>(p, &tmap_c, &tmap_aux, &tmap_r, stg, bias_s, &res_bar[e], res_phase, taddr, t, q, lane, g, bias);
          else
            epilogue_group16<32, CFG>(p, &tmap_c, &tmap_aux, &tmap_r, stg, bias_s, &res_bar[e], res_phase, taddr, t, q, lane, g, bias);
        }
      } else
      for (int g = half; g < n_groups; g += 2) {
        if (p.gw == 64)
          epilogue_group<64, CFG>(p, &tmap_c, &tmap_aux, &tmap_r, stg, bias_s, &res_bar[e], res_phase, taddr, t, q, lane, g,
                                  bias, rvec, rscale, row_ok, c_off, r_off, run_max, run_sum, run_idx);
        else
          epilogue_group<32, CFG>(p, &tmap_c, &tmap_aux, &tmap_r, stg, bias_s, &res_bar[e], res_phase, taddr, t, q, lane, g,
                                  bias, rvec, rscale, row_ok, c_off, r_off, run_max, run_sum, run_idx);
      }
      // all TMEM reads of this tile are done (tcgen05.wait::ld inside every group): hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (NCTA == 2) mbar_arrive_leader(&tmem_empty[acc]);   // the leader's MMA thread waits for both CTAs' epilogues
        else mbar_arrive(&tmem_empty[acc]);
      }
      if (mode == 1 && row_ok) {
        p.lse_part[((batch * p.M + row) * p.n_tiles + t.n_tile) * 2 + half] =
            make_float4(run_max, run_sum, __int_as_float(run_idx), 0.f);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (E::tma(p) && lane == 0) bulk_wait_all<0>();   // outstanding bulk stores must complete before the CTA exits
  }

  tc_fence_before();
  if (NCTA == 2) cluster_sync_all();   // neither CTA may free TMEM / exit while the pair's MMAs or remote arrives are in flight
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (NCTA == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef void (*GemmKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const GemmKP);

struct KernelEntry {
  int cfg;
  GemmKernelFn fn;      // one CTA per tile
  GemmKernelFn fn2;     // CTA pair per 256-row tile
  GemmKernelFn fn16;    // 16 epilogue warps (mode 0 only; nullptr otherwise)
  GemmKernelFn fn16_2;
};
#define MTASR_GEMM_CFG(mode, act, aux, res) \
  { make_cfg(mode, act, aux, res), gemm_bf16_kernel<make_cfg(mode, act, aux, res), 1, 8>, gemm_bf16_kernel<make_cfg(mode, act, aux, res), 2, 8>, \
    nullptr, nullptr }
#define MTASR_GEMM_CFG16(mode, act, aux, res) \
  { make_cfg(mode, act, aux, res), gemm_bf16_kernel<make_cfg(mode, act, aux, res), 1, 8>, gemm_bf16_kernel<make_cfg(mode, act, aux, res), 2, 8>, \
    gemm_bf16_kernel<make_cfg(mode, act, aux, res), 1, 16>, gemm_bf16_kernel<make_cfg(mode, act, aux, res), 2, 16> }
// the epilogue combinations the model issues (TMA-staged outputs); anything else runs the generic kernel
static const KernelEntry kKernels[] = {
    // (measured at M = 15968, N = 4096, K = 1024, TFLOP/s with 8 -> 16 epilogue warps: plain bf16 1234 -> 1147, + residual
    //  1180 -> 1109, GELU + tap 976 -> 970, GELU backward 856 -> 962: the 16-warp variant pays for its single staging box with
    //  an extra store-read wait per group and for its 64 KB of staging with two pipeline stages, and only wins where the
    //  per-element work is longest -- it is enabled for the GELU-backward epilogue only)
    MTASR_GEMM_CFG(0, 0, false, false),    // plain (+bias): QKV, every dgrad / wgrad, attention contractions, conv FE
    MTASR_GEMM_CFG(0, 0, false, true),     // + residual: out-proj, FFN2, pos-conv dgrad
    MTASR_GEMM_CFG(0, 1, false, false),    // GELU: conv FE (group-norm variant)
    MTASR_GEMM_CFG(0, 1, true, false),     // GELU + pre-activation tap: FFN1
    MTASR_GEMM_CFG(0, 1, true, true),      // GELU + tap + residual: pos-conv
    MTASR_GEMM_CFG(0, 2, false, false),    // ReLU
    MTASR_GEMM_CFG(0, 2, true, false),     // ReLU + tap: separator projections
    MTASR_GEMM_CFG16(0, 3, false, true),   // GELU backward: FFN2 dgrad
    MTASR_GEMM_CFG(0, 4, false, true),     // ReLU backward
    MTASR_GEMM_CFG(1, 0, false, false),  // row LSE / argmax partials: greedy argmax, CTC head forward without autograd
    MTASR_GEMM_CFG(1, 0, true, false),   // row LSE partials + fp16 logits tile: CTC head forward for training
    MTASR_GEMM_CFG(2, 0, false, false),  // softmax regeneration: CTC head backward
    {CFG_GENERIC, gemm_bf16_kernel<CFG_GENERIC, 1, 8>, gemm_bf16_kernel<CFG_GENERIC, 2, 8>, nullptr, nullptr},
};

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                      const uint32_t* box, const char* what, int dtype = MTASR_DT_BF16, bool quiet = false,
                      bool swizzle = true) {
  const uint64_t esz = dtype == MTASR_DT_BF16 ? 2 : 4;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(MTASR_ERR_DRIVER, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i] ? dims[i] : 1;
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 1; i < rank; ++i) {
    gstr[i - 1] = strides_elems[i] * esz;
    if (gstr[i - 1] % 16 != 0 || gstr[i - 1] == 0) {
      if (quiet) return MTASR_ERR_INVALID_ARG;
      return set_error(MTASR_ERR_INVALID_ARG, "gemm: %s stride[%d]=%llu bytes is not a positive multiple of 16", what, i,
                       (unsigned long long)gstr[i - 1]);
    }
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) {
    if (quiet) return MTASR_ERR_INVALID_ARG;
    return set_error(MTASR_ERR_INVALID_ARG, "gemm: %s base pointer not 16-byte aligned", what);
  }
  CUresult r = fn(map, dtype == MTASR_DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                  const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (quiet) return MTASR_ERR_DRIVER;
    return set_error(MTASR_ERR_DRIVER,
                     "gemm: cuTensorMapEncodeTiled(%s) failed with %d (rank %d dims %llu,%llu,%llu,%llu strides %llu,%llu,%llu box %u,%u,%u)",
                     what, (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1],
                     (unsigned long long)gdim[2], (unsigned long long)(rank > 3 ? gdim[3] : 0),
                     (unsigned long long)gstr[0], (unsigned long long)gstr[1], (unsigned long long)(rank > 3 ? gstr[2] : 0),
                     bx[0], bx[1], bx[2]);
  }
  return 0;
}

// Tensor map of an epilogue tensor x[b0][b1][m][n] (element strides ld / sb0 / sb1) with a (gw cols x 32 rows) box.
static int encode_out_map(CUtensorMap* map, const void* base, int dtype, int M, int N, int batch0, int batch1, long long ld,
                          long long sb0, long long sb1, int gw) {
  if (ld <= 0) return MTASR_ERR_INVALID_ARG;
  const uint64_t d2 = (batch0 > 1 && sb0 > 0) ? batch0 : 1, d3 = (batch1 > 1 && sb1 > 0) ? batch1 : 1;
  if ((batch0 > 1 && sb0 <= 0) || (batch1 > 1 && sb1 <= 0)) return MTASR_ERR_INVALID_ARG;
  const uint64_t dims[4] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M), d2, d3};
  const uint64_t str[4] = {1, static_cast<uint64_t>(ld), d2 > 1 ? static_cast<uint64_t>(sb0) : static_cast<uint64_t>(ld),
                           d3 > 1 ? static_cast<uint64_t>(sb1) : static_cast<uint64_t>(ld)};
  const uint32_t box[4] = {static_cast<uint32_t>(gw), 32, 1, 1};
  const bool sw128 = gw * (dtype == MTASR_DT_BF16 ? 2 : 4) == 128;
  return encode_map(map, base, 4, dims, str, box, "epilogue tensor", dtype, true, sw128);
}

std::atomic<long long> g_launches{0};

// Optional per-launch timing of the GEMM kernel (bench.py roofline): CUDA events recorded on the launching stream
// around every gemm_bf16_kernel launch while profiling is on.
struct ProfRec {
  cudaEvent_t e0, e1;
  double flops;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

}  // namespace mtasr

using namespace mtasr;

extern "C" int64_t mtasr_launch_count(void) { return g_launches.load(); }

// Number of mode-1 partial slots per row: two epilogue warps (column halves) per N tile.
extern "C" int mtasr_gemm_n_tiles(int32_t N, int32_t block_n) {
  if (block_n != 64 && block_n != 128 && block_n != 256) block_n = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  return 2 * ((N + block_n - 1) / block_n);
}

extern "C" int mtasr_gemm_bf16(const mtasr_gemm_desc* d, void* stream) {
  MTASR_CHECK_ARG(d != nullptr, "gemm: null descriptor");
  MTASR_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "gemm: M,N,K must be positive (got %d,%d,%d)", d->M, d->N, d->K);
  MTASR_CHECK_ARG(d->batch0 > 0 && d->batch1 > 0, "gemm: batch dims must be positive");
  MTASR_CHECK_ARG(d->a && d->b, "gemm: null operand");
  MTASR_CHECK_ARG(d->mode >= 0 && d->mode <= 2, "gemm: bad mode %d", d->mode);
  MTASR_CHECK_ARG(d->mode == 1 ? d->lse_part != nullptr : d->c != nullptr, "gemm: missing output buffer");
  MTASR_CHECK_ARG(d->mode == 1 ? (d->c == nullptr || d->c_dtype == MTASR_DT_F16) : (d->c_dtype == MTASR_DT_BF16 || d->c_dtype == MTASR_DT_F32),
                  "gemm: C must be bf16 / f32 (mode 0, 2) or f16 (optional logits output of mode 1), got dtype %d", d->c_dtype);
  MTASR_CHECK_ARG(d->mode != 1 || d->c == nullptr || (!d->accumulate && !d->aux && !d->residual), "gemm: mode 1 takes no aux / residual / accumulate");
  MTASR_CHECK_ARG(d->mode != 2 || d->row_vec != nullptr, "gemm: mode 2 needs row_vec");

  GemmKP p;
  p.M = d->M; p.N = d->N; p.K = d->K; p.batch0 = d->batch0; p.batch1 = d->batch1;
  p.a_major = d->a_major ? 1 : 0;
  p.b_major = d->b_major ? 1 : 0;
  int bn = d->block_n;
  if (bn != 64 && bn != 128 && bn != 256) bn = d->N <= 64 ? 64 : (d->N <= 128 ? 128 : 256);
  p.block_n = bn;
  p.a_inner = d->a_inner > 0 ? d->a_inner : d->K;
  p.a_phase = d->a_phase > 0 ? d->a_phase : 1;
  if (p.a_major == 0 && p.a_inner < d->K)
    MTASR_CHECK_ARG(p.a_inner % BK == 0 && d->K % p.a_inner == 0,
                    "gemm: implicit-conv A needs a_inner %% 64 == 0 and K %% a_inner == 0 (a_inner=%d K=%d)", p.a_inner, d->K);
  p.a_use0 = (d->a_sb0 != 0 && d->batch0 > 1);
  p.a_use1 = (d->a_sb1 != 0 && d->batch1 > 1);
  p.b_use0 = (d->b_sb0 != 0 && d->batch0 > 1);
  p.b_use1 = (d->b_sb1 != 0 && d->batch1 > 1);
  // CTA pairs (256-row tiles, cta_group::2) whenever the M extent gives at least two full pairs of row tiles and each CTA's
  // half of the B tile is a whole number of 64-wide swizzle chunks
  const int m_tiles1 = (d->M + BM - 1) / BM;
  const char* pair_env = getenv("MTASR_GEMM_PAIR");
  const bool pair_allowed = pair_env == nullptr || pair_env[0] != '0';
  const int ncta = (pair_allowed && bn >= 128 && m_tiles1 >= 4) ? 2 : 1;
  p.m_tiles = (m_tiles1 + ncta - 1) / ncta;          // tiles of 128 * ncta rows
  p.n_tiles = (d->N + bn - 1) / bn;
  const long long nt = static_cast<long long>(p.m_tiles) * p.n_tiles * d->batch0 * d->batch1;
  MTASR_CHECK_ARG(nt < (1LL << 31), "gemm: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.num_kb = (d->K + BK - 1) / BK;
  {
    const char* rz = getenv("MTASR_GEMM_RASTER");   // debugging / A-B switch: "m" or "n" forces the fast index
    p.n_fast = rz ? (rz[0] == 'n') : (p.n_tiles < p.m_tiles);
  }
  p.splits = 1;
  p.kb_per_split = p.num_kb;
  p.c = d->c; p.c_dtype = d->c_dtype; p.c_ld = d->c_ld; p.c_sb0 = d->c_sb0; p.c_sb1 = d->c_sb1;
  p.aux = reinterpret_cast<__nv_bfloat16*>(d->aux);
  p.bias = d->bias; p.bias_sb0 = d->bias_sb0;
  p.residual = d->residual; p.res_dtype = d->res_dtype; p.r_ld = d->r_ld; p.r_sb0 = d->r_sb0; p.r_sb1 = d->r_sb1;
  p.act = d->act; p.alpha = d->alpha; p.accumulate = d->accumulate; p.mode = d->mode;
  p.row_vec = d->row_vec; p.row_scale = d->row_scale;
  p.lse_part = reinterpret_cast<float4*>(d->lse_part);
  p.dbg = 0;
  if (const char* dbg_env = getenv("MTASR_GEMM_DBG")) p.dbg = atoi(dbg_env);
  p.drop.seed = nullptr;
  if (d->drop_seed != nullptr) {
    MTASR_CHECK_ARG(d->mode == 0 && d->drop_keep16 > 0 && d->drop_keep16 <= 65536, "gemm: fused dropout needs mode 0 and keep16 in (0, 65536]");
    p.drop.seed = reinterpret_cast<const uint32_t*>(d->drop_seed);
    p.drop.site = d->drop_site;
    p.drop.thresh = d->drop_keep16;
    p.drop.scale = 65536.0f / static_cast<float>(d->drop_keep16);
    p.drop.ld = (static_cast<long long>(d->N) + 1) & ~1LL;
  }

  CUtensorMap ma, mb;
  const uint64_t nb0a = p.a_use0 ? d->batch0 : 1, nb1a = p.a_use1 ? d->batch1 : 1;
  const uint64_t nb0b = p.b_use0 ? d->batch0 : 1, nb1b = p.b_use1 ? d->batch1 : 1;
  const uint64_t safe_a = static_cast<uint64_t>(d->a_ld), safe_b = static_cast<uint64_t>(d->b_ld);
  int rc;
  if (p.a_major == 0) {
    const uint64_t taps = d->K / p.a_inner;
    const uint64_t rows = d->a_rows > 0 ? static_cast<uint64_t>(d->a_rows)
                                        : static_cast<uint64_t>(d->M) + (taps - 1) / p.a_phase;
    const uint64_t dims[5] = {static_cast<uint64_t>(p.a_inner), static_cast<uint64_t>(p.a_phase), rows, nb0a, nb1a};
    const uint64_t str[5] = {1, p.a_phase > 1 ? static_cast<uint64_t>(p.a_inner) : safe_a, safe_a,
                             p.a_use0 ? static_cast<uint64_t>(d->a_sb0) : safe_a,
                             p.a_use1 ? static_cast<uint64_t>(d->a_sb1) : safe_a};
    const uint32_t box[5] = {BK, 1, BM, 1, 1};
    rc = encode_map(&ma, d->a, 5, dims, str, box, "A(k-major)");
  } else {
    const uint64_t rows = d->a_rows > 0 ? static_cast<uint64_t>(d->a_rows) : static_cast<uint64_t>(d->K);
    const uint64_t dims[4] = {static_cast<uint64_t>(d->M), rows, nb0a, nb1a};
    const uint64_t str[4] = {1, safe_a, p.a_use0 ? static_cast<uint64_t>(d->a_sb0) : safe_a,
                             p.a_use1 ? static_cast<uint64_t>(d->a_sb1) : safe_a};
    const uint32_t box[4] = {64, BK, 1, 1};
    rc = encode_map(&ma, d->a, 4, dims, str, box, "A(mn-major)");
  }
  if (rc) return rc;
  if (p.b_major == 0) {
    const uint64_t rows = d->b_rows > 0 ? static_cast<uint64_t>(d->b_rows) : static_cast<uint64_t>(d->N);
    const uint64_t dims[4] = {static_cast<uint64_t>(d->K), rows, nb0b, nb1b};
    const uint64_t str[4] = {1, safe_b, p.b_use0 ? static_cast<uint64_t>(d->b_sb0) : safe_b,
                             p.b_use1 ? static_cast<uint64_t>(d->b_sb1) : safe_b};
    const uint32_t box[4] = {BK, static_cast<uint32_t>(bn / ncta), 1, 1};   // a CTA of a pair loads half of the B tile
    rc = encode_map(&mb, d->b, 4, dims, str, box, "B(k-major)");
  } else {
    const uint64_t rows = d->b_rows > 0 ? static_cast<uint64_t>(d->b_rows) : static_cast<uint64_t>(d->K);
    const uint64_t dims[4] = {static_cast<uint64_t>(d->N), rows, nb0b, nb1b};
    const uint64_t str[4] = {1, safe_b, p.b_use0 ? static_cast<uint64_t>(d->b_sb0) : safe_b,
                             p.b_use1 ? static_cast<uint64_t>(d->b_sb1) : safe_b};
    const uint32_t box[4] = {64, BK, 1, 1};
    rc = encode_map(&mb, d->b, 4, dims, str, box, "B(mn-major)");
  }
  if (rc) return rc;

  // ---- epilogue path: swizzled smem staging + TMA tiled stores (and TMA-loaded residual) whenever the tensors can be
  // described by a tensor map (16-byte aligned base / strides); otherwise per-thread vector loads/stores.
  CUtensorMap mc, maux, mr;
  memset(&mc, 0, sizeof(mc));
  memset(&maux, 0, sizeof(maux));
  memset(&mr, 0, sizeof(mr));
  p.tma_epi = 0;
  p.stg_aux = p.stg_res = -1;
  p.n_stg = 0;
  const bool any_f32 = (d->mode != 1 && d->c_dtype == MTASR_DT_F32) || (d->residual && d->res_dtype == MTASR_DT_F32);
  p.gw = any_f32 ? 32 : 64;
  if (bn < p.gw) p.gw = bn;
  const bool store1 = d->mode == 1 && d->c != nullptr;
  if (store1 || (d->mode != 1 && !d->accumulate && getenv("MTASR_GEMM_DIRECT_EPILOGUE") == nullptr)) {
    bool ok = encode_out_map(&mc, d->c, store1 ? MTASR_DT_BF16 /* 2-byte elements */ : d->c_dtype, d->M, d->N, d->batch0, d->batch1, d->c_ld, d->c_sb0, d->c_sb1, p.gw) == 0;
    int n = 1;
    if (ok && d->aux) {
      ok = encode_out_map(&maux, d->aux, MTASR_DT_BF16, d->M, d->N, d->batch0, d->batch1, d->c_ld, d->c_sb0, d->c_sb1, p.gw) == 0;
      p.stg_aux = n++;
    }
    bool res_ok = false;
    if (ok && d->residual) {
      res_ok = encode_out_map(&mr, d->residual, d->res_dtype, d->M, d->N, d->batch0, d->batch1, d->r_ld, d->r_sb0, d->r_sb1,
                              p.gw) == 0;
      if (res_ok) p.stg_res = n++;
    }
    if (ok) {
      p.tma_epi = 1;
      p.n_stg = n;
    } else {
      p.stg_aux = p.stg_res = -1;
      if (store1) return set_error(MTASR_ERR_INVALID_ARG, "gemm: the mode-1 logits output needs a 16-byte aligned C with c_ld %% 8 == 0");
    }
  }
  // split-K: an un-batched plain fp32 GEMM that would leave more than half of the SMs idle (the K = B*T weight-gradient
  // contractions with small M x N) is cut along K; partial tiles are summed by TMA reduce-add into a zeroed C.
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p.tma_epi && d->mode == 0 && d->act == 0 && !d->aux && !d->residual && !d->bias && d->c_dtype == MTASR_DT_F32 &&
      d->batch0 == 1 && d->batch1 == 1 && d->alpha == 1.0f && p.num_kb >= 16 && d->drop_seed == nullptr &&
      getenv("MTASR_GEMM_NO_SPLITK") == nullptr) {
    // choose the split that best fills whole waves of SMs (wave quantisation: e.g. 500 tiles on 148 SMs run as 4 waves at
    // 84 % occupancy, 2 x 500 half-K items as 7 waves at 97 %), keeping >= 32 k-blocks (K >= 2048) per item
    const int sms = num_units(ncta);   // schedulable units: CTAs or CTA pairs
    auto eff = [&](int sp) {
      const long long items = static_cast<long long>(p.num_tiles) * sp;
      const long long waves = (items + sms - 1) / sms;
      return static_cast<double>(items) / static_cast<double>(waves * sms);
    };
    int best = 1;
    double best_eff = eff(1);
    const int max_split = p.num_tiles * 2 <= sms ? sms / p.num_tiles : 4;
    for (int sp = 2; sp <= max_split; ++sp) {
      if (p.num_kb / sp < (p.num_tiles * 2 <= sms ? 8 : 32)) break;
      const double e = eff(sp);
      if (e > best_eff + 0.04) { best = sp; best_eff = e; }
    }
    if (best > 1) {
      p.kb_per_split = (p.num_kb + best - 1) / best;
      p.splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
      p.num_tiles *= p.splits;
      if (cudaMemset2DAsync(d->c, static_cast<size_t>(d->c_ld) * 4, 0, static_cast<size_t>(d->N) * 4, d->M, st) != cudaSuccess)
        return set_error(MTASR_ERR_LAUNCH, "gemm: split-K memset failed");
    }
  }
  // tail splitting (decode_tile): only for un-split, un-batched-or-batched mode 0 / 2 launches whose N is a whole number of
  // full tiles and whose last round is at most half full
  p.tail_start = p.num_tiles;
  if (p.splits == 1 && d->mode != 1 && bn == 256 && d->N % bn == 0 && getenv("MTASR_GEMM_NO_TAILSPLIT") == nullptr) {
    const int units = num_units(ncta);
    const int rem = p.num_tiles % units;
    if (p.num_tiles > units && rem > 0 && 2 * rem <= units) {
      p.tail_start = p.num_tiles - rem;
      p.num_tiles += rem;               // every tile of the last round becomes two half tiles
    }
  }
  // specialised kernel when the outputs are TMA-staged (mode 1 has no tensor output) and the residual is TMA-loaded
  const KernelEntry* entry = &kKernels[sizeof(kKernels) / sizeof(kKernels[0]) - 1];
  const bool special_ok = getenv("MTASR_GEMM_GENERIC") == nullptr && !d->accumulate &&
                          (d->mode == 1 || (p.tma_epi && (!d->residual || p.stg_res >= 0)));
  if (special_ok) {
    const int want = make_cfg(d->mode, d->mode == 0 ? d->act : 0, (d->mode == 0 && d->aux != nullptr) || store1,
                              d->mode == 0 && d->residual != nullptr);
    for (const KernelEntry& k : kKernels)
      if (k.cfg == want) entry = &k;
  }
  // 16 epilogue warps (epilogue_group16): mode-0 configurations compiled for it, 256-wide tiles, one staging box per warp --
  // so a TMA-loaded residual must have the output's dtype, and tap + residual together stay on the 8-warp kernel
  const char* ew_env = getenv("MTASR_GEMM_EW");
  const bool ew16 = special_ok && entry->fn16 != nullptr && d->mode == 0 && bn == 256 && p.tma_epi &&
                    !(d->aux && d->residual) && (!d->residual || d->res_dtype == d->c_dtype) &&
                    (!d->aux || d->c_dtype == MTASR_DT_BF16) && !(ew_env && ew_env[0] == '8');
  const int ew = ew16 ? 16 : EPI_WARPS;
  if (ew16) {
    p.n_stg = 1;
    p.stg_aux = d->aux ? 0 : -1;
    p.stg_res = d->residual ? 0 : -1;
  }
  p.stage_bytes = A_STAGE_BYTES + (bn / ncta) * BK * 2;
  const int fixed = BAR_BYTES + ew * 64 * 4 + ew * p.n_stg * STG_BYTES;
  int stages = (SMEM_LIMIT - fixed) / p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return set_error(MTASR_ERR_UNSUPPORTED, "gemm: shared memory budget allows %d pipeline stages", stages);
  p.stages = stages;
  const int smem_bytes = fixed + stages * p.stage_bytes;

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    for (const KernelEntry& k : kKernels) {
      for (GemmKernelFn f : {k.fn, k.fn2, k.fn16, k.fn16_2}) {
        if (f == nullptr) continue;
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
        if (e != cudaSuccess) attr_err = e;
      }
    }
  });
  if (attr_err != cudaSuccess)
    return set_error(MTASR_ERR_LAUNCH, "gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
  GemmKernelFn kernel = ew16 ? (ncta == 2 ? entry->fn16_2 : entry->fn16) : (ncta == 2 ? entry->fn2 : entry->fn);
  const int threads = (CTRL_WARPS + ew) * 32;

  const int units = num_units(ncta);
  const int grid = (p.num_tiles < units ? p.num_tiles : units) * ncta;
  ProfRec rec{};
  bool prof = false;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof = g_prof_on;
  }
  if (prof) {
    cudaEventCreate(&rec.e0);
    cudaEventCreate(&rec.e1);
    rec.flops = 2.0 * d->M * static_cast<double>(d->N) * d->K * d->batch0 * d->batch1;
    cudaEventRecord(rec.e0, st);
  }
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (ncta == 2) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    if (pdl_enabled()) {   // the kernel's prologue ends in pdl_wait()
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kernel, ma, mb, mc, maux, mr, p);
    if (le != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "gemm: launch failed: %s", cudaGetErrorString(le));
  }
  g_launches.fetch_add(1);
  if (prof) {
    cudaEventRecord(rec.e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(rec);
  }
  MTASR_CHECK_LAUNCH("gemm_bf16");
  return MTASR_OK;
}

extern "C" int mtasr_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  g_prof_on = true;
  return MTASR_OK;
}

extern "C" int mtasr_profile_end(double* gemm_ms, double* gemm_flops, int64_t* gemm_launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  double ms = 0.0, fl = 0.0;
  for (auto& r : g_prof) {
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "profile_end: event sync failed");
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    ms += t;
    fl += r.flops;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (gemm_ms) *gemm_ms = ms;
  if (gemm_flops) *gemm_flops = fl;
  if (gemm_launches) *gemm_launches = static_cast<int64_t>(g_prof.size());
  g_prof.clear();
  return MTASR_OK;
}
