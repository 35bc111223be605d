// Fused self-attention with the gated relative-position bias folded into the softmax tile (hf:147-271,
// torch F.multi_head_attention_forward hf:206-228) on tcgen05 / TMEM / TMA.
//
//   O[b,q,h,:] = softmax_k( Q K^T / sqrt(d) + gate[b,h,q] * table[h, k-q+T-1]  (k < klen[b]) ) V        d = 64
//
// The reference materialises the (B*H,T,T) fp32 bias and SDPA's score tensors every layer; the unfused path of this
// repo wrote S (fp32) and P (bf16).  Here nothing of size T x T ever reaches HBM: one CTA owns a 128-query tile of one
// (utterance, head); S tiles live in TMEM, P tiles in swizzled shared memory, O accumulates in TMEM.
//
// Forward, warp-specialised (640 threads): warp 0 TMA producer (Q, then K/V tiles through a 4-slot ring), warp 1
// single-thread tcgen05.mma issuer, warp 2 TMEM allocator, warps 4..19 softmax (thread = query row x 64-key half; four
// warpgroups = 2 S/P buffers x 2 column halves).
//   * attn_fwd_sp_kernel (T <= 512, the headline shape): single pass -- every key tile has its own O accumulator in TMEM
//     and its own tile-local row maximum; the epilogue combines the (<= 4) partial results exactly.
//   * attn_fwd_kernel (any T): exact two-pass softmax -- pass A runs QK^T only and reduces the row maximum, pass B
//     recomputes the S tiles, forms P = exp(z - max) once and accumulates O += P V in TMEM without rescaling.
// Both are bound by the shared-memory pipe (one table look-up per score + P staging + MMA operand reads), not by MUFU or
// the tensor pipe: ncu line profiles under profiles/.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace mtasr {

static constexpr int AT_D = 64;          // head dim
static constexpr int AT_BQ = 128;        // queries per CTA tile
static constexpr int AT_BK = 128;        // keys per inner tile
static constexpr int AT_TILE = 128 * 64 * 2;   // one 128 x 64 bf16 operand tile = 16 KB
static constexpr int AT_KV_SLOTS = 4;
static constexpr int AT_P_BYTES = 128 * 128 * 2;   // P tile: two 64-key chunks of 128 rows x 128 B
static constexpr int AT_THREADS = 640;         // 4 control warps + 4 softmax warpgroups (2 S/P buffers x 2 column halves)
static constexpr float LOG2E = 1.4426950408889634f;

struct AttnFwdP {
  int B, H, T, nq, nk;            // nq / nk = number of 128-row query / key tiles
  int n_items;
  float scale_log2;               // softmax scale * log2(e)
  const float* gate;              // (B,H,T)
  const float* table;             // (H, 2T-1)
  const int* klen;                // (B) or null
  __nv_bfloat16* out;             // (B*T, H*64)
  float* lse;                     // (B,H,T) natural-log row LSE of the biased scores (saved for the backward)
  DropP drop;                     // dropout on the attention probabilities (hf:217); index ((b*H + h)*T + q) * drop.ld + k
};

// Dropout multipliers (0 or 1/keep) of the n consecutive keys kb .. kb+n-1 (kb even, n a multiple of 2) of probability row `row`.
// Work items are ordered (head, utterance, tile) and every CTA takes ONE CONTIGUOUS chunk of them: a CTA then stays on one
// head for (almost) its whole run -- the relative-position table (and, in the backward, its gradient accumulator) in shared
// memory is reloaded / flushed once or twice per launch instead of once per item (with the former round-robin order over
// (utterance, head, tile) the head changed on every item: 148 / nq is not a multiple of H), and consecutive items of a CTA
// re-use the K / V tiles of the same (utterance, head).
__device__ __forceinline__ int item_begin(int n_items) {
  const int base = n_items / static_cast<int>(gridDim.x), rem = n_items % static_cast<int>(gridDim.x);
  return static_cast<int>(blockIdx.x) * base + min(static_cast<int>(blockIdx.x), rem);
}
__device__ __forceinline__ int item_end(int n_items) {
  const int base = n_items / static_cast<int>(gridDim.x), rem = n_items % static_cast<int>(gridDim.x);
  return item_begin(n_items) + base + (static_cast<int>(blockIdx.x) < rem ? 1 : 0);
}

template <int N>
__device__ __forceinline__ void attn_drop_mults(uint32_t s0, uint32_t s1, const DropP& d, unsigned long long row, int kb, float (&m)[N]) {
  const unsigned long long pair0 = (row * static_cast<unsigned long long>(d.ld) + static_cast<unsigned long long>(kb)) >> 1;
#pragma unroll
  for (int j = 0; j < N / 2; ++j) {
    const uint32_t h = drop_hash(s0, s1, d.site, pair0 + j);
    m[2 * j] = (h & 0xffffu) < d.thresh ? d.scale : 0.f;
    m[2 * j + 1] = (h >> 16) < d.thresh ? d.scale : 0.f;
  }
}

// K-major SW128 operand tile [rows][64] bf16: UMMA_K step ks (16 elements) starts 32 bytes further in the swizzle row.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int ks) { return make_smem_desc(base + ks * 32, 16, 1024); }
// MN-major view of a [k-rows][64 mn] tile (rows are the contraction index): 16 k-rows = 2048 bytes per UMMA_K step;
// `lbo` = byte distance between 64-wide MN chunks.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int ks, uint32_t lbo) {
  return make_smem_desc(base + ks * 2048, lbo, 1024);
}
__device__ __forceinline__ uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

__host__ __device__ __forceinline__ int attn_table_bytes(int T) { return ((2 * T - 1 + 256) * 4 + 15) / 16 * 16; }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool DROP>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnFwdP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* q_s = smem;                                   // 16 KB
  uint8_t* kv_s = q_s + AT_TILE;                         // 4 x 16 KB
  uint8_t* p_s = kv_s + AT_KV_SLOTS * AT_TILE;           // 2 x 32 KB
  float* tbl_s = reinterpret_cast<float*>(p_s + 2 * AT_P_BYTES);   // 2T-1 floats
  // (+256 floats of slack: masked lanes of the last key tile may form addresses up to 128 entries past the table)
  float* xch_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tbl_s) + attn_table_bytes(p.T));   // 4 x 128 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch_s + 512);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;                  // [4]
  uint64_t* kv_empty = kv_full + AT_KV_SLOTS;    // [4]
  uint64_t* s_full = kv_empty + AT_KV_SLOTS;     // [2]
  uint64_t* s_empty = s_full + 2;                // [2]
  uint64_t* p_full = s_empty + 2;                // [2]
  uint64_t* p_empty = p_full + 2;                // [2]
  uint64_t* o_full = p_empty + 2;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (elect_one()) prefetch_tmap(&tmap_qkv);
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      mbar_init(q_empty, 1);
      for (int i = 0; i < AT_KV_SLOTS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
        mbar_init(&p_full[i], 8); mbar_init(&p_empty[i], 1);
      }
      mbar_init(o_full, 1);
      mbar_init(o_empty, 16);
      fence_barrier_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // PDL (common.cuh): the set-up above overlaps the predecessor's tail; q/k/v, gate and table are its results
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s[2] = {tmem_base, tmem_base + 128};
  const uint32_t tm_o = tmem_base + 256;
  const int nk = p.nk;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      int slot = 0;
      uint32_t ph = 0, qph = 0;
      for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
        const int qt = item % p.nq, b = (item / p.nq) % p.B, h = item / (p.nq * p.B);
        mbar_wait(q_empty, qph ^ 1);
        qph ^= 1;
        mbar_arrive_expect_tx(q_full, AT_TILE);
        tma_load_4d(q_s, &tmap_qkv, q_full, 0, qt * AT_BQ, h, b);
        auto load = [&](int row0, int slice) {
          mbar_wait(&kv_empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&kv_full[slot], AT_TILE);
          tma_load_4d(kv_s + slot * AT_TILE, &tmap_qkv, &kv_full[slot], 0, row0, slice, b);
          if (++slot == AT_KV_SLOTS) { slot = 0; ph ^= 1; }
        };
        for (int j = 0; j < nk; ++j) load(j * AT_BK, p.H + h);            // pass A: K tiles
        load(0, p.H + h);                                                // pass B: K_0, then K_{j+1}, V_j
        for (int j = 0; j < nk; ++j) {
          if (j + 1 < nk) load((j + 1) * AT_BK, p.H + h);
          load(j * AT_BK, 2 * p.H + h);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);          // S = Q K^T: both K-major
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);           // O += P V: A K-major, B (V) MN-major
      int slot = 0;
      uint32_t ph = 0, qph = 0, oph = 0;
      uint32_t sph[2] = {0, 0}, pph[2] = {0, 0};
      int sb = 0;   // S buffer / P buffer alternate per key tile, continuing across passes and items
      auto issue_s = [&](int buf) {
        mbar_wait(&kv_full[slot], ph);
        mbar_wait(&s_empty[buf], sph[buf] ^ 1);
        sph[buf] ^= 1;
        tc_fence_after();
        const uint32_t qa = smem_u32(q_s), ka = smem_u32(kv_s + slot * AT_TILE);
#pragma unroll
        for (int ks = 0; ks < AT_D / 16; ++ks) umma_f16(tm_s[buf], desc_kmajor(qa, ks), desc_kmajor(ka, ks), idesc_s, ks != 0);
        umma_commit(&kv_empty[slot]);
        umma_commit(&s_full[buf]);
        if (++slot == AT_KV_SLOTS) { slot = 0; ph ^= 1; }
      };
      for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
        mbar_wait(q_full, qph);
        qph ^= 1;
        for (int j = 0; j < nk; ++j) { issue_s(sb); sb ^= 1; }          // pass A
        int pb = sb;                                                     // pass B: P buffer index follows the S buffer
        issue_s(sb); sb ^= 1;
        for (int j = 0; j < nk; ++j) {
          if (j + 1 < nk) { issue_s(sb); sb ^= 1; }
          mbar_wait(&p_full[pb], pph[pb]);
          pph[pb] ^= 1;
          mbar_wait(&kv_full[slot], ph);
          if (j == 0) { mbar_wait(o_empty, oph ^ 1); oph ^= 1; }
          tc_fence_after();
          const uint32_t pa = smem_u32(p_s + pb * AT_P_BYTES), va = smem_u32(kv_s + slot * AT_TILE);
#pragma unroll
          for (int ks = 0; ks < AT_BK / 16; ++ks)
            umma_f16(tm_o, make_smem_desc(pa + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), desc_mnmajor(va, ks, 8192), idesc_o,
                     (j | ks) != 0);
          umma_commit(&kv_empty[slot]);
          umma_commit(&p_empty[pb]);
          if (++slot == AT_KV_SLOTS) { slot = 0; ph ^= 1; }
          pb ^= 1;
        }
        umma_commit(o_full);
        umma_commit(q_empty);
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax warps: thread = (query row, 64-key half).
    // Four warpgroups: group (grp, sub) owns key columns [64 sub, 64 sub + 64) of every key tile that lands in S / P buffer
    // grp, so consecutive tiles are exponentiated concurrently and every scheduler holds four softmax warps (the per-element
    // chain LDS -> FFMA -> MUFU -> pack is latency-bound: with two warps per scheduler the kernel ran at ~0.1 IPC per
    // warp).  The partial row max / row sum of the four groups are combined through shared memory.
    const int wq = warp & 3;
    const int grp = ((warp - 4) >> 2) & 1;
    const int sub = (warp - 4) >> 3;
    const int gid = sub * 2 + grp;                 // 0..3
    const int r = wq * 32 + lane;
    const int st = threadIdx.x - 128;              // 0..511 over the four groups
    uint32_t sph = 0, pph = 0, oph = 0;            // phases of this group's S / P buffer and of the O accumulator
    uint32_t drop_s0 = 0, drop_s1 = 0;
    if (DROP) { drop_s0 = __ldg(p.drop.seed); drop_s1 = __ldg(p.drop.seed + 1); }
    int cur_h = -1;
    for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
      const int qt = item % p.nq, b = (item / p.nq) % p.B, h = item / (p.nq * p.B);
      named_bar_sync(1, 512);                      // every softmax thread is done with the previous item's table / xch
      if (h != cur_h) {
        const float* trow = p.table + static_cast<long long>(h) * (2 * p.T - 1);
        for (int i = st; i < 2 * p.T - 1; i += 512) tbl_s[i] = trow[i] * LOG2E;
        cur_h = h;
      }
      named_bar_sync(1, 512);
      const int q = qt * AT_BQ + r;
      const bool q_ok = q < p.T;
      const int qc = q_ok ? q : p.T - 1;
      const int kl = p.klen ? min(p.klen[b], p.T) : p.T;
      const float g = p.gate[(static_cast<long long>(b) * p.H + h) * p.T + qc];
      const float* trel = tbl_s + (p.T - 1 - qc);   // trel[k] = log2e * table[h, k - q + T - 1]
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tm_mine = tm_s[grp] + lane_off + sub * 64;

      // pass A: exact row maximum of z (log2 units) over this group's half tiles (tile j lives in buffer j & 1)
      float m = -INFINITY;
      for (int j = grp; j < nk; j += 2) {
        mbar_wait(&s_full[grp], sph);
        sph ^= 1;
        tc_fence_after();
        const int k0 = j * AT_BK + sub * 64;
        uint32_t va[32], vb[32];
        tmem_ld32(tm_mine, va);
        tmem_ld32(tm_mine + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[grp]);   // both chunks are in registers: the buffer may be overwritten
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t(&cur)[32] = c ? vb : va;
          const int kb = k0 + c * 32;
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // independent chains (latency-bound warps)
          if (kb + 32 <= kl) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], fmaf(__uint_as_float(cur[i]), p.scale_log2, g * trel[kb + i]));
          } else if (kb < kl) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kb + i < kl) mx4[i & 3] = fmaxf(mx4[i & 3], fmaf(__uint_as_float(cur[i]), p.scale_log2, g * trel[kb + i]));
          }
          m = fmaxf(m, fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])));
        }
      }
      xch_s[gid * 128 + r] = m;
      named_bar_sync(1, 512);
      m = fmaxf(fmaxf(xch_s[r], xch_s[128 + r]), fmaxf(xch_s[256 + r], xch_s[384 + r]));
      const float mm = m == -INFINITY ? 0.f : m;
      named_bar_sync(1, 512);                      // all groups have read the maxima before xch is reused for the sums

      // pass B: P = exp2(z - m) -> bf16 smem tile, l = partial row sum (pass-B tile j lives in buffer (nk + j) & 1)
      float l = 0.f;
      for (int j = (grp + nk) & 1; j < nk; j += 2) {
        mbar_wait(&s_full[grp], sph);
        sph ^= 1;
        mbar_wait(&p_empty[grp], pph ^ 1);
        pph ^= 1;
        tc_fence_after();
        const int k0 = j * AT_BK + sub * 64;
        uint8_t* chunk = p_s + grp * AT_P_BYTES + sub * 16384;   // this group's 64-key swizzle chunk of the P tile
        uint32_t va[32], vb[32];
        tmem_ld32(tm_mine, va);
        tmem_ld32(tm_mine + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[grp]);   // the next QK^T may overwrite this S buffer
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t(&cur)[32] = c ? vb : va;
          const int kb = k0 + c * 32;
          float e[32];
          if (kb + 32 <= kl) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              e[i] = ex2_approx(fmaf(__uint_as_float(cur[i]), p.scale_log2, fmaf(g, trel[kb + i], -mm)));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              e[i] = kb + i < kl ? ex2_approx(fmaf(__uint_as_float(cur[i]), p.scale_log2, fmaf(g, trel[kb + i], -mm))) : 0.f;
          }
          {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 32; ++i) s4[i & 3] += e[i];
            l += (s4[0] + s4[1]) + (s4[2] + s4[3]);
          }
          if (DROP) {   // the row sum (softmax denominator) is over the UNDROPPED probabilities; P V uses the dropped ones
            float dm[32];
            attn_drop_mults<32>(drop_s0, drop_s1, p.drop, (static_cast<unsigned long long>(b) * p.H + h) * p.T + qc, kb, dm);
#pragma unroll
            for (int i = 0; i < 32; ++i) e[i] *= dm[i];
          }
#pragma unroll
          for (int g16 = 0; g16 < 4; ++g16) {
            uint4 u;
            u.x = pack_bf16x2(e[g16 * 8 + 0], e[g16 * 8 + 1]); u.y = pack_bf16x2(e[g16 * 8 + 2], e[g16 * 8 + 3]);
            u.z = pack_bf16x2(e[g16 * 8 + 4], e[g16 * 8 + 5]); u.w = pack_bf16x2(e[g16 * 8 + 6], e[g16 * 8 + 7]);
            *reinterpret_cast<uint4*>(chunk + swz128(static_cast<uint32_t>(r * 128 + c * 64 + g16 * 16))) = u;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[grp]);
      }
      xch_s[gid * 128 + r] = l;
      named_bar_sync(1, 512);
      l = (xch_s[r] + xch_s[128 + r]) + (xch_s[256 + r] + xch_s[384 + r]);

      // epilogue: O / l -> bf16 (group gid stores head-dim columns 16 gid .. 16 gid + 15), LSE
      mbar_wait(o_full, oph);
      oph ^= 1;
      tc_fence_after();
      uint32_t o0[16];
      tmem_ld16(tm_o + lane_off + gid * 16, o0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (q_ok) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.T + q) * (p.H * AT_D) + h * AT_D + gid * 16;
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o0[g16 * 8 + 0]) * inv, __uint_as_float(o0[g16 * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o0[g16 * 8 + 2]) * inv, __uint_as_float(o0[g16 * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o0[g16 * 8 + 4]) * inv, __uint_as_float(o0[g16 * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o0[g16 * 8 + 6]) * inv, __uint_as_float(o0[g16 * 8 + 7]) * inv);
          reinterpret_cast<uint4*>(orow)[g16] = u;
        }
        // natural-log LSE of the biased scores: z_log2 = z * log2e  =>  lse = (m + log2(l)) / log2e
        if (gid == 0)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.T + q] = l > 0.f ? (mm + log2f(l)) * 0.6931471805599453f : -INFINITY;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ forward, single pass
// T <= 512 (at most four 128-key tiles per row -- the 10 s utterances of the headline configuration): every key tile gets
// its OWN 64-column O accumulator in TMEM (2 S buffers + 4 accumulators = all 512 columns) and is exponentiated against
// its own tile-local row maximum m_j, so there is neither the max-only first pass of attn_fwd_kernel (a second QK^T per
// tile and twice the MMA <-> softmax hand-offs, which is what that kernel's time is made of) nor an online rescale of a
// running accumulator.  The epilogue combines exactly:  O = sum_j 2^(m_j - m) O_j / l,  l = sum_j 2^(m_j - m) l_j.
template <bool DROP>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_sp_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnFwdP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* q_s = smem;                                   // 2 x 16 KB: the next item's Q tile is prefetched
  uint8_t* kv_s = q_s + 2 * AT_TILE;                     // 4 x 16 KB
  uint8_t* p_s = kv_s + AT_KV_SLOTS * AT_TILE;           // 2 x 32 KB
  float* tbl_s = reinterpret_cast<float*>(p_s + 2 * AT_P_BYTES);   // 2T-1 floats (+ slack, see attn_table_bytes)
  float* xm_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tbl_s) + attn_table_bytes(p.T));   // [2 sets][2 grp][2 sub][128]
  float* tm_sm = xm_s + 1024;                            // [2 item parities][4 tiles][128]  tile-local row maxima
  float* tl_sm = tm_sm + 1024;                           // [2 item parities][4 tiles][2 halves][128]  partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(tl_sm + 2048);
  uint64_t* q_full = bars;                       // [2]
  uint64_t* q_empty = bars + 2;                  // [2]
  uint64_t* kv_full = bars + 4;                  // [4]
  uint64_t* kv_empty = kv_full + AT_KV_SLOTS;    // [4]
  uint64_t* s_full = kv_empty + AT_KV_SLOTS;     // [2]
  uint64_t* s_empty = s_full + 2;                // [2]
  uint64_t* p_full = s_empty + 2;                // [2]
  uint64_t* p_empty = p_full + 2;                // [2]
  uint64_t* o_full = p_empty + 2;
  uint64_t* o_empty = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (elect_one()) prefetch_tmap(&tmap_qkv);
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
      for (int i = 0; i < AT_KV_SLOTS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
        mbar_init(&p_full[i], 8); mbar_init(&p_empty[i], 1);
      }
      mbar_init(o_full, 1);
      mbar_init(o_empty, 16);
      fence_barrier_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // PDL (common.cuh): the set-up above overlaps the predecessor's tail; q/k/v, gate and table are its results
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s[2] = {tmem_base, tmem_base + 128};
  const uint32_t tm_o = tmem_base + 256;         // accumulator of key tile j: tm_o + 64 j
  const int nk = p.nk;                           // <= 4

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      int slot = 0, qb = 0;
      uint32_t ph = 0, qph[2] = {0, 0};
      for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
        const int qt = item % p.nq, b = (item / p.nq) % p.B, h = item / (p.nq * p.B);
        mbar_wait(&q_empty[qb], qph[qb] ^ 1);
        qph[qb] ^= 1;
        mbar_arrive_expect_tx(&q_full[qb], AT_TILE);
        tma_load_4d(q_s + qb * AT_TILE, &tmap_qkv, &q_full[qb], 0, qt * AT_BQ, h, b);
        qb ^= 1;
        auto load = [&](int row0, int slice) {
          mbar_wait(&kv_empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&kv_full[slot], AT_TILE);
          tma_load_4d(kv_s + slot * AT_TILE, &tmap_qkv, &kv_full[slot], 0, row0, slice, b);
          if (++slot == AT_KV_SLOTS) { slot = 0; ph ^= 1; }
        };
        load(0, p.H + h);                                                // K_0, then K_{j+1}, V_j
        for (int j = 0; j < nk; ++j) {
          if (j + 1 < nk) load((j + 1) * AT_BK, p.H + h);
          load(j * AT_BK, 2 * p.H + h);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);          // S = Q K^T: both K-major
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);           // O_j = P V: A K-major, B (V) MN-major
      int slot = 0, qb = 0;
      uint32_t ph = 0, qph[2] = {0, 0}, oph = 0;
      uint32_t sph[2] = {0, 0}, pph[2] = {0, 0};
      auto issue_s = [&](int buf) {
        mbar_wait(&kv_full[slot], ph);
        mbar_wait(&s_empty[buf], sph[buf] ^ 1);
        sph[buf] ^= 1;
        tc_fence_after();
        const uint32_t qa = smem_u32(q_s + qb * AT_TILE), ka = smem_u32(kv_s + slot * AT_TILE);
#pragma unroll
        for (int ks = 0; ks < AT_D / 16; ++ks) umma_f16(tm_s[buf], desc_kmajor(qa, ks), desc_kmajor(ka, ks), idesc_s, ks != 0);
        umma_commit(&kv_empty[slot]);
        umma_commit(&s_full[buf]);
        if (++slot == AT_KV_SLOTS) { slot = 0; ph ^= 1; }
      };
      for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
        mbar_wait(&q_full[qb], qph[qb]);
        qph[qb] ^= 1;
        issue_s(0);                                                      // key tile j always uses S / P buffer j & 1
        for (int j = 0; j < nk; ++j) {
          const int pb = j & 1;
          if (j + 1 < nk) issue_s(pb ^ 1);
          mbar_wait(&p_full[pb], pph[pb]);
          pph[pb] ^= 1;
          mbar_wait(&kv_full[slot], ph);
          if (j == 0) { mbar_wait(o_empty, oph ^ 1); oph ^= 1; }
          tc_fence_after();
          const uint32_t pa = smem_u32(p_s + pb * AT_P_BYTES), va = smem_u32(kv_s + slot * AT_TILE);
#pragma unroll
          for (int ks = 0; ks < AT_BK / 16; ++ks)
            umma_f16(tm_o + 64 * j, make_smem_desc(pa + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), desc_mnmajor(va, ks, 8192),
                     idesc_o, ks != 0);
          umma_commit(&kv_empty[slot]);
          umma_commit(&p_empty[pb]);
          if (++slot == AT_KV_SLOTS) { slot = 0; ph ^= 1; }
        }
        umma_commit(o_full);
        umma_commit(&q_empty[qb]);
        qb ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax warps: thread = (query row, 64-key half)
    const int wq = warp & 3;
    const int grp = ((warp - 4) >> 2) & 1;         // S / P buffer = key-tile parity
    const int sub = (warp - 4) >> 3;               // key columns [64 sub, 64 sub + 64) of the tile
    const int gid = sub * 2 + grp;                 // 0..3
    const int r = wq * 32 + lane;
    const int st = threadIdx.x - 128;              // 0..511
    uint32_t sph = 0, pph = 0, oph = 0;
    uint32_t drop_s0 = 0, drop_s1 = 0;
    if (DROP) { drop_s0 = __ldg(p.drop.seed); drop_s1 = __ldg(p.drop.seed + 1); }
    int xset = 0;                                  // exchange-slot set, alternating per tile of this buffer
    int cur_h = -1;
    int ipar = 0;                                  // item parity: the combine arrays are double-buffered, so the only
                                                   // CTA-wide barrier per item is the one in front of the combination
    // the row's gate value of item i + 1 is requested while item i is processed (it sat as a dependent L2 round trip in
    // front of the first score of every item: 11 % of the softmax warps' stall samples)
    auto gate_of = [&](int it) {
      const int qt_ = it % p.nq, b_ = (it / p.nq) % p.B, h_ = it / (p.nq * p.B);
      const int q_ = min(qt_ * AT_BQ + r, p.T - 1);
      return p.gate[(static_cast<long long>(b_) * p.H + h_) * p.T + q_];
    };
    const int it_end = item_end(p.n_items);
    float g_next = item_begin(p.n_items) < it_end ? gate_of(item_begin(p.n_items)) : 0.f;
    for (int item = item_begin(p.n_items); item < it_end; ++item) {
      const int qt = item % p.nq, b = (item / p.nq) % p.B, h = item / (p.nq * p.B);
      if (h != cur_h) {                            // uniform over the CTA: the table changes once per head
        named_bar_sync(1, 512);                    // every softmax thread is done with the previous head's table
        const float* trow = p.table + static_cast<long long>(h) * (2 * p.T - 1);
        for (int i = st; i < 2 * p.T - 1; i += 512) tbl_s[i] = trow[i] * LOG2E;
        cur_h = h;
        named_bar_sync(1, 512);
      }
      float* tmx = tm_sm + ipar * 512;
      float* tlx = tl_sm + ipar * 1024;
      ipar ^= 1;
      const int q = qt * AT_BQ + r;
      const bool q_ok = q < p.T;
      const int qc = q_ok ? q : p.T - 1;
      const int kl = p.klen ? min(p.klen[b], p.T) : p.T;
      const float g = g_next;
      if (item + 1 < it_end) g_next = gate_of(item + 1);
      const float* trel = tbl_s + (p.T - 1 - qc);   // trel[k] = log2e * table[h, k - q + T - 1]
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tm_mine = tm_s[grp] + lane_off + sub * 64;

      for (int j = grp; j < nk; j += 2) {
        mbar_wait(&s_full[grp], sph);
        sph ^= 1;
        tc_fence_after();
        const int k0 = j * AT_BK + sub * 64;
        uint32_t va[32], vb[32];
        tmem_ld32(tm_mine, va);
        tmem_ld32(tm_mine + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[grp]);   // S is in registers: the buffer may take tile j + 2
        // z (log2 units) in place, masked keys -> -inf; tile-local max of this half
        // four independent max chains (a single running max is a 64-deep dependent FMNMX chain per thread and tile: the
        // softmax warps are latency-bound, not issue-bound)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t(&cur)[32] = c ? vb : va;
          const int kb = k0 + c * 32;
          if (kb + 32 <= kl) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float z = fmaf(__uint_as_float(cur[i]), p.scale_log2, g * trel[kb + i]);
              cur[i] = __float_as_uint(z);
              mx4[i & 3] = fmaxf(mx4[i & 3], z);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float z = -INFINITY;
              if (kb + i < kl) z = fmaf(__uint_as_float(cur[i]), p.scale_log2, g * trel[kb + i]);
              cur[i] = __float_as_uint(z);
              mx4[i & 3] = fmaxf(mx4[i & 3], z);
            }
          }
        }
        const float mh = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // the two halves of a tile share one maximum (one scale per row of P V): exchange through shared memory
        float* xs = xm_s + xset * 512 + grp * 256;
        xs[sub * 128 + r] = mh;
        named_bar_sync(2 + grp, 256);
        const float mj = fmaxf(mh, xs[(sub ^ 1) * 128 + r]);
        xset ^= 1;
        const float mm = mj == -INFINITY ? 0.f : mj;
        mbar_wait(&p_empty[grp], pph ^ 1);
        pph ^= 1;
        uint8_t* chunk = p_s + grp * AT_P_BYTES + sub * 16384;   // this group's 64-key swizzle chunk of the P tile
        float lh = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t(&cur)[32] = c ? vb : va;
          float e[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) e[i] = ex2_approx(__uint_as_float(cur[i]) - mm);   // exp2(-inf) = 0 for masked keys
          {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};   // four independent partial sums instead of one 32-deep FADD chain
#pragma unroll
            for (int i = 0; i < 32; ++i) s4[i & 3] += e[i];
            lh += (s4[0] + s4[1]) + (s4[2] + s4[3]);
          }
          if (DROP) {   // denominator over the undropped probabilities, P V over the dropped ones
            float dm[32];
            attn_drop_mults<32>(drop_s0, drop_s1, p.drop, (static_cast<unsigned long long>(b) * p.H + h) * p.T + qc, k0 + c * 32, dm);
#pragma unroll
            for (int i = 0; i < 32; ++i) e[i] *= dm[i];
          }
#pragma unroll
          for (int g16 = 0; g16 < 4; ++g16) {
            uint4 u;
            u.x = pack_bf16x2(e[g16 * 8 + 0], e[g16 * 8 + 1]); u.y = pack_bf16x2(e[g16 * 8 + 2], e[g16 * 8 + 3]);
            u.z = pack_bf16x2(e[g16 * 8 + 4], e[g16 * 8 + 5]); u.w = pack_bf16x2(e[g16 * 8 + 6], e[g16 * 8 + 7]);
            *reinterpret_cast<uint4*>(chunk + swz128(static_cast<uint32_t>(r * 128 + c * 64 + g16 * 16))) = u;
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[grp]);
        if (sub == 0) tmx[j * 128 + r] = mm;
        tlx[(j * 2 + sub) * 128 + r] = lh;
      }
      named_bar_sync(1, 512);
      // exact combination over the key tiles
      float m = -INFINITY;
      for (int j = 0; j < nk; ++j) {
        const float lj = tlx[(j * 2) * 128 + r] + tlx[(j * 2 + 1) * 128 + r];
        if (lj > 0.f) m = fmaxf(m, tmx[j * 128 + r]);
      }
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nk) {
          const float lj = tlx[(j * 2) * 128 + r] + tlx[(j * 2 + 1) * 128 + r];
          f[j] = lj > 0.f ? ex2_approx(tmx[j * 128 + r] - m) : 0.f;
          l = fmaf(f[j], lj, l);
        }
      }
      // epilogue: sum_j f_j O_j / l -> bf16 (group gid stores head-dim columns 16 gid .. 16 gid + 15), LSE
      mbar_wait(o_full, oph);
      oph ^= 1;
      tc_fence_after();
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nk) {
          uint32_t o0[16];
          tmem_ld16(tm_o + 64 * j + lane_off + gid * 16, o0);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[i] = fmaf(f[j], __uint_as_float(o0[i]), acc[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      if (q_ok) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        __nv_bfloat16* orow = p.out + (static_cast<long long>(b) * p.T + q) * (p.H * AT_D) + h * AT_D + gid * 16;
#pragma unroll
        for (int g16 = 0; g16 < 2; ++g16) {
          uint4 u;
          u.x = pack_bf16x2(acc[g16 * 8 + 0] * inv, acc[g16 * 8 + 1] * inv);
          u.y = pack_bf16x2(acc[g16 * 8 + 2] * inv, acc[g16 * 8 + 3] * inv);
          u.z = pack_bf16x2(acc[g16 * 8 + 4] * inv, acc[g16 * 8 + 5] * inv);
          u.w = pack_bf16x2(acc[g16 * 8 + 6] * inv, acc[g16 * 8 + 7] * inv);
          reinterpret_cast<uint4*>(orow)[g16] = u;
        }
        if (gid == 0)
          p.lse[(static_cast<long long>(b) * p.H + h) * p.T + q] = l > 0.f ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA owns a 128-key tile j of one (utterance, head) and walks the query tiles i.  Per (i, j):
//   S  = Q_i K_j^T, dP = dO_i V_j^T                     (TMEM, 2 x 128 columns)
//   P  = exp(z - lse),  dZ = P * (dP - delta),  dS = scale * dZ        (softmax warps, thread = query row)
//   dV_j += P^T dO_i,  dK_j += dS^T Q_i,  dQ_i = dS K_j                (P / dS staged once in swizzled smem and consumed
//                                                                       K-major or MN-major through the descriptors)
// dK_j / dV_j accumulate in TMEM over i and are stored once; dQ_i partials are added to an fp32 scratch with vector
// reductions (red.global.add.v4.f32); dgate[q] = sum_k dZ * table and the Toeplitz table gradient (sums of gate * dZ
// along diagonals, read back from the dS tile by 4 reducer warps and accumulated in shared memory) complete the bias path.
static constexpr int ATB_THREADS = 512;   // 4 control warps + 2 softmax warpgroups (key columns 0-63 / 64-127) + 4 reducer warps

struct AttnBwdP {
  int B, H, T, nq, nk;
  int n_items;
  float scale, scale_log2;
  const float* gate;       // (B,H,T)
  const float* table;      // (H,2T-1)
  const int* klen;         // (B) or null
  const float* lse;        // (B,H,T)
  const float* delta;      // (B,H,T) = sum_d dO * O
  __nv_bfloat16* dqkv;     // (B*T, 3*H*64): dK / dV written here
  float* dq32;             // (B*T, H*64) fp32, zero-initialised, receives dQ partial sums
  float* dgate;            // (B,H,T) zero-initialised (atomics)
  float* dtable;           // (H,2T-1) zero-initialised (atomics)
  DropP drop;              // dropout on the attention probabilities: must be the forward's (seed, site, keep)
  int dbg;                 // timing experiments only (MTASR_ATTN_DBG): 1 skip table-gradient diagonals, 2 skip dQ reductions, 4 skip softmax math
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool DROP>
__global__ void __launch_bounds__(ATB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_do, const AttnBwdP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  pdl_trigger();
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* k_s = smem;                          // 16 KB
  uint8_t* v_s = k_s + AT_TILE;                 // 16 KB
  uint8_t* q_s = v_s + AT_TILE;                 // 2 x 16 KB
  uint8_t* do_s = q_s + 2 * AT_TILE;            // 2 x 16 KB
  uint8_t* p_s = do_s + 2 * AT_TILE;            // 32 KB (single: only MMA 3 reads it, and quickly)
  uint8_t* ds_s = p_s + AT_P_BYTES;             // 2 x 32 KB (double: MMA 4/5 AND the reducer warps read it, slowly)
  float* tbl_s = reinterpret_cast<float*>(ds_s + 2 * AT_P_BYTES);                              // 2T-1 (+slack)
  float* acc_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tbl_s) + attn_table_bytes(p.T));   // 2T-1
  float* g_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(acc_s) + attn_table_bytes(p.T));     // 2 x 128
  uint64_t* bars = reinterpret_cast<uint64_t*>(g_s + 256);
  uint64_t* kv_full = bars, *kv_empty = bars + 1;
  uint64_t* qdo_full = bars + 2, *qdo_empty = bars + 4;     // [2] each
  uint64_t* sdp_full = bars + 6, *sdp_empty = bars + 7;
  uint64_t* pds_full = bars + 8;                            // [2] P + dS[buf] (+ gate[buf]) written by the softmax warps
  uint64_t* ds_empty = bars + 10;                           // [2] dS[buf] consumed by MMA 4/5 and the 4 reducer warps
  uint64_t* p_empty = bars + 12;                            // P consumed by MMA 3
  uint64_t* dq_full = bars + 13, *dq_empty = bars + 14;
  uint64_t* dkv_full = bars + 15, *dkv_empty = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (elect_one()) { prefetch_tmap(&tmap_qkv); prefetch_tmap(&tmap_do); }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
      mbar_init(sdp_full, 1); mbar_init(sdp_empty, 8);
      for (int i = 0; i < 2; ++i) { mbar_init(&pds_full[i], 8); mbar_init(&ds_empty[i], 5); }   // MMA commit + 4 reducer warps
      mbar_init(p_empty, 1);
      mbar_init(dq_full, 1); mbar_init(dq_empty, 4);
      mbar_init(dkv_full, 1); mbar_init(dkv_empty, 8);
      fence_barrier_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // PDL (common.cuh): everything read below (q/k/v, O-gradient, lse, delta, zeroed accumulators) is a predecessor's result
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 128, tm_dv = tmem_base + 256, tm_dk = tmem_base + 320,
                 tm_dq = tmem_base + 384;
  const int nq = p.nq;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      uint32_t kvph = 0, qph[2] = {0, 0};
      int st = 0;
      for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
        const int jt = item % p.nk, b = (item / p.nk) % p.B, h = item / (p.nk * p.B);
        mbar_wait(kv_empty, kvph ^ 1);
        kvph ^= 1;
        mbar_arrive_expect_tx(kv_full, 2 * AT_TILE);
        tma_load_4d(k_s, &tmap_qkv, kv_full, 0, jt * AT_BK, p.H + h, b);
        tma_load_4d(v_s, &tmap_qkv, kv_full, 0, jt * AT_BK, 2 * p.H + h, b);
        for (int i = 0; i < nq; ++i) {
          mbar_wait(&qdo_empty[st], qph[st] ^ 1);
          qph[st] ^= 1;
          mbar_arrive_expect_tx(&qdo_full[st], 2 * AT_TILE);
          tma_load_4d(q_s + st * AT_TILE, &tmap_qkv, &qdo_full[st], 0, i * AT_BQ, h, b);
          tma_load_4d(do_s + st * AT_TILE, &tmap_do, &qdo_full[st], 0, i * AT_BQ, h, b);
          st ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      const uint32_t idesc_kk = make_idesc_bf16(128, 128, 0, 0);   // S, dP: both operands K-major
      const uint32_t idesc_mm = make_idesc_bf16(128, 64, 1, 1);    // dV, dK: both MN-major
      const uint32_t idesc_km = make_idesc_bf16(128, 64, 0, 1);    // dQ: dS K-major, K_j MN-major
      uint32_t kvph = 0, qph[2] = {0, 0}, sdp_ph = 0, pds_ph[2] = {0, 0}, dq_ph = 0, dkv_ph = 0;
      int st = 0;   // stage of the NEXT S/dP issue
      int db = 0;   // dS buffer of the current tile (alternates per tile, continuing across items)
      const uint32_t ka = smem_u32(k_s), va = smem_u32(v_s), pa = smem_u32(p_s);
      auto issue_sdp = [&]() {
        mbar_wait(&qdo_full[st], qph[st]);
        qph[st] ^= 1;
        mbar_wait(sdp_empty, sdp_ph ^ 1);
        sdp_ph ^= 1;
        tc_fence_after();
        const uint32_t qa = smem_u32(q_s + st * AT_TILE), doa = smem_u32(do_s + st * AT_TILE);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16(tm_s, desc_kmajor(qa, ks), desc_kmajor(ka, ks), idesc_kk, ks != 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16(tm_dp, desc_kmajor(doa, ks), desc_kmajor(va, ks), idesc_kk, ks != 0);
        umma_commit(sdp_full);
        st ^= 1;
      };
      for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
        mbar_wait(kv_full, kvph);
        kvph ^= 1;
        issue_sdp();
        for (int i = 0; i < nq; ++i) {
          const int cur = st ^ 1;                     // stage holding Q_i / dO_i
          if (i + 1 < nq) issue_sdp();
          const int stage_i = (i + 1 < nq) ? (st) : cur;   // after the extra issue `st` points back at tile i's stage
          mbar_wait(&pds_full[db], pds_ph[db]);
          pds_ph[db] ^= 1;
          mbar_wait(dq_empty, dq_ph ^ 1);
          dq_ph ^= 1;
          if (i == 0) { mbar_wait(dkv_empty, dkv_ph ^ 1); dkv_ph ^= 1; }
          tc_fence_after();
          const uint32_t qa = smem_u32(q_s + stage_i * AT_TILE), doa = smem_u32(do_s + stage_i * AT_TILE);
          const uint32_t dsa = smem_u32(ds_s + db * AT_P_BYTES);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // dV += P^T dO_i
            umma_f16(tm_dv, desc_mnmajor(pa, ks, 16384), desc_mnmajor(doa, ks, 8192), idesc_mm, (i | ks) != 0);
          umma_commit(p_empty);            // P may be overwritten as soon as MMA 3 has read it
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // dK += dS^T Q_i
            umma_f16(tm_dk, desc_mnmajor(dsa, ks, 16384), desc_mnmajor(qa, ks, 8192), idesc_mm, (i | ks) != 0);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // dQ_i = dS K_j
            umma_f16(tm_dq, make_smem_desc(dsa + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), desc_mnmajor(ka, ks, 8192),
                     idesc_km, ks != 0);
          umma_commit(dq_full);
          umma_commit(&ds_empty[db]);
          umma_commit(&qdo_empty[stage_i]);
          db ^= 1;
        }
        umma_commit(dkv_full);
        umma_commit(kv_empty);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ---------------------------------------------------------------- softmax / dZ warps: thread = query row; warpgroup
    // `grp` owns key columns [64 grp, 64 grp + 64) of every tile (= one 64-key swizzle chunk of the P / dS staging)
    const int wq = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int r = wq * 32 + lane;
    const int tid = threadIdx.x - 128;             // 0..255 over both groups
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    uint32_t sdp_ph = 0, pe_ph = 0, dse_ph[2] = {0, 0}, dkv_ph = 0;
    int db = 0;
    int cur_h = -1;
    uint32_t drop_s0 = 0, drop_s1 = 0;
    if (DROP) { drop_s0 = __ldg(p.drop.seed); drop_s1 = __ldg(p.drop.seed + 1); }
    for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
      const int jt = item % p.nk, b = (item / p.nk) % p.B, h = item / (p.nk * p.B);
      named_bar_sync(1, 256);
      if (h != cur_h) {
        const float* trow = p.table + static_cast<long long>(h) * (2 * p.T - 1);
        for (int i = tid; i < 2 * p.T - 1; i += 256) tbl_s[i] = trow[i] * LOG2E;
        cur_h = h;
      }
      named_bar_sync(1, 256);
      const int kl = p.klen ? min(p.klen[b], p.T) : p.T;
      const int k0 = jt * AT_BK + grp * 64;
      const long long bh = static_cast<long long>(b) * p.H + h;
      // per-row scalars of query tile i + 1 are requested while tile i is processed (three dependent L2 round trips
      // per tile otherwise sit between the S / dP MMAs and the first exponential)
      float g_n, lse_n, delta_n;
      {
        const int qc0 = min(r, p.T - 1);
        g_n = p.gate[bh * p.T + qc0]; lse_n = p.lse[bh * p.T + qc0]; delta_n = p.delta[bh * p.T + qc0];
      }
      for (int i = 0; i < nq; ++i) {
        const int q = i * AT_BQ + r;
        const bool q_ok = q < p.T;
        const int qc = q_ok ? q : p.T - 1;
        const float g = g_n;
        const float nlse2 = -lse_n * LOG2E;
        const float delta = delta_n;
        if (i + 1 < nq) {
          const int qn = min((i + 1) * AT_BQ + r, p.T - 1);
          g_n = p.gate[bh * p.T + qn]; lse_n = p.lse[bh * p.T + qn]; delta_n = p.delta[bh * p.T + qn];
        }
        const float* trel = tbl_s + (p.T - 1 - qc);
        float dg = 0.f;
        mbar_wait(sdp_full, sdp_ph);
        sdp_ph ^= 1;
        tc_fence_after();
        uint8_t* pc = p_s + grp * 16384;
        uint8_t* dc = ds_s + db * AT_P_BYTES + grp * 16384;
        uint32_t svb[2][16], dvb[2][16];
        tmem_ld16(tm_s + lane_off + grp * 64, svb[0]);
        tmem_ld16(tm_dp + lane_off + grp * 64, dvb[0]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t(&sv)[16] = svb[c & 1];
          uint32_t(&dv)[16] = dvb[c & 1];
          if (c + 1 < 4) {   // next chunk's TMEM loads fly while this chunk is processed
            tmem_ld16(tm_s + lane_off + grp * 64 + (c + 1) * 16, svb[(c + 1) & 1]);
            tmem_ld16(tm_dp + lane_off + grp * 64 + (c + 1) * 16, dvb[(c + 1) & 1]);
          }
          if (c == 0) {   // P of the previous tile (MMA 3) and dS / gate of two tiles ago (MMA 4/5 + reducers) consumed
            mbar_wait(p_empty, pe_ph ^ 1);
            pe_ph ^= 1;
            mbar_wait(&ds_empty[db], dse_ph[db] ^ 1);
            dse_ph[db] ^= 1;
            if (grp == 0) g_s[db * 128 + r] = q_ok ? g : 0.f;
          }
          const int kb = k0 + c * 16;
          float pe[16], de[16];
          // dropout on the probabilities: O = (P o M) V with M = mask / keep, so dP = M o (dO V^T), the P^T dO contraction of dV
          // takes P o M, and delta = rowsum(dO o O) is unchanged
          float dm[16];
          if (DROP) attn_drop_mults<16>(drop_s0, drop_s1, p.drop, static_cast<unsigned long long>(bh) * p.T + qc, kb, dm);
          float dg4[4] = {0.f, 0.f, 0.f, 0.f};   // independent partial sums of the gate gradient (no 64-deep FFMA chain)
          if (p.dbg & 4) {   // timing experiment: no per-element math
#pragma unroll
            for (int e = 0; e < 16; ++e) { pe[e] = __uint_as_float(sv[e]); de[e] = __uint_as_float(dv[e]); }
          } else if (q_ok && kb + 16 <= kl) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float t = trel[kb + e];
              const float pv = ex2_approx(fmaf(__uint_as_float(sv[e]), p.scale_log2, fmaf(g, t, nlse2)));
              const float dpv = DROP ? __uint_as_float(dv[e]) * dm[e] : __uint_as_float(dv[e]);
              const float dz = pv * (dpv - delta);
              dg4[e & 3] = fmaf(dz, t, dg4[e & 3]);
              pe[e] = DROP ? pv * dm[e] : pv;
              de[e] = dz * p.scale;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const bool ok = q_ok && (kb + e < kl);
              // masked keys may index past the 2T-1 table entries (uninitialised shared memory): 0 * NaN would poison dg
              const float t = ok ? trel[kb + e] : 0.f;
              const float pv = ok ? ex2_approx(fmaf(__uint_as_float(sv[e]), p.scale_log2, fmaf(g, t, nlse2))) : 0.f;
              const float dpv = DROP ? __uint_as_float(dv[e]) * dm[e] : __uint_as_float(dv[e]);
              const float dz = pv * (dpv - delta);
              dg4[e & 3] = fmaf(dz, t, dg4[e & 3]);
              pe[e] = DROP ? pv * dm[e] : pv;
              de[e] = dz * p.scale;
            }
          }
          dg += (dg4[0] + dg4[1]) + (dg4[2] + dg4[3]);
#pragma unroll
          for (int g16 = 0; g16 < 2; ++g16) {
            const uint32_t off = swz128(static_cast<uint32_t>(r * 128 + c * 32 + g16 * 16));
            uint4 u;
            u.x = pack_bf16x2(pe[g16 * 8 + 0], pe[g16 * 8 + 1]); u.y = pack_bf16x2(pe[g16 * 8 + 2], pe[g16 * 8 + 3]);
            u.z = pack_bf16x2(pe[g16 * 8 + 4], pe[g16 * 8 + 5]); u.w = pack_bf16x2(pe[g16 * 8 + 6], pe[g16 * 8 + 7]);
            *reinterpret_cast<uint4*>(pc + off) = u;
            u.x = pack_bf16x2(de[g16 * 8 + 0], de[g16 * 8 + 1]); u.y = pack_bf16x2(de[g16 * 8 + 2], de[g16 * 8 + 3]);
            u.z = pack_bf16x2(de[g16 * 8 + 4], de[g16 * 8 + 5]); u.w = pack_bf16x2(de[g16 * 8 + 6], de[g16 * 8 + 7]);
            *reinterpret_cast<uint4*>(dc + off) = u;
          }
          if (c + 1 < 4) tmem_ld_wait();
          if (c == 2) {   // all TMEM reads of this tile are in registers: the next S / dP may be issued
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sdp_empty);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_full[db]);
        db ^= 1;
        if (q_ok && dg != 0.f) atomicAdd(p.dgate + bh * p.T + q, dg * 0.6931471805599453f);   // table was scaled by log2e
      }
      // item epilogue: group 0 stores dK_j, group 1 stores dV_j (rows = keys of this tile) as bf16 into dqkv
      mbar_wait(dkv_full, dkv_ph);
      dkv_ph ^= 1;
      tc_fence_after();
      const int key = jt * AT_BK + r;
      {
        const uint32_t src = grp == 0 ? tm_dk : tm_dv;
        __nv_bfloat16* dst = p.dqkv + (static_cast<long long>(b) * p.T + (key < p.T ? key : 0)) * (3 * p.H * AT_D) +
                             (grp == 0 ? 1 : 2) * p.H * AT_D + h * AT_D;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t a0[16];
          tmem_ld16(src + lane_off + c * 16, a0);
          tmem_ld_wait();
          if (key < p.T) {
#pragma unroll
            for (int g16 = 0; g16 < 2; ++g16) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(a0[g16 * 8 + 0]), __uint_as_float(a0[g16 * 8 + 1]));
              u.y = pack_bf16x2(__uint_as_float(a0[g16 * 8 + 2]), __uint_as_float(a0[g16 * 8 + 3]));
              u.z = pack_bf16x2(__uint_as_float(a0[g16 * 8 + 4]), __uint_as_float(a0[g16 * 8 + 5]));
              u.w = pack_bf16x2(__uint_as_float(a0[g16 * 8 + 6]), __uint_as_float(a0[g16 * 8 + 7]));
              reinterpret_cast<uint4*>(dst)[c * 2 + g16] = u;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkv_empty);
    }
  } else if (warp >= 12) {
    // ---------------------------------------------------------------- reducer warps: table-gradient diagonals + dQ drain
    const int wq = warp & 3;
    const int t = wq * 32 + lane;                 // 0..127
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    uint32_t pds_ph[2] = {0, 0}, dq_ph = 0;
    int db = 0;
    const float inv_scale = 1.f / p.scale;
    int red_h = -1;
    for (int item = item_begin(p.n_items); item < item_end(p.n_items); ++item) {
      const int jt = item % p.nk, b = (item / p.nk) % p.B, h = item / (p.nk * p.B);
      if (h != red_h) {                            // uniform: the accumulator follows the head (once or twice per CTA)
        named_bar_sync(2, 128);                    // every diagonal sum of the previous head has landed
        float* dt = p.dtable + static_cast<long long>(red_h < 0 ? 0 : red_h) * (2 * p.T - 1);
        for (int i = t; i < 2 * p.T - 1; i += 128) {
          const float v = acc_s[i];
          if (red_h >= 0 && v != 0.f) atomicAdd(dt + i, v * inv_scale);
          acc_s[i] = 0.f;
        }
        red_h = h;
        named_bar_sync(2, 128);
      }
      for (int i = 0; i < nq; ++i) {
        mbar_wait(&pds_full[db], pds_ph[db]);
        pds_ph[db] ^= 1;
        const uint8_t* dsb = ds_s + db * AT_P_BYTES;
        const float* gsb = g_s + db * 128;
        // Diagonal sums of gate * dS over the 128 x 128 tile, two bf16 per 32-bit shared-memory load: thread beta = t - 64
        // walks the aligned column pairs (2w, 2w+1), w = r + beta, of the row pair (2r, 2r+1).  On the even row they lie on
        // the local diagonals (2 beta, 2 beta + 1), on the odd row on (2 beta - 1, 2 beta): three running sums per thread,
        // 3 instructions per element instead of 11 with one 16-bit load per element (the reducer warps were the largest
        // consumer of issue slots of the kernel).  Odd diagonals have two owners -> shared-memory atomics at the end.
        const int base = (jt - i) * 128 + p.T - 1;
        if (!(p.dbg & 1)) {
          const int beta = t - 64;
          const int r_lo = beta < 0 ? -beta : 0, r_hi = beta > 0 ? 64 - beta : 64;
          float aM = 0.f, a0 = 0.f, aP = 0.f;
#pragma unroll 4
          for (int r = r_lo; r < r_hi; ++r) {
            const int w = r + beta;
            // element (row, col) of a 64-column chunk lives at row * 128 + ((col * 2) ^ ((row & 7) << 4)) (128-byte swizzle)
            const uint32_t chunk = static_cast<uint32_t>(w >> 5) * 16384u + static_cast<uint32_t>(2 * r) * 128u;
            const uint32_t cb = static_cast<uint32_t>(w & 31) * 4u;
            const uint32_t we = *reinterpret_cast<const uint32_t*>(dsb + chunk + (cb ^ ((static_cast<uint32_t>(2 * r) & 7u) << 4)));
            const uint32_t wo = *reinterpret_cast<const uint32_t*>(dsb + chunk + 128u + (cb ^ ((static_cast<uint32_t>(2 * r + 1) & 7u) << 4)));
            const float2 g2 = *reinterpret_cast<const float2*>(gsb + 2 * r);
            a0 = fmaf(__uint_as_float(we << 16), g2.x, a0);
            aP = fmaf(__uint_as_float(we & 0xffff0000u), g2.x, aP);
            aM = fmaf(__uint_as_float(wo << 16), g2.y, aM);
            a0 = fmaf(__uint_as_float(wo & 0xffff0000u), g2.y, a0);
          }
          const int g0 = base + 2 * beta;
          if (r_lo < r_hi) {
            if (g0 - 1 >= 0 && g0 - 1 <= 2 * p.T - 2) atomicAdd(acc_s + g0 - 1, aM);
            if (g0 >= 0 && g0 <= 2 * p.T - 2) atomicAdd(acc_s + g0, a0);
            if (g0 + 1 >= 0 && g0 + 1 <= 2 * p.T - 2) atomicAdd(acc_s + g0 + 1, aP);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_empty[db]);
        db ^= 1;
        // dQ_i partial: TMEM -> fp32 vector reductions into the scratch
        mbar_wait(dq_full, dq_ph);
        dq_ph ^= 1;
        tc_fence_after();
        uint32_t a0[32], a1[32];
        tmem_ld32(tm_dq + lane_off, a0);
        tmem_ld32(tm_dq + lane_off + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_empty);
        const int q = i * AT_BQ + t;
        if (q < p.T && !(p.dbg & 2)) {
          float* dst = p.dq32 + (static_cast<long long>(b) * p.T + q) * (p.H * AT_D) + h * AT_D;
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4)
            red_add_v4(dst + v4 * 4, __uint_as_float(a0[v4 * 4]), __uint_as_float(a0[v4 * 4 + 1]), __uint_as_float(a0[v4 * 4 + 2]),
                       __uint_as_float(a0[v4 * 4 + 3]));
#pragma unroll
          for (int v4 = 0; v4 < 8; ++v4)
            red_add_v4(dst + 32 + v4 * 4, __uint_as_float(a1[v4 * 4]), __uint_as_float(a1[v4 * 4 + 1]),
                       __uint_as_float(a1[v4 * 4 + 2]), __uint_as_float(a1[v4 * 4 + 3]));
        }
      }
    }
    named_bar_sync(2, 128);
    if (red_h >= 0) {
      float* dt = p.dtable + static_cast<long long>(red_h) * (2 * p.T - 1);
      for (int i = t; i < 2 * p.T - 1; i += 128) {
        const float v = acc_s[i];
        if (v != 0.f) atomicAdd(dt + i, v * inv_scale);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]  (the softmax-backward row term), one thread per (b,q,h)
// The same launch zeroes the three accumulation targets of the backward kernel that follows it (dQ scratch, gate and table
// gradients): they were two fill launches per layer in front of this kernel.
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, int B, int T, int H,
                                  float* __restrict__ delta, float4* __restrict__ zero4, long long n_zero4,
                                  float* __restrict__ zero_a, long long n_zero_a, float* __restrict__ zero_b, long long n_zero_b) {
  pdl_trigger();
  const long long n = static_cast<long long>(B) * T * H;
  const long long tid0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = tid0; i < n_zero4; i += nthr) zero4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = tid0; i < n_zero_a; i += nthr) zero_a[i] = 0.f;
  for (long long i = tid0; i < n_zero_b; i += nthr) zero_b[i] = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int h = static_cast<int>(i % H);
    const long long bq = i / H;
    const uint4* a = reinterpret_cast<const uint4*>(o + bq * H * AT_D + h * AT_D);
    const uint4* d = reinterpret_cast<const uint4*>(dout + bq * H * AT_D + h * AT_D);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 x = a[c], y = d[c];
      float2 u, v;
      u = unpack_bf16x2(x.x); v = unpack_bf16x2(y.x); s += u.x * v.x + u.y * v.y;
      u = unpack_bf16x2(x.y); v = unpack_bf16x2(y.y); s += u.x * v.x + u.y * v.y;
      u = unpack_bf16x2(x.z); v = unpack_bf16x2(y.z); s += u.x * v.x + u.y * v.y;
      u = unpack_bf16x2(x.w); v = unpack_bf16x2(y.w); s += u.x * v.x + u.y * v.y;
    }
    const int b = static_cast<int>(bq / T), q = static_cast<int>(bq % T);
    delta[(static_cast<long long>(b) * H + h) * T + q] = s;
  }
}

// y[r][0..cols) (bf16, row stride ldy) = x[r][0..cols) (fp32, row stride ldx): dQ scratch -> the q block of dqkv
__global__ void cast2d_kernel(const float* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy, long long rows,
                              int cols8) {
  pdl_trigger();
  const long long n = rows * cols8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols8;
    const int c = static_cast<int>(i % cols8) * 8;
    const float4 a = *reinterpret_cast<const float4*>(x + r * ldx + c);
    const float4 b = *reinterpret_cast<const float4*>(x + r * ldx + c + 4);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w); u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(y + r * ldy + c) = u;
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 attn_encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(ptr);
  }
  return fn;
}

// (B*T, n_slices*64) bf16 matrix viewed as [b][slice][t][64]: box = 128 rows x 64 columns of one (b, slice);
// rows >= T are out of bounds of dimension 1 and therefore zero-filled (never the next utterance's rows).
static int encode_heads_map(CUtensorMap* map, const void* base, int B, int T, int n_slices, long long ld, const char* what) {
  EncodeTiledFn2 fn = attn_encode_fn();
  if (!fn) return set_error(MTASR_ERR_DRIVER, "attention: cuTensorMapEncodeTiled unavailable");
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0 || (ld * 2) % 16 != 0)
    return set_error(MTASR_ERR_INVALID_ARG, "attention: %s must be 16-byte aligned with a 16-byte multiple row stride", what);
  cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(n_slices), static_cast<cuuint64_t>(B)};
  cuuint64_t str[3] = {static_cast<cuuint64_t>(ld) * 2, 128, static_cast<cuuint64_t>(ld) * 2 * T};
  cuuint32_t box[4] = {64, 128, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, str, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(MTASR_ERR_DRIVER, "attention: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return 0;
}

}  // namespace mtasr

using namespace mtasr;

static int attn_drop_params(DropP* d, const void* seed, uint32_t site, uint32_t keep16, int T) {
  d->seed = nullptr;
  d->site = site;
  d->thresh = 65536;
  d->scale = 1.f;
  d->ld = (static_cast<long long>(T) + 1) & ~1LL;
  if (seed == nullptr || keep16 >= 65536) return 0;
  if (keep16 == 0) return set_error(MTASR_ERR_INVALID_ARG, "attention: dropout keep16 must be in (0, 65536]");
  d->seed = reinterpret_cast<const uint32_t*>(seed);
  d->thresh = keep16;
  d->scale = 65536.0f / static_cast<float>(keep16);
  return 0;
}

extern "C" int mtasr_attn_fwd(const void* qkv, const float* gate, const float* table, const int32_t* klen, int32_t B, int32_t H,
                              int32_t T, float scale, void* out, float* lse, const void* drop_seed, uint32_t drop_site,
                              uint32_t drop_keep16, void* stream) {
  MTASR_CHECK_ARG(qkv && gate && table && out && lse && B > 0 && H > 0 && T > 0, "attn_fwd: bad arguments");
  MTASR_CHECK_ARG(T <= 8192, "attn_fwd: T=%d too long for the shared-memory bias table", T);
  CUtensorMap map;
  if (int rc = encode_heads_map(&map, qkv, B, T, 3 * H, 3LL * H * AT_D, "qkv")) return rc;
  AttnFwdP p;
  p.B = B; p.H = H; p.T = T;
  p.nq = (T + AT_BQ - 1) / AT_BQ;
  p.nk = (T + AT_BK - 1) / AT_BK;
  p.n_items = p.nq * H * B;
  p.scale_log2 = scale * LOG2E;
  p.gate = gate; p.table = table; p.klen = klen;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  if (int rc = attn_drop_params(&p.drop, drop_seed, drop_site, drop_keep16, T)) return rc;
  const bool drop = p.drop.seed != nullptr;
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  if (p.nk <= 4 && getenv("MTASR_ATTN_TWO_PASS") == nullptr) {
    // T <= 512: one O accumulator per key tile, no max-only pass (attn_fwd_sp_kernel)
    const int smem = 2 * AT_TILE + AT_KV_SLOTS * AT_TILE + 2 * AT_P_BYTES + attn_table_bytes(T) + (1024 + 1024 + 2048) * 4 + 256;
    auto kern = drop ? attn_fwd_sp_kernel<true> : attn_fwd_sp_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return set_error(MTASR_ERR_LAUNCH, "attn_fwd: cannot set the shared-memory attribute");
    if (launch_pdl(kern, dim3(grid), dim3(AT_THREADS), smem, static_cast<cudaStream_t>(stream), map, p) != cudaSuccess)
      return set_error(MTASR_ERR_LAUNCH, "attn_fwd: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    MTASR_COUNT_LAUNCH();
    MTASR_CHECK_LAUNCH("attn_fwd");
    return MTASR_OK;
  }
  const int smem = AT_TILE + AT_KV_SLOTS * AT_TILE + 2 * AT_P_BYTES + attn_table_bytes(T) + 2048 + 256;
  MTASR_CHECK_ARG(smem <= 232448, "attn_fwd: T=%d needs %d bytes of shared memory", T, smem);
  auto kern2 = drop ? attn_fwd_kernel<true> : attn_fwd_kernel<false>;
  if (cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return set_error(MTASR_ERR_LAUNCH, "attn_fwd: cannot set the shared-memory attribute");
  if (launch_pdl(kern2, dim3(grid), dim3(AT_THREADS), smem, static_cast<cudaStream_t>(stream), map, p) != cudaSuccess)
    return set_error(MTASR_ERR_LAUNCH, "attn_fwd: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("attn_fwd");
  return MTASR_OK;
}

extern "C" int mtasr_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const float* gate,
                              const float* table, const int32_t* klen, int32_t B, int32_t H, int32_t T, float scale, void* dqkv,
                              float* dq32, float* delta, float* dgate, float* dtable, const void* drop_seed, uint32_t drop_site,
                              uint32_t drop_keep16, void* stream) {
  MTASR_CHECK_ARG(qkv && out && dout && lse && gate && table && dqkv && dq32 && delta && dgate && dtable && B > 0 && H > 0 && T > 0,
                  "attn_bwd: bad arguments");
  MTASR_CHECK_ARG(T <= 4096, "attn_bwd: T=%d too long for the shared-memory bias tables", T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = static_cast<long long>(B) * T * H;
  long long g = (n + 255) / 256;
  if (g > num_sms() * 8) g = num_sms() * 8;
  MTASR_CHECK_ARG((reinterpret_cast<uintptr_t>(dq32) & 15) == 0, "attn_bwd: dq32 must be 16-byte aligned");
  attn_delta_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), B, T, H, delta,
      reinterpret_cast<float4*>(dq32), n * (AT_D / 4), dgate, n, dtable, static_cast<long long>(H) * (2 * T - 1));
  MTASR_COUNT_LAUNCH();
  CUtensorMap mq, mdo;
  if (int rc = encode_heads_map(&mq, qkv, B, T, 3 * H, 3LL * H * AT_D, "qkv")) return rc;
  if (int rc = encode_heads_map(&mdo, dout, B, T, H, 1LL * H * AT_D, "dout")) return rc;
  AttnBwdP p;
  p.B = B; p.H = H; p.T = T;
  p.nq = (T + AT_BQ - 1) / AT_BQ;
  p.nk = (T + AT_BK - 1) / AT_BK;
  p.n_items = p.nk * H * B;
  p.scale = scale;
  p.scale_log2 = scale * LOG2E;
  p.gate = gate; p.table = table; p.klen = klen; p.lse = lse; p.delta = delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.dq32 = dq32; p.dgate = dgate; p.dtable = dtable;
  p.dbg = getenv("MTASR_ATTN_DBG") ? atoi(getenv("MTASR_ATTN_DBG")) : 0;
  if (int rc = attn_drop_params(&p.drop, drop_seed, drop_site, drop_keep16, T)) return rc;
  const int smem = 6 * AT_TILE + 3 * AT_P_BYTES + 2 * attn_table_bytes(T) + 256 * 4 + 256;
  MTASR_CHECK_ARG(smem <= 232448, "attn_bwd: T=%d needs %d bytes of shared memory", T, smem);
  auto kern = p.drop.seed != nullptr ? attn_bwd_kernel<true> : attn_bwd_kernel<false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return set_error(MTASR_ERR_LAUNCH, "attn_bwd: cannot set the shared-memory attribute");
  const int grid = p.n_items < num_sms() ? p.n_items : num_sms();
  if (launch_pdl(kern, dim3(grid), dim3(ATB_THREADS), smem, st, mq, mdo, p) != cudaSuccess)
    return set_error(MTASR_ERR_LAUNCH, "attn_bwd: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  MTASR_COUNT_LAUNCH();
  // dQ scratch (fp32) -> q block of dqkv (bf16)
  const long long rows = static_cast<long long>(B) * T;
  const int cols8 = H * AT_D / 8;
  long long g2 = (rows * cols8 + 255) / 256;
  if (g2 > num_sms() * 8) g2 = num_sms() * 8;
  cast2d_kernel<<<static_cast<unsigned>(g2), 256, 0, st>>>(dq32, 1LL * H * AT_D, reinterpret_cast<__nv_bfloat16*>(dqkv),
                                                           3LL * H * AT_D, rows, cols8);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("attn_bwd");
  return MTASR_OK;
}
