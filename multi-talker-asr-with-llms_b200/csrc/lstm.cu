// Persistent LSTM recurrence for the separator (ref:models/separator.py:6-59: CustomLSTMCell / StackedCustomLSTM).
//
// The reference runs a Python `for t in range(T)` over ~10 tiny kernels per layer.  Here the
// input half x_t W_ih^T + b of every step is ONE batched tcgen05 GEMM (csrc/gemm.cu) done beforehand, and the
// sequential half runs in a single cooperative kernel per layer:
//   * grid = Hs/8 CTAs; CTA j owns hidden units [8j, 8j+8) = 32 gate columns (i,f,g,o) and keeps that slice of
//     W_hh (32 x Hs bf16) resident in shared memory for all T steps;
//   * per step: h_{t-1} (B x Hs bf16) is pulled from L2 into smem, the 8 warps do the (B x Hs)(Hs x 32) product
//     with mma.sync m16n8k16 (latency-bound GEMV-like step: not a tcgen05 shape), gates/cell update in fp32
//     registers, h_t is published and a grid-wide arrive/spin barrier (one atomic per CTA) orders the steps.
// Backward (BPTT) mirrors it: CTA j owns dh_{t-1}[:, 8j:8j+8], streams dgates_t (B x 4Hs bf16) through smem in
// four Hs-wide chunks against its resident W_hh^T slice, split-K over the 8 warps with an smem reduction.
// dW / dx / db are batched GEMMs / column sums over the saved dgates afterwards.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace mtasr {

static constexpr int LU = 8;           // hidden units per CTA
static constexpr int LTHREADS = 256;   // 8 warps
static constexpr int LPAD = 8;         // bf16 row padding (16 B) -> conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Grid barrier: every CTA adds 1 after publishing step data; waiters spin until the count reaches `target`.
__device__ __forceinline__ void grid_arrive(unsigned int* bar) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(bar, 1u);
}
// A waiter that spins longer than `limit` SM clocks traps (a missing co-resident CTA would otherwise hang the GPU until the
// watchdog).  MTASR_LSTM_SPIN_TIMEOUT_MS sets the limit (default 2000 ms at ~2 GHz); 0 disables it -- needed under ncu
// kernel replay, a debugger or GPU time-slicing, where a CTA can legitimately be descheduled for seconds.  The limit
// travels as a kernel parameter (constant bank): read from a __device__ variable it added a dependent L2 round trip in
// front of every barrier poll (lstm_bwd 7.4 -> 11.1 ms per cfg2 step, profiles/launch_summary_r2_v4.txt).
__device__ __forceinline__ void grid_wait(unsigned int* bar, unsigned int target, long long limit) {
  if (threadIdx.x == 0) {
    unsigned int v;
    long long t0 = clock64();
    unsigned int spins = 0;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if ((++spins & 0xfff) == 0 && limit > 0 && clock64() - t0 > limit) __trap();
    }
  }
  __syncthreads();
}

__device__ __forceinline__ uint4 ldcg16(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

struct LstmFwdP {
  const float* xg;            // (B,T,4Hs)
  const __nv_bfloat16* whh;   // (4Hs, ldw): whh[col*ldw + k]
  __nv_bfloat16* h_bf16;      // (B,T,Hs)
  float* h_f32;               // (B,T,Hs) optional
  float* c_all;               // (B,T,Hs)
  float* gates;               // (B,T,4Hs) activations i,f,g,o
  unsigned int* bar;
  int B, T, Hs, ldw, Bp;
  long long spin_limit;       // grid_wait trap threshold in SM clocks (0 = never)
};

__global__ void __launch_bounds__(LTHREADS, 1) lstm_fwd_kernel(const LstmFwdP p) {
  extern __shared__ __align__(16) uint8_t lsm[];
  const int Hs = p.Hs, B = p.B, T = p.T, Bp = p.Bp;
  const int rs = Hs + LPAD;  // smem row stride (elements)
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(lsm);   // [32][rs]
  __nv_bfloat16* Hsm = Ws + 32 * rs;                           // [Bp][rs]
  float* G = reinterpret_cast<float*>(Hsm + Bp * rs);          // [Bp][33]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = blockIdx.x * LU;

  // resident W_hh slice: local column lc = gate*8 + u  <-  global column gate*Hs + u0 + u
  for (int i = tid; i < 32 * (Hs / 8); i += LTHREADS) {
    const int lc = i / (Hs / 8), kc = i % (Hs / 8);
    const int gcol = (lc >> 3) * Hs + u0 + (lc & 7);
    *reinterpret_cast<uint4*>(Ws + lc * rs + kc * 8) =
        *reinterpret_cast<const uint4*>(p.whh + static_cast<long long>(gcol) * p.ldw + kc * 8);
  }
  for (int i = tid; i < Bp * rs; i += LTHREADS) Hsm[i] = f2bf(0.f);
  for (int i = tid; i < Bp * 33; i += LTHREADS) G[i] = 0.f;
  __syncthreads();

  const int npair = B * LU;
  float creg[2] = {0.f, 0.f};  // cell state of the (b, unit) pairs this thread owns (B <= 64)
  const int m_tiles = Bp / 16;

  for (int t = 0; t < T; ++t) {
    float xv[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pr = tid + i * LTHREADS;
      if (pr < npair) {
        const int b = pr >> 3, u = pr & 7;
        const float* xr = p.xg + (static_cast<long long>(b) * T + t) * 4 * Hs + u0 + u;
#pragma unroll
        for (int g = 0; g < 4; ++g) xv[i][g] = xr[g * Hs];
      }
    }
    if (t > 0) {
      grid_wait(p.bar, gridDim.x * static_cast<unsigned>(t), p.spin_limit);
      for (int i = tid; i < B * (Hs / 8); i += LTHREADS) {
        const int b = i / (Hs / 8), kc = i % (Hs / 8);
        *reinterpret_cast<uint4*>(Hsm + b * rs + kc * 8) =
            ldcg16(p.h_bf16 + (static_cast<long long>(b) * T + (t - 1)) * Hs + kc * 8);
      }
      __syncthreads();
      for (int tile = warp; tile < m_tiles * 4; tile += 8) {
        const int mt = tile >> 2, nt = tile & 3;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t a_base = smem_addr(Hsm + (mt * 16 + (lane & 15)) * rs + (lane >> 4) * 8);
        const uint32_t b_base = smem_addr(Ws + (nt * 8 + (lane & 7)) * rs + ((lane >> 3) & 1) * 8);
        for (int k0 = 0; k0 < Hs; k0 += 16) {
          uint32_t a0, a1, a2, a3, b0, b1;
          ldsm_x4(a_base + k0 * 2, a0, a1, a2, a3);
          ldsm_x2(b_base + k0 * 2, b0, b1);
          mma16816(acc, a0, a1, a2, a3, b0, b1);
        }
        const int r = mt * 16 + (lane >> 2), c = nt * 8 + (lane & 3) * 2;
        G[r * 33 + c] = acc[0];
        G[r * 33 + c + 1] = acc[1];
        G[(r + 8) * 33 + c] = acc[2];
        G[(r + 8) * 33 + c + 1] = acc[3];
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pr = tid + i * LTHREADS;
      if (pr < npair) {
        const int b = pr >> 3, u = pr & 7;
        const float ai = xv[i][0] + G[b * 33 + u];
        const float af = xv[i][1] + G[b * 33 + 8 + u];
        const float ag = xv[i][2] + G[b * 33 + 16 + u];
        const float ao = xv[i][3] + G[b * 33 + 24 + u];
        const float ig = 1.f / (1.f + expf(-ai)), fg = 1.f / (1.f + expf(-af));
        const float gg = tanhf(ag), og = 1.f / (1.f + expf(-ao));
        const float c = fg * creg[i] + ig * gg;
        creg[i] = c;
        const float h = og * tanhf(c);
        const long long o = (static_cast<long long>(b) * T + t) * Hs + u0 + u;
        p.h_bf16[o] = f2bf(h);
        if (p.h_f32) p.h_f32[o] = h;
        p.c_all[o] = c;
        float* gr = p.gates + (static_cast<long long>(b) * T + t) * 4 * Hs + u0 + u;
        gr[0] = ig; gr[Hs] = fg; gr[2 * Hs] = gg; gr[3 * Hs] = og;
      }
    }
    if (t + 1 < T) grid_arrive(p.bar);
  }
}

struct LstmBwdP {
  const float* dh_out;        // (B,T,Hs)
  const float* gates;         // (B,T,4Hs)
  const float* c_all;         // (B,T,Hs)
  const __nv_bfloat16* whh;   // (4Hs, ldw)
  __nv_bfloat16* dgates;      // (B,T,4Hs) out, pre-activation gradients
  unsigned int* bar;
  int B, T, Hs, ldw, Bp;
  long long spin_limit;       // grid_wait trap threshold in SM clocks (0 = never)
};

__global__ void __launch_bounds__(LTHREADS, 1) lstm_bwd_kernel(const LstmBwdP p) {
  extern __shared__ __align__(16) uint8_t lsm[];
  const int Hs = p.Hs, B = p.B, T = p.T, Bp = p.Bp;
  const int K4 = 4 * Hs;
  const int wrs = K4 + LPAD;  // W^T slice row stride
  const int ars = Hs + LPAD;  // A chunk row stride
  __nv_bfloat16* Wt = reinterpret_cast<__nv_bfloat16*>(lsm);  // [8][wrs]: Wt[u][col] = whh[col][u0+u]
  __nv_bfloat16* Asm = Wt + LU * wrs;                         // [Bp][ars]
  float* R = reinterpret_cast<float*>(Asm + Bp * ars);        // [8 warps][Bp][8] partials
  float* Dh = R + 8 * Bp * 8;                                 // [Bp][8] reduced dh_rec
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = blockIdx.x * LU;

  for (int i = tid; i < LU * K4; i += LTHREADS) {
    const int u = i % LU, col = i / LU;
    Wt[u * wrs + col] = p.whh[static_cast<long long>(col) * p.ldw + u0 + u];
  }
  for (int i = tid; i < Bp * ars; i += LTHREADS) Asm[i] = f2bf(0.f);
  for (int i = tid; i < Bp * 8; i += LTHREADS) Dh[i] = 0.f;
  __syncthreads();

  const int npair = B * LU;
  float dc_carry[2] = {0.f, 0.f};
  const int m_tiles = Bp / 16;
  unsigned int step = 0;

  for (int t = T - 1; t >= 0; --t) {
    if (t < T - 1) {
      ++step;
      grid_wait(p.bar, gridDim.x * step, p.spin_limit);
      float acc[4][4];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
      for (int g = 0; g < 4; ++g) {
        __syncthreads();  // previous chunk fully consumed
        for (int i = tid; i < B * (Hs / 8); i += LTHREADS) {
          const int b = i / (Hs / 8), kc = i % (Hs / 8);
          *reinterpret_cast<uint4*>(Asm + b * ars + kc * 8) =
              ldcg16(p.dgates + (static_cast<long long>(b) * T + (t + 1)) * K4 + g * Hs + kc * 8);
        }
        __syncthreads();
        // split-K over warps: 16-wide k-steps dealt round-robin
        for (int ks = warp; ks < Hs / 16; ks += 8) {
          const int k0 = ks * 16;
          uint32_t b0, b1;
          ldsm_x2(smem_addr(Wt + (lane & 7) * wrs + g * Hs + k0 + ((lane >> 3) & 1) * 8), b0, b1);
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            if (mt < m_tiles) {
              uint32_t a0, a1, a2, a3;
              ldsm_x4(smem_addr(Asm + (mt * 16 + (lane & 15)) * ars + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
              mma16816(acc[mt], a0, a1, a2, a3, b0, b1);
            }
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        if (mt < m_tiles) {
          const int r = mt * 16 + (lane >> 2), c = (lane & 3) * 2;
          float* Rw = R + warp * Bp * 8;
          Rw[r * 8 + c] = acc[mt][0];
          Rw[r * 8 + c + 1] = acc[mt][1];
          Rw[(r + 8) * 8 + c] = acc[mt][2];
          Rw[(r + 8) * 8 + c + 1] = acc[mt][3];
        }
      }
      __syncthreads();
      for (int i = tid; i < Bp * 8; i += LTHREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += R[w * Bp * 8 + i];
        Dh[i] = s;
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pr = tid + i * LTHREADS;
      if (pr < npair) {
        const int b = pr >> 3, u = pr & 7;
        const long long o = (static_cast<long long>(b) * T + t) * Hs + u0 + u;
        const float* gr = p.gates + (static_cast<long long>(b) * T + t) * K4 + u0 + u;
        const float ig = gr[0], fg = gr[Hs], gg = gr[2 * Hs], og = gr[3 * Hs];
        const float c = p.c_all[o];
        const float cprev = t > 0 ? p.c_all[o - Hs] : 0.f;
        const float dh = p.dh_out[o] + (t < T - 1 ? Dh[b * 8 + u] : 0.f);
        const float tc = tanhf(c);
        const float dc = dh * og * (1.f - tc * tc) + dc_carry[i];
        dc_carry[i] = dc * fg;
        const float dai = dc * gg * ig * (1.f - ig);
        const float daf = dc * cprev * fg * (1.f - fg);
        const float dag = dc * ig * (1.f - gg * gg);
        const float dao = dh * tc * og * (1.f - og);
        __nv_bfloat16* dg = p.dgates + (static_cast<long long>(b) * T + t) * K4 + u0 + u;
        dg[0] = f2bf(dai); dg[Hs] = f2bf(daf); dg[2 * Hs] = f2bf(dag); dg[3 * Hs] = f2bf(dao);
      }
    }
    if (t > 0) grid_arrive(p.bar);
  }
}

// ------------------------------------------------------------------------------------------------ batch-sliced variant
// The utterances of a batch are independent, so the recurrence is cut into slices of RB = 8 utterances; a GROUP of G CTAs
// owns one slice and splits the hidden units (U = Hs / G per CTA, e.g. 28 of 896 with G = 32).  Compared with one
// group over the whole batch this divides the per-step all-gather (h_{t-1}, or dgates_{t+1} in the backward) by the
// number of slices and shrinks every barrier to G arrivals.  The per-step product is computed "transposed":
//   gates^T[4U, 8] = W_slice[4U, Hs] . h^T[Hs, 8]            (mma.sync m16n8k16: weights are the M operand, the 8
//   dh^T[U, 8]     = W_hh^T_slice[U, 4Hs] . dgates^T[4Hs, 8]   utterances exactly fill N -- no padded batch rows)
// with K split over the 8 warps (independent accumulators per warp -> no long dependent MMA chain) and a shared-memory
// reduction of the 8 partials.
static constexpr int RB = 8;

// Arrive on the group counter after this CTA's exchange stores: bar.sync orders every thread's stores before thread 0,
// whose single gpu-scope fence (cumulative) then publishes them -- one fence per CTA instead of one per thread, and it
// only has to wait for the few exchange stores issued so far (the bulky per-step saves are stored AFTER the arrive).
__device__ __forceinline__ void group_arrive(unsigned int* bar) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
  }
}

// Gate non-linearities on the SFU (one MUFU.EX2 + one MUFU.RCP each, |abs err| ~1e-7): libm expf/tanhf cost ~40
// instructions with range-reduction branches apiece, five times per step on the critical path of the recurrence.
__device__ __forceinline__ float sigmoid_sfu(float x) { return rcp_approx(1.f + ex2_approx(x * -1.4426950408889634f)); }
__device__ __forceinline__ float tanh_sfu(float x) { return fmaf(2.f, sigmoid_sfu(2.f * x), -1.f); }

struct LstmBsP {
  const float* xg;            // (B,T,4Hs)   fwd
  const __nv_bfloat16* whh;   // (4Hs, ldw)
  __nv_bfloat16* h_bf16;      // (B,T,Hs)    fwd out / exchange buffer
  float* h_f32;               // optional
  float* c_all;               // (B,T,Hs)
  float* gates;               // (B,T,4Hs)
  const float* dh_out;        // (B,T,Hs)    bwd
  __nv_bfloat16* dgates;      // (B,T,4Hs)   bwd out / exchange buffer
  unsigned int* bar;          // [S] group counters
  int B, T, Hs, ldw, G, U;
  int dbg;                    // timing experiments only (MTASR_LSTM_DBG): 1 skip MMA, 2 skip exchange load, 4 skip barrier
  long long spin_limit;       // grid_wait trap threshold in SM clocks (0 = never)
};

template <int KPER>
__global__ void __launch_bounds__(LTHREADS, 1) lstm_fwd_bs_kernel(const LstmBsP p) {
  extern __shared__ __align__(16) uint8_t lsm[];
  const int Hs = p.Hs, T = p.T, U = p.U, G = p.G;
  const int rs = Hs + LPAD;
  const int M = 4 * U, m_tiles = (M + 15) / 16;
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(lsm);          // [4U][rs]
  __nv_bfloat16* Hsm = Ws + M * rs;                                   // [8][rs]
  float* P = reinterpret_cast<float*>(Hsm + RB * rs);                 // [4 K-quarters][m_tiles*16][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x / G, j = blockIdx.x % G;
  const int b0 = s * RB, nb = min(RB, p.B - b0);
  const int u0 = j * U;
  unsigned int* bar = p.bar + s;

  for (int i = tid; i < M * (Hs / 8); i += LTHREADS) {
    const int lc = i / (Hs / 8), kc = i % (Hs / 8);
    const int gcol = (lc / U) * Hs + u0 + (lc % U);
    *reinterpret_cast<uint4*>(Ws + lc * rs + kc * 8) =
        *reinterpret_cast<const uint4*>(p.whh + static_cast<long long>(gcol) * p.ldw + kc * 8);
  }
  for (int i = tid; i < RB * rs; i += LTHREADS) Hsm[i] = f2bf(0.f);
  __syncthreads();

  // warp (kq, mh): K quarter kq = warp & 3, M-tile half mh = warp >> 2 (<= 4 independent accumulators per warp)
  const int ksteps = Hs / 16;
  const int kper = (ksteps + 3) / 4;
  const int kq = warp & 3, mh = warp >> 2;
  const int k_lo = kq * kper, k_hi = min(ksteps, k_lo + kper);
  const int mt_half = (m_tiles + 1) / 2;
  const int mt_lo = mh * mt_half;
  const int pr = tid;                       // (b, u) pair owned by this thread
  const bool has_pair = pr < nb * U;
  const int pb = pr / U, pu = pr % U;
  float creg = 0.f;
  const int prow = m_tiles * 16;
  float sv_h = 0.f, sv_c = 0.f, sv_i = 0.f, sv_f = 0.f, sv_g = 0.f, sv_o = 0.f;

  for (int t = 0; t < T; ++t) {
    float xv[4] = {0.f, 0.f, 0.f, 0.f};
    if (has_pair) {
      const float* xr = p.xg + (static_cast<long long>(b0 + pb) * T + t) * 4 * Hs + u0 + pu;
#pragma unroll
      for (int g = 0; g < 4; ++g) xv[g] = xr[g * Hs];
    }
    float gsum[4] = {0.f, 0.f, 0.f, 0.f};
    if (t > 0) {
      if (!(p.dbg & 4)) grid_wait(bar, static_cast<unsigned>(G) * static_cast<unsigned>(t), p.spin_limit);
      if (!(p.dbg & 2))
      for (int i = tid; i < nb * (Hs / 8); i += LTHREADS) {
        const int b = i / (Hs / 8), kc = i % (Hs / 8);
        *reinterpret_cast<uint4*>(Hsm + b * rs + kc * 8) =
            ldcg16(p.h_bf16 + (static_cast<long long>(b0 + b) * T + (t - 1)) * Hs + kc * 8);
      }
      __syncthreads();
      float acc[4][4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mi][q] = 0.f;
      if constexpr (KPER > 0) {
        // fixed trip count (Hs = 64 KPER): fully unrolled so that the ldmatrix of later k-steps are issued ahead of the
        // dependent MMA chain instead of ldmatrix -> mma -> ldmatrix ... serialisation (1.5 us of a 5 us step before)
        uint32_t hb[KPER][2];
#pragma unroll
        for (int i = 0; i < KPER; ++i)
          ldsm_x2(smem_addr(Hsm + (lane & 7) * rs + (k_lo + i) * 16 + ((lane >> 3) & 1) * 8), hb[i][0], hb[i][1]);
        if (!(p.dbg & 1)) {
#pragma unroll
          for (int i = 0; i < KPER; ++i) {
            const int k0 = (k_lo + i) * 16;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
              const int mt = mt_lo + mi;
              if (mi < mt_half && mt < m_tiles) {
                const int row = min(mt * 16 + (lane & 15), M - 1);
                uint32_t a0, a1, a2, a3;
                ldsm_x4(smem_addr(Ws + row * rs + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
                mma16816(acc[mi], a0, a1, a2, a3, hb[i][0], hb[i][1]);
              }
            }
          }
        }
      } else {
      for (int ks = (p.dbg & 1) ? k_hi : k_lo; ks < k_hi; ++ks) {
          const int k0 = ks * 16;
          uint32_t hb0, hb1;
          ldsm_x2(smem_addr(Hsm + (lane & 7) * rs + k0 + ((lane >> 3) & 1) * 8), hb0, hb1);
  #pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            const int mt = mt_lo + mi;
            if (mi < mt_half && mt < m_tiles) {
              const int row = min(mt * 16 + (lane & 15), M - 1);
              uint32_t a0, a1, a2, a3;
              ldsm_x4(smem_addr(Ws + row * rs + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
              mma16816(acc[mi], a0, a1, a2, a3, hb0, hb1);
            }
          }
        }
      }
      float* Pw = P + kq * prow * 8;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int mt = mt_lo + mi;
        if (mi < mt_half && mt < m_tiles) {
          const int r = mt * 16 + (lane >> 2), c = (lane & 3) * 2;
          *reinterpret_cast<float2*>(Pw + r * 8 + c) = make_float2(acc[mi][0], acc[mi][1]);
          *reinterpret_cast<float2*>(Pw + (r + 8) * 8 + c) = make_float2(acc[mi][2], acc[mi][3]);
        }
      }
      __syncthreads();
      if (has_pair) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int lc = g * U + pu;
          float a = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) a += P[(w * prow + lc) * 8 + pb];
          gsum[g] = a;
        }
      }
    }
    if (has_pair) {
      const float ai = xv[0] + gsum[0], af = xv[1] + gsum[1], ag = xv[2] + gsum[2], ao = xv[3] + gsum[3];
      const float ig = sigmoid_sfu(ai), fg = sigmoid_sfu(af);
      const float gg = tanh_sfu(ag), og = sigmoid_sfu(ao);
      const float c = fg * creg + ig * gg;
      creg = c;
      const float h = og * tanh_sfu(c);
      const long long o = (static_cast<long long>(b0 + pb) * T + t) * Hs + u0 + pu;
      p.h_bf16[o] = f2bf(h);                       // the exchange store: published by the arrive below
      sv_h = h; sv_c = c; sv_i = ig; sv_f = fg; sv_g = gg; sv_o = og;
    }
    if (t + 1 < T) group_arrive(bar);
    if (has_pair) {                                // saves for the backward: off the critical path
      const long long o = (static_cast<long long>(b0 + pb) * T + t) * Hs + u0 + pu;
      if (p.h_f32) p.h_f32[o] = sv_h;
      p.c_all[o] = sv_c;
      float* gr = p.gates + (static_cast<long long>(b0 + pb) * T + t) * 4 * Hs + u0 + pu;
      gr[0] = sv_i; gr[Hs] = sv_f; gr[2 * Hs] = sv_g; gr[3 * Hs] = sv_o;
    }
  }
}

template <int KPER>
__global__ void __launch_bounds__(LTHREADS, 1) lstm_bwd_bs_kernel(const LstmBsP p) {
  extern __shared__ __align__(16) uint8_t lsm[];
  const int Hs = p.Hs, T = p.T, U = p.U, G = p.G;
  const int K4 = 4 * Hs;
  const int wrs = K4 + LPAD, ars = Hs + LPAD;
  const int m_tiles = (U + 15) / 16;        // <= 2
  __nv_bfloat16* Wt = reinterpret_cast<__nv_bfloat16*>(lsm);        // [U][wrs]: Wt[u][col] = whh[col][u0+u]
  __nv_bfloat16* Asm = Wt + U * wrs;                                // [8][ars] one Hs-wide chunk of the dgates slice
  float* P = reinterpret_cast<float*>(Asm + RB * ars);              // [8 warps][m_tiles*16][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s = blockIdx.x / G, j = blockIdx.x % G;
  const int b0 = s * RB, nb = min(RB, p.B - b0);
  const int u0 = j * U;
  unsigned int* bar = p.bar + s;

  for (int i = tid; i < U * K4; i += LTHREADS) {
    const int u = i % U, col = i / U;
    Wt[u * wrs + col] = p.whh[static_cast<long long>(col) * p.ldw + u0 + u];
  }
  for (int i = tid; i < RB * ars; i += LTHREADS) Asm[i] = f2bf(0.f);
  __syncthreads();

  const int ksteps = Hs / 16;
  const int kper = (ksteps + 7) / 8;
  const int k_lo = warp * kper, k_hi = min(ksteps, k_lo + kper);
  const int pr = tid;
  const bool has_pair = pr < nb * U;
  const int pb = pr / U, pu = pr % U;
  const int prow = m_tiles * 16;
  float dc_carry = 0.f;
  unsigned int step = 0;

  const int nload = nb * (Hs / 8);               // 16-byte pieces of one Hs-wide chunk of the dgates slice
  for (int t = T - 1; t >= 0; --t) {
    float dh_rec = 0.f;
    // this step's saved activations / upstream gradient: issued before the barrier wait so HBM latency is hidden
    float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, cv = 0.f, cprev = 0.f, dho = 0.f;
    if (has_pair) {
      const long long o = (static_cast<long long>(b0 + pb) * T + t) * Hs + u0 + pu;
      const float* gr = p.gates + (static_cast<long long>(b0 + pb) * T + t) * K4 + u0 + pu;
      ig = gr[0]; fg = gr[Hs]; gg = gr[2 * Hs]; og = gr[3 * Hs];
      cv = p.c_all[o];
      cprev = t > 0 ? p.c_all[o - Hs] : 0.f;
      dho = p.dh_out[o];
    }
    if (t < T - 1) {
      ++step;
      grid_wait(bar, static_cast<unsigned>(G) * step, p.spin_limit);
      // all four chunks of dgates_{t+1} are requested at once (registers) so the L2 latency is paid once per step
      uint4 stage[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int i = tid + v * LTHREADS;
          if (i < nload) {
            const int b = i / (Hs / 8), kc = i % (Hs / 8);
            stage[g][v] = ldcg16(p.dgates + (static_cast<long long>(b0 + b) * T + (t + 1)) * K4 + g * Hs + kc * 8);
          }
        }
      float acc[2][2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[a][b][q] = 0.f;
      for (int g = 0; g < 4; ++g) {
        __syncthreads();   // previous chunk fully consumed
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int i = tid + v * LTHREADS;
          if (i < nload) {
            const int b = i / (Hs / 8), kc = i % (Hs / 8);
            *reinterpret_cast<uint4*>(Asm + b * ars + kc * 8) = stage[g][v];
          }
        }
        __syncthreads();
        if constexpr (KPER > 0) {
          // fixed trip count (Hs = 128 KPER): unrolled, B fragments of the whole chunk loaded ahead of the MMA chain
          uint32_t db[KPER][2];
#pragma unroll
          for (int i = 0; i < KPER; ++i)
            ldsm_x2(smem_addr(Asm + (lane & 7) * ars + (k_lo + i) * 16 + ((lane >> 3) & 1) * 8), db[i][0], db[i][1]);
#pragma unroll
          for (int i = 0; i < KPER; ++i) {
            const int k0 = (k_lo + i) * 16;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              if (mt < m_tiles) {
                const int row = min(mt * 16 + (lane & 15), U - 1);
                uint32_t a0, a1, a2, a3;
                ldsm_x4(smem_addr(Wt + row * wrs + g * Hs + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
                mma16816(acc[mt][i & 1], a0, a1, a2, a3, db[i][0], db[i][1]);
              }
            }
          }
        } else {
          for (int ks = k_lo; ks < k_hi; ks += 2) {
#pragma unroll
            for (int par = 0; par < 2; ++par) {     // two independent accumulator sets halve the dependent MMA chain
              if (ks + par < k_hi) {
                const int k0 = (ks + par) * 16;
                uint32_t d0, d1;
                ldsm_x2(smem_addr(Asm + (lane & 7) * ars + k0 + ((lane >> 3) & 1) * 8), d0, d1);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                  if (mt < m_tiles) {
                    const int row = min(mt * 16 + (lane & 15), U - 1);
                    uint32_t a0, a1, a2, a3;
                    ldsm_x4(smem_addr(Wt + row * wrs + g * Hs + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
                    mma16816(acc[mt][par], a0, a1, a2, a3, d0, d1);
                  }
                }
              }
            }
          }
        }
      }
      float* Pw = P + warp * prow * 8;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt < m_tiles) {
          const int r = mt * 16 + (lane >> 2), c = (lane & 3) * 2;
          *reinterpret_cast<float2*>(Pw + r * 8 + c) = make_float2(acc[mt][0][0] + acc[mt][1][0], acc[mt][0][1] + acc[mt][1][1]);
          *reinterpret_cast<float2*>(Pw + (r + 8) * 8 + c) =
              make_float2(acc[mt][0][2] + acc[mt][1][2], acc[mt][0][3] + acc[mt][1][3]);
        }
      }
      __syncthreads();
      if (has_pair) {
#pragma unroll
        for (int w = 0; w < 8; ++w) dh_rec += P[(w * prow + pu) * 8 + pb];
      }
    }
    if (has_pair) {
      const float dh = dho + dh_rec;
      const float tc = tanhf(cv);
      const float dc = dh * og * (1.f - tc * tc) + dc_carry;
      dc_carry = dc * fg;
      const float dai = dc * gg * ig * (1.f - ig);
      const float daf = dc * cprev * fg * (1.f - fg);
      const float dag = dc * ig * (1.f - gg * gg);
      const float dao = dh * tc * og * (1.f - og);
      __nv_bfloat16* dg = p.dgates + (static_cast<long long>(b0 + pb) * T + t) * K4 + u0 + pu;
      dg[0] = f2bf(dai); dg[Hs] = f2bf(daf); dg[2 * Hs] = f2bf(dag); dg[3 * Hs] = f2bf(dao);
    }
    if (t > 0) group_arrive(bar);
  }
}

// ------------------------------------------------------------------------------------------------ flag-in-data exchange
// The batch-sliced kernels above pay three dependent L2 round trips per step: exchange store -> fence -> atomic arrive,
// the waiter's poll of the counter, then the exchange loads -- plus five CTA-wide barriers around the smem staging.  The
// *_ll kernels below exchange h_t (forward) / dgates_t (backward) the way NCCL's LL protocol moves small messages: every
// 8-byte word of the exchange buffer is {two bf16 values, 32-bit step flag}, written by ONE 64-bit store and read by
// 64-bit loads (single-copy atomic), so a consumer that sees the flag of the step it waits for has the data -- no fence,
// no counter, ONE L2 round trip per step.  The words are laid out so that one 16-byte load is exactly the {b0, b1}
// B-fragment pair of an mma.sync m16n8k16 lane (units 16 ks + 2 j + {0,1} and 16 ks + 8 + 2 j + {0,1} of utterance
// lane / 4): the consumer warps poll their own fragments straight from L2 into registers -- no shared-memory staging, no
// barrier in front of the MMAs -- and the per-warp partial products are double-buffered, which leaves ONE __syncthreads
// per step.  The buffer holds two step parities (a CTA can run at most one step ahead of the slowest CTA of its group).
__device__ __forceinline__ void ll_store(unsigned long long* p, uint32_t data, uint32_t flag) {
  const unsigned long long v = static_cast<unsigned long long>(data) | (static_cast<unsigned long long>(flag) << 32);
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ll_load2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// Spin until both words of the fragment pair carry `flag`; -> {b0, b1}.  A waiter that spins past `limit` clocks traps.
__device__ __forceinline__ void ll_wait2(const unsigned long long* p, unsigned long long& a, unsigned long long& b, uint32_t flag,
                                         long long limit) {
  if (static_cast<uint32_t>(a >> 32) == flag && static_cast<uint32_t>(b >> 32) == flag) return;
  const long long t0 = clock64();
  unsigned int spins = 0;
  do {
    ll_load2(p, a, b);
    if ((++spins & 0x3ff) == 0 && limit > 0 && clock64() - t0 > limit) __trap();
  } while (static_cast<uint32_t>(a >> 32) != flag || static_cast<uint32_t>(b >> 32) != flag);
}
// Batched wait over N fragment pairs at p + i * stride (64-bit words): every pair that does not carry `flag` yet is
// re-requested in the SAME round, so a round costs one L2 round trip however many pairs were early (waiting for them one
// after the other serialised up to N round trips per step).
template <int N>
__device__ __forceinline__ void ll_wait_all(const unsigned long long* p, int stride, unsigned long long (&a)[N],
                                            unsigned long long (&b)[N], uint32_t flag, long long limit) {
  long long t0 = 0;
  unsigned int spins = 0;
  while (true) {
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (static_cast<uint32_t>(a[i] >> 32) != flag || static_cast<uint32_t>(b[i] >> 32) != flag) {
        ok = false;
        ll_load2(p + i * stride, a[i], b[i]);
      }
    }
    if (ok) return;
    if (spins == 0) t0 = clock64();
    if ((++spins & 0x3ff) == 0 && limit > 0 && clock64() - t0 > limit) __trap();
  }
}
// 64-bit word index of unit pair (k, k+1), k even, inside one utterance's row of `ksteps` 16-unit k-steps:
// [ks][j = (k % 8) / 2][half = (k % 16) / 8]
__device__ __forceinline__ int ll_word(int k) { return (k >> 4) * 8 + ((k & 7) >> 1) * 2 + ((k >> 3) & 1); }

struct LstmLlP {
  const float* xg;            // (B,T,4Hs)   fwd
  const __nv_bfloat16* whh;   // (4Hs, ldw)
  __nv_bfloat16* h_bf16;      // (B,T,Hs)    fwd out
  float* h_f32;               // optional
  float* c_all;               // (B,T,Hs)
  float* gates;               // (B,T,4Hs)
  const float* dh_out;        // (B,T,Hs)    bwd
  __nv_bfloat16* dgates;      // (B,T,4Hs)   bwd out
  unsigned long long* ll;     // exchange words: [2 parities][S][8 utterances][Hs/2 (fwd) | 4Hs/2 (bwd)]
  int B, T, Hs, ldw, G, U;
  long long spin_limit;
};

// KPER > 0: every warp owns exactly KPER k-steps (fully unrolled, all fragment loads of a step in flight at once).
// KREG > 0: the weight fragments (MMA A operands) of the warp's first KREG k-steps stay in REGISTERS for all T steps.  The
// per-step product is bound by streaming the CTA's 200 KB weight slice out of shared memory through ldmatrix (0.8 us at
// 128 B/clk); the register file (256 KB per SM, one CTA per SM) is otherwise idle, so part of the slice lives there.
template <int KPER, int KREG = 0>
__global__ void __launch_bounds__(LTHREADS, 1) lstm_fwd_ll_kernel(const LstmLlP p) {
  extern __shared__ __align__(16) uint8_t lsm[];
  const int Hs = p.Hs, T = p.T, U = p.U, G = p.G;
  const int rs = Hs + LPAD;
  const int M = 4 * U, m_tiles = (M + 15) / 16;
  const int prow = m_tiles * 16;
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(lsm);          // [4U][rs]
  float* P = reinterpret_cast<float*>(Ws + M * rs);                   // [2 step parities][4 K-quarters][prow][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = gridDim.x / G;
  const int s = blockIdx.x / G, j = blockIdx.x % G;
  const int b0 = s * RB, nb = min(RB, p.B - b0);
  const int u0 = j * U;

  for (int i = tid; i < M * (Hs / 8); i += LTHREADS) {
    const int lc = i / (Hs / 8), kc = i % (Hs / 8);
    const int gcol = (lc / U) * Hs + u0 + (lc % U);
    *reinterpret_cast<uint4*>(Ws + lc * rs + kc * 8) =
        *reinterpret_cast<const uint4*>(p.whh + static_cast<long long>(gcol) * p.ldw + kc * 8);
  }
  __syncthreads();

  const int ksteps = Hs / 16;
  const int kper = (ksteps + 3) / 4;
  const int kq = warp & 3, mh = warp >> 2;
  const int k_lo = kq * kper, k_hi = min(ksteps, k_lo + kper);
  const int mt_half = (m_tiles + 1) / 2;
  const int mt_lo = mh * mt_half;
  const int pr = tid;                       // (b, u) pair owned by this thread
  const bool has_pair = pr < nb * U;
  const int pb = has_pair ? pr / U : 0, pu = has_pair ? pr % U : 0;
  float creg = 0.f;
  const int wpr = Hs / 2;                   // exchange words per utterance
  const bool frag_ok = (lane >> 2) < nb;    // utterance (MMA column) lane / 4 exists
  // this lane's fragment pair of k-step ks lives at frag0 + parity * S * RB * wpr + ks * 8
  const unsigned long long* frag0 = p.ll + (static_cast<long long>(s) * RB + (lane >> 2)) * wpr + (lane & 3) * 2;
  unsigned long long* const out0 = p.ll + (static_cast<long long>(s) * RB + pb) * wpr + ll_word(u0 + pu);
  const long long par_stride = static_cast<long long>(S) * RB * wpr;

  // register-resident weight fragments: k-steps k_lo .. k_lo + KREG - 1 of this warp's (up to) four m-tiles
  uint32_t wreg[KREG > 0 ? KREG : 1][4][4];
  if constexpr (KPER > 0 && KREG > 0) {
#pragma unroll
    for (int i = 0; i < KREG; ++i) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int mt = mt_lo + mi;
        wreg[i][mi][0] = wreg[i][mi][1] = wreg[i][mi][2] = wreg[i][mi][3] = 0u;
        if (mi < mt_half && mt < m_tiles) {
          const int row = min(mt * 16 + (lane & 15), M - 1);
          ldsm_x4(smem_addr(Ws + row * rs + (k_lo + i) * 16 + (lane >> 4) * 8), wreg[i][mi][0], wreg[i][mi][1], wreg[i][mi][2],
                  wreg[i][mi][3]);
        }
      }
    }
  }

  for (int t = 0; t < T; ++t) {
    float xv[4] = {0.f, 0.f, 0.f, 0.f};
    if (has_pair) {
      const float* xr = p.xg + (static_cast<long long>(b0 + pb) * T + t) * 4 * Hs + u0 + pu;
#pragma unroll
      for (int g = 0; g < 4; ++g) xv[g] = xr[g * Hs];
    }
    float gsum[4] = {0.f, 0.f, 0.f, 0.f};
    if (t > 0) {
      const uint32_t flag = static_cast<uint32_t>(t);                 // h_{t-1} was published with flag (t-1) + 1
      const unsigned long long* fr = frag0 + ((t - 1) & 1) * par_stride;
      float acc[4][4];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mi][q] = 0.f;
      if constexpr (KPER > 0) {
        unsigned long long fa[KPER], fb[KPER];
#pragma unroll
        for (int i = 0; i < KPER; ++i) {
          fa[i] = fb[i] = 0ull;
          if (frag_ok) ll_load2(fr + (k_lo + i) * 8, fa[i], fb[i]);
        }
        if (frag_ok) ll_wait_all<KPER>(fr + k_lo * 8, 8, fa, fb, flag, p.spin_limit);
#pragma unroll
        for (int i = 0; i < KPER; ++i) {
          const int k0 = (k_lo + i) * 16;
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            const int mt = mt_lo + mi;
            if (mi < mt_half && mt < m_tiles) {
              if (i < KREG) {
                mma16816(acc[mi], wreg[i][mi][0], wreg[i][mi][1], wreg[i][mi][2], wreg[i][mi][3], static_cast<uint32_t>(fa[i]),
                         static_cast<uint32_t>(fb[i]));
              } else {
                const int row = min(mt * 16 + (lane & 15), M - 1);
                uint32_t a0, a1, a2, a3;
                ldsm_x4(smem_addr(Ws + row * rs + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
                mma16816(acc[mi], a0, a1, a2, a3, static_cast<uint32_t>(fa[i]), static_cast<uint32_t>(fb[i]));
              }
            }
          }
        }
      } else {
        for (int ks = k_lo; ks < k_hi; ++ks) {
          unsigned long long fa = 0ull, fb = 0ull;
          if (frag_ok) {
            ll_load2(fr + ks * 8, fa, fb);
            ll_wait2(fr + ks * 8, fa, fb, flag, p.spin_limit);
          }
          const int k0 = ks * 16;
#pragma unroll
          for (int mi = 0; mi < 4; ++mi) {
            const int mt = mt_lo + mi;
            if (mi < mt_half && mt < m_tiles) {
              const int row = min(mt * 16 + (lane & 15), M - 1);
              uint32_t a0, a1, a2, a3;
              ldsm_x4(smem_addr(Ws + row * rs + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
              mma16816(acc[mi], a0, a1, a2, a3, static_cast<uint32_t>(fa), static_cast<uint32_t>(fb));
            }
          }
        }
      }
      float* Pt = P + (t & 1) * 4 * prow * 8;
      float* Pw = Pt + kq * prow * 8;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int mt = mt_lo + mi;
        if (mi < mt_half && mt < m_tiles) {
          const int r = mt * 16 + (lane >> 2), c = (lane & 3) * 2;
          *reinterpret_cast<float2*>(Pw + r * 8 + c) = make_float2(acc[mi][0], acc[mi][1]);
          *reinterpret_cast<float2*>(Pw + (r + 8) * 8 + c) = make_float2(acc[mi][2], acc[mi][3]);
        }
      }
      __syncthreads();
      if (has_pair) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int lc = g * U + pu;
          float a = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) a += Pt[(w * prow + lc) * 8 + pb];
          gsum[g] = a;
        }
      }
    }
    const float ai = xv[0] + gsum[0], af = xv[1] + gsum[1], ag = xv[2] + gsum[2], ao = xv[3] + gsum[3];
    const float ig = sigmoid_sfu(ai), fg = sigmoid_sfu(af);
    const float gg = tanh_sfu(ag), og = sigmoid_sfu(ao);
    const float c = fg * creg + ig * gg;
    creg = c;
    const float h = has_pair ? og * tanh_sfu(c) : 0.f;
    const __nv_bfloat16 hb = f2bf(h);
    // exchange: the even unit of a pair publishes {h_u, h_{u+1}, flag}; U is even, so the partner is the next lane
    const uint32_t mine = static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(&hb));
    const uint32_t next = __shfl_down_sync(0xffffffffu, mine, 1);
    if (has_pair && !(pu & 1) && t + 1 < T) ll_store(out0 + (t & 1) * par_stride, mine | (next << 16), static_cast<uint32_t>(t + 1));
    if (has_pair) {                                // outputs / saves for the backward: off the critical path
      const long long o = (static_cast<long long>(b0 + pb) * T + t) * Hs + u0 + pu;
      p.h_bf16[o] = hb;
      if (p.h_f32) p.h_f32[o] = h;
      p.c_all[o] = c;
      float* gr = p.gates + (static_cast<long long>(b0 + pb) * T + t) * 4 * Hs + u0 + pu;
      gr[0] = ig; gr[Hs] = fg; gr[2 * Hs] = gg; gr[3 * Hs] = og;
    }
  }
}

template <int KPER>
__global__ void __launch_bounds__(LTHREADS, 1) lstm_bwd_ll_kernel(const LstmLlP p) {
  extern __shared__ __align__(16) uint8_t lsm[];
  const int Hs = p.Hs, T = p.T, U = p.U, G = p.G;
  const int K4 = 4 * Hs;
  const int wrs = K4 + LPAD;
  const int m_tiles = (U + 15) / 16;        // <= 2
  const int prow = m_tiles * 16;
  __nv_bfloat16* Wt = reinterpret_cast<__nv_bfloat16*>(lsm);        // [U][wrs]: Wt[u][col] = whh[col][u0+u]
  float* P = reinterpret_cast<float*>(Wt + U * wrs);                // [2 step parities][8 warps][prow][8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = gridDim.x / G;
  const int s = blockIdx.x / G, j = blockIdx.x % G;
  const int b0 = s * RB, nb = min(RB, p.B - b0);
  const int u0 = j * U;

  for (int i = tid; i < U * K4; i += LTHREADS) {
    const int u = i % U, col = i / U;
    Wt[u * wrs + col] = p.whh[static_cast<long long>(col) * p.ldw + u0 + u];
  }
  __syncthreads();

  const int ksteps = Hs / 16;
  const int kper = (ksteps + 7) / 8;
  const int k_lo = warp * kper, k_hi = min(ksteps, k_lo + kper);
  const int pr = tid;
  const bool has_pair = pr < nb * U;
  const int pb = has_pair ? pr / U : 0, pu = has_pair ? pr % U : 0;
  float dc_carry = 0.f;
  const int wpr = K4 / 2;                   // exchange words per utterance: [gate][Hs / 2]
  const bool frag_ok = (lane >> 2) < nb;
  const unsigned long long* frag0 = p.ll + (static_cast<long long>(s) * RB + (lane >> 2)) * wpr + (lane & 3) * 2;
  unsigned long long* const out0 = p.ll + (static_cast<long long>(s) * RB + pb) * wpr + ll_word(u0 + pu);
  const long long par_stride = static_cast<long long>(S) * RB * wpr;
  unsigned int step = 0;

  for (int t = T - 1; t >= 0; --t, ++step) {
    float dh_rec = 0.f;
    // this step's saved activations / upstream gradient: issued before the exchange wait so HBM latency is hidden
    float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, cv = 0.f, cprev = 0.f, dho = 0.f;
    if (has_pair) {
      const long long o = (static_cast<long long>(b0 + pb) * T + t) * Hs + u0 + pu;
      const float* gr = p.gates + (static_cast<long long>(b0 + pb) * T + t) * K4 + u0 + pu;
      ig = gr[0]; fg = gr[Hs]; gg = gr[2 * Hs]; og = gr[3 * Hs];
      cv = p.c_all[o];
      cprev = t > 0 ? p.c_all[o - Hs] : 0.f;
      dho = p.dh_out[o];
    }
    if (step > 0) {
      const uint32_t flag = step;                                     // dgates_{t+1} were published with flag (step-1) + 1
      const unsigned long long* fr = frag0 + ((step - 1) & 1) * par_stride;
      float acc[2][2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[a][b][q] = 0.f;
      if constexpr (KPER > 0) {
        // gate chunk g + 1's fragments are requested before chunk g's MMAs (requesting chunks 1..3 all at once after chunk 0
        // has arrived was measured slower: 4.88 -> 5.09 us per step)
        unsigned long long fa[2][KPER], fb[2][KPER];
#pragma unroll
        for (int i = 0; i < KPER; ++i) {
          fa[0][i] = fb[0][i] = 0ull;
          if (frag_ok) ll_load2(fr + (k_lo + i) * 8, fa[0][i], fb[0][i]);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cur = g & 1;
          if (g + 1 < 4) {
#pragma unroll
            for (int i = 0; i < KPER; ++i) {
              fa[cur ^ 1][i] = fb[cur ^ 1][i] = 0ull;
              if (frag_ok) ll_load2(fr + (g + 1) * (Hs / 2) + (k_lo + i) * 8, fa[cur ^ 1][i], fb[cur ^ 1][i]);
            }
          }
          if (frag_ok) ll_wait_all<KPER>(fr + g * (Hs / 2) + k_lo * 8, 8, fa[cur], fb[cur], flag, p.spin_limit);
#pragma unroll
          for (int i = 0; i < KPER; ++i) {
            const int k0 = (k_lo + i) * 16;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              if (mt < m_tiles) {
                const int row = min(mt * 16 + (lane & 15), U - 1);
                uint32_t a0, a1, a2, a3;
                ldsm_x4(smem_addr(Wt + row * wrs + g * Hs + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
                mma16816(acc[mt][i & 1], a0, a1, a2, a3, static_cast<uint32_t>(fa[cur][i]), static_cast<uint32_t>(fb[cur][i]));
              }
            }
          }
        }
      } else {
        for (int g = 0; g < 4; ++g) {
          for (int ks = k_lo; ks < k_hi; ++ks) {
            unsigned long long fa = 0ull, fb = 0ull;
            if (frag_ok) {
              ll_load2(fr + g * (Hs / 2) + ks * 8, fa, fb);
              ll_wait2(fr + g * (Hs / 2) + ks * 8, fa, fb, flag, p.spin_limit);
            }
            const int k0 = ks * 16;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              if (mt < m_tiles) {
                const int row = min(mt * 16 + (lane & 15), U - 1);
                uint32_t a0, a1, a2, a3;
                ldsm_x4(smem_addr(Wt + row * wrs + g * Hs + k0 + (lane >> 4) * 8), a0, a1, a2, a3);
                mma16816(acc[mt][0], a0, a1, a2, a3, static_cast<uint32_t>(fa), static_cast<uint32_t>(fb));
              }
            }
          }
        }
      }
      float* Pt = P + (step & 1) * 8 * prow * 8;
      float* Pw = Pt + warp * prow * 8;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (mt < m_tiles) {
          const int r = mt * 16 + (lane >> 2), c = (lane & 3) * 2;
          *reinterpret_cast<float2*>(Pw + r * 8 + c) = make_float2(acc[mt][0][0] + acc[mt][1][0], acc[mt][0][1] + acc[mt][1][1]);
          *reinterpret_cast<float2*>(Pw + (r + 8) * 8 + c) =
              make_float2(acc[mt][0][2] + acc[mt][1][2], acc[mt][0][3] + acc[mt][1][3]);
        }
      }
      __syncthreads();
      if (has_pair) {
#pragma unroll
        for (int w = 0; w < 8; ++w) dh_rec += Pt[(w * prow + pu) * 8 + pb];
      }
    }
    const float dh = dho + dh_rec;
    const float tc = tanhf(cv);
    const float dc = dh * og * (1.f - tc * tc) + dc_carry;
    dc_carry = dc * fg;
    const float dg4[4] = {dc * gg * ig * (1.f - ig), dc * cprev * fg * (1.f - fg), dc * ig * (1.f - gg * gg), dh * tc * og * (1.f - og)};
    uint32_t mine[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const __nv_bfloat16 v = f2bf(has_pair ? dg4[g] : 0.f);
      mine[g] = static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(&v));
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t next = __shfl_down_sync(0xffffffffu, mine[g], 1);
      if (has_pair && !(pu & 1) && t > 0)
        ll_store(out0 + (step & 1) * par_stride + g * (Hs / 2), mine[g] | (next << 16), step + 1);
    }
    if (has_pair) {                                // full dgates tensor for the dW / dx GEMMs: off the critical path
      unsigned short* dg = reinterpret_cast<unsigned short*>(p.dgates) + (static_cast<long long>(b0 + pb) * T + t) * K4 + u0 + pu;
      dg[0] = static_cast<unsigned short>(mine[0]); dg[Hs] = static_cast<unsigned short>(mine[1]);
      dg[2 * Hs] = static_cast<unsigned short>(mine[2]); dg[3 * Hs] = static_cast<unsigned short>(mine[3]);
    }
  }
}

// Pick the group size G (CTAs per 8-utterance slice): U = Hs / G hidden units per CTA must be integral, the 8 x U
// (utterance, unit) pairs must fit one thread each, the weight slice must fit shared memory and all S * G CTAs must be
// co-resident.  Returns 0 when the batch-sliced kernels do not apply (the whole-batch kernels are used instead).
static int lstm_bs_pick(int B, int Hs, bool bwd, size_t* smem_out) {
  if (Hs % 16 != 0) return 0;
  const int S = (B + RB - 1) / RB;
  for (int G = 64; G >= 1; --G) {
    if (Hs % G != 0 || S * G > num_sms()) continue;
    const int U = Hs / G;
    if (RB * U > LTHREADS || RB * (Hs / 8) > 4 * LTHREADS) continue;
    size_t smem;
    if (!bwd) {
      const int M = 4 * U, mt = (M + 15) / 16;
      if (mt > 8) continue;
      smem = static_cast<size_t>(M + RB) * (Hs + LPAD) * 2 + static_cast<size_t>(4) * mt * 16 * 8 * 4;
    } else {
      const int mt = (U + 15) / 16;
      if (mt > 2) continue;
      smem = static_cast<size_t>(U) * (4 * Hs + LPAD) * 2 + static_cast<size_t>(RB) * (Hs + LPAD) * 2 +
             static_cast<size_t>(8) * mt * 16 * 8 * 4;
    }
    if (smem > 232448) continue;
    *smem_out = smem;
    return G;
  }
  return 0;
}

// MTASR_LSTM_WREG: 0 = all weight fragments of the forward from shared memory, 1 / 2 (default) = 6 / 8 of every warp's 14
// k-steps register-resident (measured at cfg2: 4.36 -> 4.08 -> 3.98 us per step; the same trick does nothing for the
// backward -- 4.98 us per step either way, it waits on the per-gate-chunk exchange, not on ldmatrix -- and is not applied there).
static int lstm_wreg_mode() {
  const char* e = getenv("MTASR_LSTM_WREG");
  return e ? atoi(e) : 2;
}

// The flag-in-data kernels apply when the batch-sliced partition exists, U is even (unit pairs never straddle CTAs) and
// their shared-memory layout (weights + double-buffered partials, no staging tile) fits.
static int lstm_ll_pick(int B, int Hs, bool bwd, size_t* smem_out) {
  if (getenv("MTASR_LSTM_NO_LL")) return 0;
  size_t dummy = 0;
  const int G = lstm_bs_pick(B, Hs, bwd, &dummy);
  if (G <= 0) return 0;
  const int U = Hs / G;
  if (U % 2 != 0) return 0;
  size_t smem;
  if (!bwd) {
    const int M = 4 * U, mt = (M + 15) / 16;
    smem = static_cast<size_t>(M) * (Hs + LPAD) * 2 + static_cast<size_t>(2) * 4 * mt * 16 * 8 * 4;
  } else {
    const int mt = (U + 15) / 16;
    smem = static_cast<size_t>(U) * (4 * Hs + LPAD) * 2 + static_cast<size_t>(2) * 8 * mt * 16 * 8 * 4;
  }
  if (smem > 232448) return 0;
  *smem_out = smem;
  return G;
}
static constexpr size_t LL_OFFSET = 256;   // the group counters of the counter-based kernels live in front of the words
static size_t lstm_ll_bytes(int B, int Hs, bool bwd) {
  const size_t S = (B + RB - 1) / RB;
  return static_cast<size_t>(2) * S * RB * (static_cast<size_t>(bwd ? 4 : 1) * Hs / 2) * 8;
}

static size_t lstm_fwd_smem(int Hs, int Bp) {
  return static_cast<size_t>(32 + Bp) * (Hs + LPAD) * 2 + static_cast<size_t>(Bp) * 33 * 4;
}
static size_t lstm_bwd_smem(int Hs, int Bp) {
  return static_cast<size_t>(LU) * (4 * Hs + LPAD) * 2 + static_cast<size_t>(Bp) * (Hs + LPAD) * 2 +
         static_cast<size_t>(9) * Bp * 8 * 4;
}

// grid_wait trap threshold in SM clocks from MTASR_LSTM_SPIN_TIMEOUT_MS (default 2000 ms at ~2 GHz; 0 disables the trap).
static long long lstm_spin_limit() {
  static long long cycles = -1;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* e = getenv("MTASR_LSTM_SPIN_TIMEOUT_MS");
    cycles = e ? static_cast<long long>(atof(e) * 2.0e6) : 4000000000LL;
  });
  return cycles;
}

static int lstm_check(int B, int T, int Hs, int ldw, const char* who) {
  if (B <= 0 || B > 64) return set_error(MTASR_ERR_UNSUPPORTED, "%s: batch %d not in [1,64] (chunk the batch)", who, B);
  if (T <= 0 || Hs <= 0 || Hs % 16 != 0)
    return set_error(MTASR_ERR_UNSUPPORTED, "%s: need T > 0 and Hs %% 16 == 0 (Hs=%d)", who, Hs);
  if (Hs / LU > num_sms())
    return set_error(MTASR_ERR_UNSUPPORTED, "%s: Hs=%d needs %d co-resident CTAs > %d SMs", who, Hs, Hs / LU, num_sms());
  if (ldw % 8 != 0) return set_error(MTASR_ERR_INVALID_ARG, "%s: ldw must be a multiple of 8", who);
  return 0;
}

}  // namespace mtasr

using namespace mtasr;

extern "C" int64_t mtasr_lstm_scratch_bytes(int32_t B, int32_t Hs, int32_t backward) {
  if (B <= 0 || Hs <= 0) return static_cast<int64_t>(LL_OFFSET);
  return static_cast<int64_t>(LL_OFFSET + lstm_ll_bytes(B, Hs, backward != 0));
}

extern "C" int mtasr_lstm_fwd(const float* xg, const void* whh_bf16, int32_t ldw, int32_t B, int32_t T, int32_t Hs,
                              void* h_bf16, float* h_f32, float* c_all, float* gates, uint32_t* barrier, void* stream) {
  MTASR_CHECK_ARG(xg && whh_bf16 && h_bf16 && c_all && gates && barrier, "lstm_fwd: null pointer");
  MTASR_CHECK_ARG((reinterpret_cast<uintptr_t>(barrier) & 15) == 0, "lstm_fwd: scratch must be 16-byte aligned");
  cudaStream_t st0 = static_cast<cudaStream_t>(stream);
  {
    size_t smem_ll = 0;
    const int G = lstm_ll_pick(B, Hs, false, &smem_ll);
    if (G > 0 && T > 0 && ldw % 8 == 0) {
      const int S = (B + RB - 1) / RB;
      LstmLlP q{};
      q.xg = xg; q.whh = reinterpret_cast<const __nv_bfloat16*>(whh_bf16); q.h_bf16 = reinterpret_cast<__nv_bfloat16*>(h_bf16);
      q.h_f32 = h_f32; q.c_all = c_all; q.gates = gates;
      q.ll = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(barrier) + LL_OFFSET);
      q.B = B; q.T = T; q.Hs = Hs; q.ldw = ldw; q.G = G; q.U = Hs / G;
      q.spin_limit = lstm_spin_limit();
      const int wreg = lstm_wreg_mode();
      void (*kern)(const LstmLlP) = (Hs == 896) ? (wreg == 0 ? lstm_fwd_ll_kernel<14> : wreg == 1 ? lstm_fwd_ll_kernel<14, 6> : lstm_fwd_ll_kernel<14, 8>)
                                                : lstm_fwd_ll_kernel<0>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_ll)) != cudaSuccess)
        return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: cannot set smem attribute");
      if (cudaMemsetAsync(q.ll, 0, lstm_ll_bytes(B, Hs, false), st0) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: memset failed");
      void* args[] = {&q};
      cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(S * G), dim3(LTHREADS), args, smem_ll, st0);
      MTASR_COUNT_LAUNCH();
      if (e != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: cooperative launch failed: %s", cudaGetErrorString(e));
      return MTASR_OK;
    }
  }
  {
    size_t smem_bs = 0;
    const int G = getenv("MTASR_LSTM_WHOLE_BATCH") ? 0 : lstm_bs_pick(B, Hs, false, &smem_bs);
    if (G > 0 && T > 0 && ldw % 8 == 0) {
      const int S = (B + RB - 1) / RB;
      LstmBsP q{};
      q.xg = xg; q.whh = reinterpret_cast<const __nv_bfloat16*>(whh_bf16); q.h_bf16 = reinterpret_cast<__nv_bfloat16*>(h_bf16);
      q.h_f32 = h_f32; q.c_all = c_all; q.gates = gates; q.bar = barrier;
      q.B = B; q.T = T; q.Hs = Hs; q.ldw = ldw; q.G = G; q.U = Hs / G;
      q.dbg = getenv("MTASR_LSTM_DBG") ? atoi(getenv("MTASR_LSTM_DBG")) : 0;
      q.spin_limit = lstm_spin_limit();
      // unrolled instantiation when every warp owns exactly Hs/64 k-steps (Hs = 896 -> 14), generic loop otherwise
      void (*kern)(const LstmBsP) = (Hs == 896) ? lstm_fwd_bs_kernel<14> : lstm_fwd_bs_kernel<0>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bs)) != cudaSuccess)
        return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: cannot set smem attribute");
      if (cudaMemsetAsync(barrier, 0, sizeof(uint32_t) * S, st0) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: memset failed");
      void* args[] = {&q};
      cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(S * G), dim3(LTHREADS), args, smem_bs, st0);
      MTASR_COUNT_LAUNCH();
      if (e != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: cooperative launch failed: %s", cudaGetErrorString(e));
      return MTASR_OK;
    }
  }
  if (int rc = lstm_check(B, T, Hs, ldw, "lstm_fwd")) return rc;
  LstmFwdP p;
  p.spin_limit = lstm_spin_limit();
  p.xg = xg; p.whh = reinterpret_cast<const __nv_bfloat16*>(whh_bf16); p.h_bf16 = reinterpret_cast<__nv_bfloat16*>(h_bf16);
  p.h_f32 = h_f32; p.c_all = c_all; p.gates = gates; p.bar = barrier;
  p.B = B; p.T = T; p.Hs = Hs; p.ldw = ldw; p.Bp = (B + 15) / 16 * 16;
  const size_t smem = lstm_fwd_smem(Hs, p.Bp);
  if (smem > 227 * 1024) return set_error(MTASR_ERR_UNSUPPORTED, "lstm_fwd: Hs=%d B=%d needs %zu B smem", Hs, B, smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
    return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: cannot set smem attribute");
  if (cudaMemsetAsync(barrier, 0, sizeof(uint32_t), st) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: memset failed");
  void* args[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_fwd_kernel), dim3(Hs / LU), dim3(LTHREADS), args, smem, st);
  MTASR_COUNT_LAUNCH();
  if (e != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_fwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return MTASR_OK;
}

extern "C" int mtasr_lstm_bwd(const float* dh_out, const float* gates, const float* c_all, const void* whh_bf16, int32_t ldw,
                              int32_t B, int32_t T, int32_t Hs, void* dgates_bf16, uint32_t* barrier, void* stream) {
  MTASR_CHECK_ARG(dh_out && gates && c_all && whh_bf16 && dgates_bf16 && barrier, "lstm_bwd: null pointer");
  MTASR_CHECK_ARG((reinterpret_cast<uintptr_t>(barrier) & 15) == 0, "lstm_bwd: scratch must be 16-byte aligned");
  cudaStream_t st0 = static_cast<cudaStream_t>(stream);
  {
    size_t smem_ll = 0;
    const int G = lstm_ll_pick(B, Hs, true, &smem_ll);
    if (G > 0 && T > 0 && ldw % 8 == 0) {
      const int S = (B + RB - 1) / RB;
      LstmLlP q{};
      q.whh = reinterpret_cast<const __nv_bfloat16*>(whh_bf16); q.c_all = const_cast<float*>(c_all); q.gates = const_cast<float*>(gates);
      q.dh_out = dh_out; q.dgates = reinterpret_cast<__nv_bfloat16*>(dgates_bf16);
      q.ll = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(barrier) + LL_OFFSET);
      q.B = B; q.T = T; q.Hs = Hs; q.ldw = ldw; q.G = G; q.U = Hs / G;
      q.spin_limit = lstm_spin_limit();
      void (*kern)(const LstmLlP) = (Hs == 896) ? lstm_bwd_ll_kernel<7> : lstm_bwd_ll_kernel<0>;
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_ll)) != cudaSuccess)
        return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: cannot set smem attribute");
      if (cudaMemsetAsync(q.ll, 0, lstm_ll_bytes(B, Hs, true), st0) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: memset failed");
      void* args[] = {&q};
      cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(S * G), dim3(LTHREADS), args, smem_ll, st0);
      MTASR_COUNT_LAUNCH();
      if (e != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
      return MTASR_OK;
    }
  }
  {
    size_t smem_bs = 0;
    const int G = getenv("MTASR_LSTM_WHOLE_BATCH") ? 0 : lstm_bs_pick(B, Hs, true, &smem_bs);
    if (G > 0 && T > 0 && ldw % 8 == 0) {
      const int S = (B + RB - 1) / RB;
      LstmBsP q{};
      q.whh = reinterpret_cast<const __nv_bfloat16*>(whh_bf16); q.c_all = const_cast<float*>(c_all); q.gates = const_cast<float*>(gates);
      q.dh_out = dh_out; q.dgates = reinterpret_cast<__nv_bfloat16*>(dgates_bf16); q.bar = barrier;
      q.B = B; q.T = T; q.Hs = Hs; q.ldw = ldw; q.G = G; q.U = Hs / G;
      q.spin_limit = lstm_spin_limit();
      void (*kern)(const LstmBsP) = (Hs == 896) ? lstm_bwd_bs_kernel<7> : lstm_bwd_bs_kernel<0>;   // Hs/128 k-steps per warp and chunk
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bs)) != cudaSuccess)
        return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: cannot set smem attribute");
      if (cudaMemsetAsync(barrier, 0, sizeof(uint32_t) * S, st0) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: memset failed");
      void* args[] = {&q};
      cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kern), dim3(S * G), dim3(LTHREADS), args, smem_bs, st0);
      MTASR_COUNT_LAUNCH();
      if (e != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
      return MTASR_OK;
    }
  }
  if (int rc = lstm_check(B, T, Hs, ldw, "lstm_bwd")) return rc;
  LstmBwdP p;
  p.spin_limit = lstm_spin_limit();
  p.dh_out = dh_out; p.gates = gates; p.c_all = c_all; p.whh = reinterpret_cast<const __nv_bfloat16*>(whh_bf16);
  p.dgates = reinterpret_cast<__nv_bfloat16*>(dgates_bf16); p.bar = barrier;
  p.B = B; p.T = T; p.Hs = Hs; p.ldw = ldw; p.Bp = (B + 15) / 16 * 16;
  const size_t smem = lstm_bwd_smem(Hs, p.Bp);
  if (smem > 227 * 1024) return set_error(MTASR_ERR_UNSUPPORTED, "lstm_bwd: Hs=%d B=%d needs %zu B smem", Hs, B, smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
    return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: cannot set smem attribute");
  if (cudaMemsetAsync(barrier, 0, sizeof(uint32_t), st) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: memset failed");
  void* args[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_bwd_kernel), dim3(Hs / LU), dim3(LTHREADS), args, smem, st);
  MTASR_COUNT_LAUNCH();
  if (e != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "lstm_bwd: cooperative launch failed: %s", cudaGetErrorString(e));
  return MTASR_OK;
}
