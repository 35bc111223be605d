// Serialized-CTC kernels: per-(utterance, speaker) alpha/beta recursion, greedy collapse, LSE finalisation.
//
// Layout ("compact lattice columns"): for one CTC head, glog[b][t][c] holds the LOGITS of only the columns the
// lattice of utterance b can touch: c = 0 is the blank, c = 1 + l is label y[b][l] (l < L_b).  lse[b][t] is the
// log-sum-exp of the full vocabulary row, so lp(t, c) = glog[b][t][c] - lse[b][t] is what torch's
// log_softmax + CTCLoss (ref:models/ctc.py:53-54) would read at (t, b, ext[s]).  The (B,T,V) tensor never exists.
//
// One warp per utterance; lane i owns NS consecutive extended states s = i*NS .. i*NS+NS-1 in registers, the
// s-1 / s-2 neighbours of its first two states come from lane i-1 by shuffle.  The recursion is kept NORMALISED
// (every step subtracts the warp-wide max, the running offset is accumulated in double), so fp32 state values stay
// O(1) and loss/gradients match the fp64 oracle to ~1e-6 instead of the ~1e-5 of an unnormalised fp32 lattice.
#include <cfloat>

#include "common.cuh"

namespace mtasr {

static constexpr float NEG_INF = -INFINITY;

// Label length of utterance b as the kernels may use it: never past the label row (ys_ld entries) nor past the Lp - 1 label
// columns of the compact lattice, whatever the caller put into ylens.
__device__ __forceinline__ int clamp_label_len(long long L, int Lp, int ys_ld) {
  const long long hi = ys_ld < Lp - 1 ? ys_ld : Lp - 1;
  return static_cast<int>(L < 0 ? 0 : (L > hi ? hi : L));
}

// log(exp a + exp b + exp c) on the SFU (ex2 / lg2 approximations, ~1e-7 relative): the recursion is one dependent chain
// per time step, so the instruction count of this function IS the latency of the kernel.  Branch-free on purpose: with
// a branch per call the NS calls of one step become NS serialised BSSY/BSYNC regions and their MUFU latencies add up
// instead of overlapping.  All-(-inf) input: the exponentials are 0, lg2(0) = -inf, result -inf.  The argument of the
// log otherwise lies in [1, 3], where lg2.approx is accurate to a few ulp; tests pin loss and gradients to the fp64
// oracle at 1e-5.
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m0 = fmaxf(a, fmaxf(b, c));
  const float m = m0 == NEG_INF ? 0.f : m0;
  const float k = 1.4426950408889634f;
  const float nm = -m * k;
  const float sum = ex2_approx(fmaf(a, k, nm)) + ex2_approx(fmaf(b, k, nm)) + ex2_approx(fmaf(c, k, nm));
  return fmaf(lg2_approx(sum), 0.6931471805599453f, m);
}

// Ampere-style asynchronous copies (global -> shared without a register round trip): the recursions below are one
// dependent chain per time step, so every global-memory latency inside the step is exposed.  All per-frame inputs are
// therefore streamed through a double-buffered shared-memory ring, CH frames per chunk, one chunk ahead of the recursion.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Warp-wide float max in one REDUX instead of five shuffle+max rounds: floats are mapped to integers of the same order.
__device__ __forceinline__ float warp_max_redux(float v) {
  int i = __float_as_int(v);
  i ^= (i >> 31) & 0x7fffffff;
  i = __reduce_max_sync(0xffffffffu, i);
  i ^= (i >> 31) & 0x7fffffff;
  return __int_as_float(i);
}

// ------------------------------------------------------------------------------------------------ alpha
// One warp (= one CTA) per utterance.  smem: [2][CH][Lp] lattice-column logits + [2][CH] row LSE.
template <int NS>
__global__ void __launch_bounds__(32)
ctc_alpha_kernel(const float* __restrict__ glog, const float* __restrict__ lse, const long long* __restrict__ ys,
                 const long long* __restrict__ hlens, const long long* __restrict__ ylens, int B, int T, int Lp,
                 int ys_ld, int CH, float* __restrict__ alpha_ws, double* __restrict__ coff_ws, float* __restrict__ nll_out,
                 double* __restrict__ nll_raw) {
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char ctc_smem[];
  constexpr int SP = 32 * NS;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  if (b >= B) return;
  int Tb = static_cast<int>(hlens[b]);
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  const int L = clamp_label_len(ylens[b], Lp, ys_ld);   // never walk past the label row / the lattice columns
  const int S = 2 * L + 1;
  const long long* y = ys + static_cast<long long>(b) * ys_ld;

  int col[NS];
  bool ok[NS], skip[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    ok[j] = s < S;
    col[j] = (s & 1) ? (s >> 1) + 1 : 0;
    skip[j] = ok[j] && (s & 1) && s >= 3 && (y[s >> 1] != y[(s >> 1) - 1]);
  }
  if (Tb == 0) {
    if (lane == 0) {
      const double v = L == 0 ? 0.0 : static_cast<double>(INFINITY);
      nll_raw[b] = v;
      nll_out[b] = isinf(v) ? 0.f : static_cast<float>(v);
    }
    return;
  }
  const float* g = glog + static_cast<long long>(b) * T * Lp;
  const float* ls = lse + static_cast<long long>(b) * T;
  float* aw = alpha_ws + static_cast<long long>(b) * T * SP;
  double* cw = coff_ws + static_cast<long long>(b) * T;
  float* sg = reinterpret_cast<float*>(ctc_smem);
  float* sl = sg + 2 * CH * Lp;

  auto issue = [&](int c) {
    const int t0 = c * CH;
    const int n = min(CH, Tb - t0);
    if (n > 0) {
      float* dst = sg + (c & 1) * CH * Lp;
      const float* src = g + static_cast<long long>(t0) * Lp;
      for (int i = lane; i < n * (Lp >> 2); i += 32) cp_async16(dst + i * 4, src + i * 4);
      for (int i = lane; i < n; i += 32) cp_async4(sl + (c & 1) * CH + i, ls + t0 + i);
    }
    cp_async_commit();
  };

  float a[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) a[j] = NEG_INF;
  double coff = 0.0;
  const int nchunk = (Tb + CH - 1) / CH;
  issue(0);
  for (int c = 0; c < nchunk; ++c) {
    issue(c + 1);
    cp_async_wait<1>();
    __syncwarp();
    const float* cg = sg + (c & 1) * CH * Lp;
    const float* cl = sl + (c & 1) * CH;
    const int t0 = c * CH;
    const int n = min(CH, Tb - t0);
    for (int f = 0; f < n; ++f) {
      const int t = t0 + f;
      float cur[NS];
      const float l = cl[f];
#pragma unroll
      for (int j = 0; j < NS; ++j) cur[j] = ok[j] ? cg[f * Lp + col[j]] - l : NEG_INF;
      if (t == 0) {
#pragma unroll
        for (int j = 0; j < NS; ++j) a[j] = (lane * NS + j <= 1) ? cur[j] : NEG_INF;
      } else {
        float pm1 = __shfl_up_sync(0xffffffffu, a[NS - 1], 1);
        float pm2 = __shfl_up_sync(0xffffffffu, a[NS - 2], 1);
        if (lane == 0) { pm1 = NEG_INF; pm2 = NEG_INF; }
        float na[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const float s1 = j >= 1 ? a[j - 1] : pm1;
          const float s2 = j >= 2 ? a[j - 2] : (j == 1 ? pm1 : pm2);
          na[j] = cur[j] + lse3(a[j], s1, skip[j] ? s2 : NEG_INF);   // cur = -inf for states past the lattice
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) a[j] = na[j];
      }
      // normalise
      float m = NEG_INF;
#pragma unroll
      for (int j = 0; j < NS; ++j) m = fmaxf(m, a[j]);
      m = warp_max_redux(m);
      if (m == NEG_INF) m = 0.f;  // dead lattice: stays -inf, nll becomes +inf below
#pragma unroll
      for (int j = 0; j < NS; ++j) a[j] -= m;
      coff += static_cast<double>(m);
      float4* dst = reinterpret_cast<float4*>(aw + static_cast<long long>(t) * SP + lane * NS);
      if constexpr (NS % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NS; j += 4) dst[j >> 2] = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) aw[static_cast<long long>(t) * SP + lane * NS + j] = a[j];
      }
      if (lane == 0) cw[t] = coff;
    }
    __syncwarp();   // all lanes are done with this buffer before chunk c + 2 is copied into it
  }
  float e1 = NEG_INF, e2 = NEG_INF;
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    if (s == S - 1) e1 = a[j];
    if (s == S - 2) e2 = a[j];
  }
  e1 = warp_max(e1);
  e2 = warp_max(e2);
  if (lane == 0) {
    const float tail = logaddexp_f(e1, e2);
    const double v = tail == NEG_INF ? static_cast<double>(INFINITY) : -(coff + static_cast<double>(tail));
    nll_raw[b] = v;  // kept in double: |nll| ~ 1e3 would cost 1e-4 absolute in fp32 and that error multiplies every occupancy
    nll_out[b] = isinf(v) ? 0.f : static_cast<float>(v);  // zero_infinity=True (ref:models/ctc.py:31,44-46)
  }
}

// ------------------------------------------------------------------------------------------------ beta + grad
// dG[b][t][c] = -gout[b] * occupancy(t, c) (dG arrives ZEROED: only reachable (t, c) are written);
// rowscale[b][t] = gout[b] for valid frames of feasible utterances.
// The dense part of d nll/d logits (softmax * rowscale) is regenerated by the vocab GEMM (mode 2).
// smem: [2][CH][Lp] logits + [2][CH][SP] saved alpha + [2][CH] coff (double) + [2][CH] row LSE.
template <int NS>
__global__ void __launch_bounds__(32)
ctc_beta_grad_kernel(const float* __restrict__ glog, const float* __restrict__ lse, const long long* __restrict__ ys,
                     const long long* __restrict__ hlens, const long long* __restrict__ ylens, int B, int T, int Lp,
                     int ys_ld, int CH, const float* __restrict__ alpha_ws, const double* __restrict__ coff_ws,
                     const double* __restrict__ nll_raw, const float* __restrict__ gout, float* __restrict__ dG,
                     float* __restrict__ rowscale) {
  pdl_trigger();
  extern __shared__ __align__(16) unsigned char ctc_smem[];
  constexpr int SP = 32 * NS;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  if (b >= B) return;
  int Tb = static_cast<int>(hlens[b]);
  Tb = Tb < 0 ? 0 : (Tb > T ? T : Tb);
  const int L = clamp_label_len(ylens[b], Lp, ys_ld);   // never walk past the label row / the lattice columns
  const int S = 2 * L + 1;
  const long long* y = ys + static_cast<long long>(b) * ys_ld;
  const double nll = nll_raw[b];
  const bool feasible = !isinf(nll) && Tb > 0;
  const float go = gout[b];
  float* rs = rowscale + static_cast<long long>(b) * T;
  for (int t = lane; t < T; t += 32) rs[t] = (feasible && t < Tb) ? go : 0.f;
  float* dg = dG + static_cast<long long>(b) * T * Lp;
  if (!feasible) return;

  int col[NS];
  bool ok[NS], skip[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) {
    const int s = lane * NS + j;
    ok[j] = s < S;
    col[j] = (s & 1) ? (s >> 1) + 1 : 0;
    skip[j] = ok[j] && (s & 1) && (s + 2 < S) && (y[s >> 1] != y[(s >> 1) + 1]);
  }
  const float* g = glog + static_cast<long long>(b) * T * Lp;
  const float* ls = lse + static_cast<long long>(b) * T;
  const float* aw = alpha_ws + static_cast<long long>(b) * T * SP;
  const double* cw = coff_ws + static_cast<long long>(b) * T;
  double* sc = reinterpret_cast<double*>(ctc_smem);            // [2][CH]
  float* sg = reinterpret_cast<float*>(sc + 2 * CH);           // [2][CH][Lp]
  float* sa = sg + 2 * CH * Lp;                                // [2][CH][SP]
  float* sl = sa + 2 * CH * SP;                                // [2][CH]

  const int nchunk = (Tb + CH - 1) / CH;
  auto issue = [&](int c) {      // chunk c covers frames [c*CH, c*CH + n); chunks are consumed from the last one down
    if (c >= 0) {
      const int t0 = c * CH;
      const int n = min(CH, Tb - t0);
      const int buf = c & 1;
      float* dst = sg + buf * CH * Lp;
      const float* src = g + static_cast<long long>(t0) * Lp;
      for (int i = lane; i < n * (Lp >> 2); i += 32) cp_async16(dst + i * 4, src + i * 4);
      float* dsta = sa + buf * CH * SP;
      const float* srca = aw + static_cast<long long>(t0) * SP;
      for (int i = lane; i < n * (SP >> 2); i += 32) cp_async16(dsta + i * 4, srca + i * 4);
      for (int i = lane; i < n; i += 32) {
        cp_async4(sl + buf * CH + i, ls + t0 + i);
        cp_async8(sc + buf * CH + i, cw + t0 + i);
      }
    }
    cp_async_commit();
  };

  float bt[NS];
#pragma unroll
  for (int j = 0; j < NS; ++j) bt[j] = NEG_INF;
  double boff = 0.0;
  issue(nchunk - 1);
  for (int c = nchunk - 1; c >= 0; --c) {
    issue(c - 1);
    cp_async_wait<1>();
    __syncwarp();
    const int buf = c & 1;
    const float* cg = sg + buf * CH * Lp;
    const float* ca = sa + buf * CH * SP;
    const float* cl = sl + buf * CH;
    const double* cc = sc + buf * CH;
    const int t0 = c * CH;
    const int n = min(CH, Tb - t0);
    for (int f = n - 1; f >= 0; --f) {
      const int t = t0 + f;
      float lp[NS];
      const float l = cl[f];
#pragma unroll
      for (int j = 0; j < NS; ++j) lp[j] = ok[j] ? cg[f * Lp + col[j]] - l : NEG_INF;
      if (t == Tb - 1) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const int s = lane * NS + j;
          bt[j] = (ok[j] && (s == S - 1 || s == S - 2)) ? lp[j] : NEG_INF;
        }
      } else {
        float np1 = __shfl_down_sync(0xffffffffu, bt[0], 1);
        float np2 = __shfl_down_sync(0xffffffffu, bt[1], 1);
        if (lane == 31) { np1 = NEG_INF; np2 = NEG_INF; }
        float nb[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          const float s1 = j + 1 < NS ? bt[j + 1 < NS ? j + 1 : 0] : np1;
          const float s2 = j + 2 < NS ? bt[j + 2 < NS ? j + 2 : 0] : (j + 2 == NS ? np1 : np2);
          nb[j] = lp[j] + lse3(bt[j], s1, skip[j] ? s2 : NEG_INF);   // lp = -inf for states past the lattice
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) bt[j] = nb[j];
      }
      float m = NEG_INF;
#pragma unroll
      for (int j = 0; j < NS; ++j) m = fmaxf(m, bt[j]);
      m = warp_max_redux(m);
      if (m == NEG_INF) m = 0.f;
#pragma unroll
      for (int j = 0; j < NS; ++j) bt[j] -= m;
      boff += static_cast<double>(m);
      // occupancy: exp(alpha + beta - lp + nll) with the three O(|nll|) offsets combined in double
      const float shift = static_cast<float>(cc[f] + boff + nll);
      const float* ar = ca + f * SP + lane * NS;
      float blank_occ = 0.f;
      float* dgr = dg + static_cast<long long>(t) * Lp;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        // unreachable states have alpha or beta = -inf -> e = -inf -> occupancy 0 (lp is finite wherever ok[j])
        const float e = ar[j] + bt[j] - (ok[j] ? lp[j] : 0.f) + shift;
        const float occ = ex2_approx(e * 1.4426950408889634f);
        if (ok[j]) {
          if (col[j] == 0) blank_occ += occ;
          else dgr[col[j]] = -go * occ;
        }
      }
      blank_occ = warp_sum(blank_occ);
      if (lane == 0) dgr[0] = -go * blank_occ;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ LSE finalise
// Combine the per-N-tile {max, sumexp, argmax} partials written by the vocab GEMM (mode 1).  One warp per row.
__global__ void lse_finalize_kernel(const float4* __restrict__ part, int rows, int n_tiles, float* __restrict__ lse,
                                    long long* __restrict__ argmax) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* p = part + row * n_tiles;
  float m = NEG_INF, s = 0.f;
  int idx = 0x7fffffff;
  for (int i = lane; i < n_tiles; i += 32) {
    const float4 v = p[i];
    const int vi = __float_as_int(v.z);
    if (v.x == NEG_INF) continue;   // empty partial (an epilogue warp that owned no column of this tile)
    if (v.x > m) {
      s = s * __expf(m - v.x) + v.y;
      m = v.x;
      idx = vi;
    } else {
      s += v.y * __expf(v.x - m);
      if (v.x == m && vi < idx) idx = vi;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const float os = __shfl_xor_sync(0xffffffffu, s, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    const float nm = fmaxf(m, om);
    const float sa = m == NEG_INF ? 0.f : s * __expf(m - nm);
    const float sb = om == NEG_INF ? 0.f : os * __expf(om - nm);
    if (om > m || (om == m && oi < idx)) idx = oi;
    m = nm;
    s = sa + sb;
  }
  if (lane == 0) {
    if (lse) lse[row] = m + __logf(s);
    if (argmax) argmax[row] = idx;
  }
}

// ------------------------------------------------------------------------------------------------ collapse
// Greedy collapse of ref:models/modeling_speech_encoder_decoder_llama.py:902-972: drop pad, drop blank, drop a token
// equal to the most recent non-blank/non-pad token.  One warp per row, 32 frames per step, ballot + popc ranking.
__global__ void ctc_collapse_kernel(const long long* __restrict__ ids, int B, int T, long long blank_id,
                                    long long pad_id, long long* __restrict__ out, int* __restrict__ lengths) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const long long* row = ids + static_cast<long long>(b) * T;
  long long* o = out + static_cast<long long>(b) * T;
  long long carry = 0;
  bool have_carry = false;
  int n = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const long long tok = t < T ? row[t] : pad_id;
    const bool valid = t < T && tok != pad_id && tok != blank_id;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    const unsigned below = vmask & ((1u << lane) - 1u);
    const int src = below ? 31 - __clz(below) : 0;
    const long long prev_in = __shfl_sync(0xffffffffu, tok, src);
    bool keep = false;
    if (valid) {
      if (below) keep = prev_in != tok;
      else keep = !have_carry || carry != tok;
    }
    const unsigned kmask = __ballot_sync(0xffffffffu, keep);
    if (keep) o[n + __popc(kmask & ((1u << lane) - 1u))] = tok;
    n += __popc(kmask);
    if (vmask) {
      carry = __shfl_sync(0xffffffffu, tok, 31 - __clz(vmask));
      have_carry = true;
    }
  }
  for (int t = n + lane; t < T; t += 32) o[t] = pad_id;
  if (lane == 0) lengths[b] = n;
}

// ------------------------------------------------------------------------------------------------ gathers
// dense (B,T,V) fp32 logits -> compact lattice columns (B,T,Lp): c=0 blank, c=1+l label l, rest 0.
__global__ void ctc_gather_cols_kernel(const float* __restrict__ dense, const long long* __restrict__ ys,
                                       const long long* __restrict__ ylens, int B, int T, int V, int Lp, int ys_ld,
                                       long long blank, float* __restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * T * Lp;
  if (i >= total) return;
  const int c = static_cast<int>(i % Lp);
  const long long bt = i / Lp;
  const int b = static_cast<int>(bt / T);
  const int L = clamp_label_len(ylens[b], Lp, ys_ld);
  float v = 0.f;
  long long id = -1;
  if (c == 0) id = blank;
  else if (c <= L) id = ys[static_cast<long long>(b) * ys_ld + c - 1];
  if (id >= 0 && id < V) v = dense[bt * V + id];        // ids outside [0, V) (e.g. a -100 pad inside ylens) read nothing
  out[i] = v;
}

// dense[b][t][col(c)] += src[b][t][c] for the lattice columns (inverse of the gather; repeated labels accumulate).
__global__ void ctc_scatter_cols_kernel(const float* __restrict__ src, const long long* __restrict__ ys,
                                        const long long* __restrict__ ylens, int B, int T, int V, int Lp, int ys_ld,
                                        long long blank, float* __restrict__ dense) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * T * Lp;
  if (i >= total) return;
  const int c = static_cast<int>(i % Lp);
  const long long bt = i / Lp;
  const int b = static_cast<int>(bt / T);
  const int L = clamp_label_len(ylens[b], Lp, ys_ld);
  if (c > L) return;
  const long long v = c == 0 ? blank : ys[static_cast<long long>(b) * ys_ld + c - 1];
  if (v < 0 || v >= V) return;                          // out-of-range id: nothing to scatter to
  atomicAdd(dense + bt * V + v, src[i]);
}

// Rows of the (V, D) head weight needed by each utterance's lattice -> Wg (B, Lp, D) bf16 (+ bias -> bg (B, Lp)).
__global__ void ctc_gather_rows_kernel(const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                                       const long long* __restrict__ ys, const long long* __restrict__ ylens, int B,
                                       int Lp, int D, int ys_ld, long long blank, long long V,
                                       __nv_bfloat16* __restrict__ wg, float* __restrict__ bg) {
  pdl_trigger();
  const int b = blockIdx.x / Lp, c = blockIdx.x % Lp;
  const int L = clamp_label_len(ylens[b], Lp, ys_ld);
  long long v = -1;
  if (c == 0) v = blank;
  else if (c <= L) v = ys[static_cast<long long>(b) * ys_ld + c - 1];
  if (v >= V) v = -1;                                   // ids outside [0, V) gather a zero row (never out of bounds)
  __nv_bfloat16* dst = wg + (static_cast<long long>(b) * Lp + c) * D;
  for (int i = threadIdx.x * 8; i < D; i += blockDim.x * 8) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (v >= 0) u = *reinterpret_cast<const uint4*>(w + v * D + i);
    *reinterpret_cast<uint4*>(dst + i) = u;
  }
  if (threadIdx.x == 0 && bg) bg[static_cast<long long>(b) * Lp + c] = (v >= 0 && bias) ? bias[v] : 0.f;
}

// dW[v(b,c)][:] += dWg[b][c][:]; db[v] += dbg[b][c]  (fp32 atomics; <= B*(L+1) rows touched).
__global__ void ctc_scatter_rows_kernel(const float* __restrict__ dwg, const float* __restrict__ dbg,
                                        const long long* __restrict__ ys, const long long* __restrict__ ylens, int B,
                                        int Lp, int D, int ys_ld, long long blank, long long V,
                                        float* __restrict__ dw, float* __restrict__ db) {
  pdl_trigger();
  const int b = blockIdx.x / Lp, c = blockIdx.x % Lp;
  const int L = clamp_label_len(ylens[b], Lp, ys_ld);
  if (c > L) return;
  const long long v = c == 0 ? blank : ys[static_cast<long long>(b) * ys_ld + c - 1];
  if (v < 0 || v >= V) return;                          // out-of-range id (e.g. -100 inside ylens): no write
  const float* src = dwg + (static_cast<long long>(b) * Lp + c) * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(dw + v * D + i, src[i]);
  if (threadIdx.x == 0 && db && dbg) atomicAdd(db + v, dbg[static_cast<long long>(b) * Lp + c]);
}

// ------------------------------------------------------------------------------------------------ label splitter
// ref:utils/split_labels_by_sc.py:21-75 on the device: per row, cut at the first `end_id`, split at `sep_id`
// into exactly K segments, drop `ignore_id` everywhere, right-trim `pad_id`; out (K, B, L) arrives filled with the pad value
// and receives the kept tokens, lens (K, B) the segment lengths.  status[0] = smallest failing row (INT_MAX = none),
// status[1] = failure kind of that row's last writer (1 wrong separator count, 2 empty segment), status[2] = the separator
// count / slot it saw.  One thread per row: the scan is sequential and a few hundred tokens long.
__global__ void split_labels_kernel(const long long* __restrict__ labels, int B, int L, long long ld, int K, long long sep_id,
                                    long long pad_id, int has_pad, long long ignore_id, int has_ignore, long long end_id, int has_end,
                                    int allow_empty, long long* __restrict__ out, long long* __restrict__ lens, int* __restrict__ status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long* row = labels + static_cast<long long>(b) * ld;
  int seg = 0, n = 0, last_keep = 0, nsep = 0;
  auto close = [&]() {
    if (seg < K) lens[static_cast<long long>(seg) * B + b] = has_pad ? last_keep : n;
  };
  for (int j = 0; j < L; ++j) {
    const long long tok = row[j];
    if (has_end && tok == end_id) break;
    if (tok == sep_id) {
      close();
      ++seg; ++nsep; n = 0; last_keep = 0;
      continue;
    }
    if (has_ignore && tok == ignore_id) continue;
    if (seg < K) out[(static_cast<long long>(seg) * B + b) * L + n] = tok;
    ++n;
    if (tok != pad_id) last_keep = n;
  }
  close();
  for (int s2 = seg + 1; s2 < K; ++s2) lens[static_cast<long long>(s2) * B + b] = 0;
  int kind = 0, info = 0;
  if (nsep != K - 1) { kind = 1; info = nsep; }
  else if (!allow_empty) {
    for (int s2 = 0; s2 < K; ++s2)
      if (lens[static_cast<long long>(s2) * B + b] == 0) { kind = 2; info = s2; break; }
  }
  if (kind) {
    const int prev = atomicMin(status, b);
    if (b <= prev) { status[1] = kind; status[2] = info; }   // the smallest failing row reports (ties cannot happen)
  }
}

// ------------------------------------------------------------------------------------------------ token segments
// Segmentation of ref:models/mt_ctctoken_builder.py:56-157 on the greedy path: a segment is a maximal run of one
// non-blank token and is EMITTED when a blank follows it or the valid region ends; a token change without a blank
// restarts the run and silently drops the previous one (reference behaviour, kept); scanning stops at the first
// masked frame.  One thread per utterance (the scan is sequential and tiny).
__global__ void ctc_segments_kernel(const long long* __restrict__ path, const unsigned char* __restrict__ mask, int B, int T,
                                    long long blank, int* __restrict__ seg_start, int* __restrict__ seg_end,
                                    int* __restrict__ nseg) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const long long* pr = path + static_cast<long long>(b) * T;
  const unsigned char* mr = mask + static_cast<long long>(b) * T;
  int* ss = seg_start + static_cast<long long>(b) * T;
  int* se = seg_end + static_cast<long long>(b) * T;
  long long prev = -1;
  int cs = -1, ce = -1, n = 0;
  for (int t = 0; t < T; ++t) {
    if (!mr[t]) break;
    const long long tok = pr[t];
    if (tok == blank) {
      if (cs >= 0) { ss[n] = cs; se[n] = ce; ++n; cs = -1; }
      prev = -1;
      continue;
    }
    if (prev < 0 || tok != prev) { cs = t; ce = t; prev = tok; }
    else ce = t;
  }
  if (cs >= 0) { ss[n] = cs; se[n] = ce; ++n; }
  nseg[b] = n;
}

// out[b][j][:] = mean over frames seg_start[b][j]..seg_end[b][j] of x[b][t][:]  (j < nseg[b], else 0); conf analog on a
// per-frame scalar.  One CTA per (b, j).
__global__ void segment_mean_fwd_kernel(const float* __restrict__ x, const float* __restrict__ pblank, const int* __restrict__ seg_start,
                                        const int* __restrict__ seg_end, const int* __restrict__ nseg, int T, int D, int Lmax,
                                        float* __restrict__ out, float* __restrict__ conf) {
  const int b = blockIdx.y, j = blockIdx.x;
  float* o = out + (static_cast<long long>(b) * Lmax + j) * D;
  if (j >= nseg[b]) {
    for (int c = threadIdx.x; c < D; c += blockDim.x) o[c] = 0.f;
    if (threadIdx.x == 0 && conf) conf[static_cast<long long>(b) * Lmax + j] = 0.f;
    return;
  }
  const int s = seg_start[static_cast<long long>(b) * T + j], e = seg_end[static_cast<long long>(b) * T + j];
  const float inv = 1.f / static_cast<float>(e - s + 1);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f;
    for (int t = s; t <= e; ++t) a += x[(static_cast<long long>(b) * T + t) * D + c];
    o[c] = a * inv;
  }
  if (threadIdx.x == 0 && conf) {
    float a = 0.f;
    for (int t = s; t <= e; ++t) a += pblank[static_cast<long long>(b) * T + t];
    conf[static_cast<long long>(b) * Lmax + j] = fminf(fmaxf(1.f - a * inv, 0.f), 1.f);
  }
}

// dx[b][t][:] += dout[b][j][:] / len(j) for the frames of every emitted segment (dx zero-initialised; segments are disjoint)
__global__ void segment_mean_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ seg_start,
                                        const int* __restrict__ seg_end, const int* __restrict__ nseg, int T, int D, int Lmax,
                                        float* __restrict__ dx) {
  const int b = blockIdx.y, j = blockIdx.x;
  if (j >= nseg[b]) return;
  const int s = seg_start[static_cast<long long>(b) * T + j], e = seg_end[static_cast<long long>(b) * T + j];
  const float inv = 1.f / static_cast<float>(e - s + 1);
  const float* g = dout + (static_cast<long long>(b) * Lmax + j) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = g[c] * inv;
    for (int t = s; t <= e; ++t) dx[(static_cast<long long>(b) * T + t) * D + c] = v;
  }
}

static int pick_ns(int max_states) {
  if (max_states <= 64) return 2;
  if (max_states <= 128) return 4;
  if (max_states <= 256) return 8;
  if (max_states <= 512) return 16;
  return -1;
}

}  // namespace mtasr

using namespace mtasr;

// frames per shared-memory chunk of the alpha / beta kernels: two buffers within the 48 KB static-opt-in-free limit,
// a multiple of 2, at most 32
static int ctc_chunk_frames(int frame_bytes) {
  int ch = 22000 / frame_bytes;
  ch = ch > 32 ? 32 : ch;
  ch &= ~1;
  return ch < 2 ? 2 : ch;
}

extern "C" int mtasr_ctc_state_pad(int32_t max_label_len) {
  const int ns = pick_ns(2 * max_label_len + 1);
  return ns < 0 ? -1 : 32 * ns;
}

extern "C" int mtasr_ctc_alpha_fwd(const float* glog, const float* lse, const int64_t* ys, const int64_t* hlens,
                                   const int64_t* ylens, int32_t B, int32_t T, int32_t Lp, int32_t ys_ld,
                                   int32_t max_label_len, float* alpha_ws, double* coff_ws, float* nll_out,
                                   double* nll_raw, void* stream) {
  MTASR_CHECK_ARG(glog && lse && hlens && ylens && alpha_ws && coff_ws && nll_out && nll_raw, "ctc_alpha_fwd: null pointer");
  MTASR_CHECK_ARG(ys || max_label_len == 0, "ctc_alpha_fwd: null labels");
  MTASR_CHECK_ARG(B > 0 && T > 0 && Lp >= max_label_len + 1, "ctc_alpha_fwd: bad sizes B=%d T=%d Lp=%d Lmax=%d", B, T, Lp, max_label_len);
  const int ns = pick_ns(2 * max_label_len + 1);
  if (ns < 0) return set_error(MTASR_ERR_UNSUPPORTED, "ctc_alpha_fwd: label length %d > 255 not supported", max_label_len);
  MTASR_CHECK_ARG(Lp % 4 == 0 && (reinterpret_cast<uintptr_t>(glog) & 15) == 0, "ctc_alpha_fwd: Lp %% 4 and 16-byte aligned glog required");
  const int frame_bytes = Lp * 4 + 4;
  const int CH = ctc_chunk_frames(frame_bytes);
  const size_t smem = 2 * static_cast<size_t>(CH) * frame_bytes;
  dim3 grid(B), block(32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* y = reinterpret_cast<const long long*>(ys);
  const long long* hl = reinterpret_cast<const long long*>(hlens);
  const long long* yl = reinterpret_cast<const long long*>(ylens);
  switch (ns) {
    case 2: ctc_alpha_kernel<2><<<grid, block, smem, st>>>(glog, lse, y, hl, yl, B, T, Lp, ys_ld, CH, alpha_ws, coff_ws, nll_out, nll_raw); break;
    case 4: ctc_alpha_kernel<4><<<grid, block, smem, st>>>(glog, lse, y, hl, yl, B, T, Lp, ys_ld, CH, alpha_ws, coff_ws, nll_out, nll_raw); break;
    case 8: ctc_alpha_kernel<8><<<grid, block, smem, st>>>(glog, lse, y, hl, yl, B, T, Lp, ys_ld, CH, alpha_ws, coff_ws, nll_out, nll_raw); break;
    default: ctc_alpha_kernel<16><<<grid, block, smem, st>>>(glog, lse, y, hl, yl, B, T, Lp, ys_ld, CH, alpha_ws, coff_ws, nll_out, nll_raw); break;
  }
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("ctc_alpha_fwd");
  return MTASR_OK;
}

extern "C" int mtasr_ctc_beta_bwd(const float* glog, const float* lse, const int64_t* ys, const int64_t* hlens,
                                  const int64_t* ylens, int32_t B, int32_t T, int32_t Lp, int32_t ys_ld,
                                  int32_t max_label_len, const float* alpha_ws, const double* coff_ws,
                                  const double* nll_raw, const float* gout, float* dG, float* rowscale, void* stream) {
  MTASR_CHECK_ARG(glog && lse && hlens && ylens && alpha_ws && coff_ws && nll_raw && gout && dG && rowscale, "ctc_beta_bwd: null pointer");
  MTASR_CHECK_ARG(B > 0 && T > 0 && Lp >= max_label_len + 1, "ctc_beta_bwd: bad sizes");
  const int ns = pick_ns(2 * max_label_len + 1);
  if (ns < 0) return set_error(MTASR_ERR_UNSUPPORTED, "ctc_beta_bwd: label length %d > 255 not supported", max_label_len);
  MTASR_CHECK_ARG(Lp % 4 == 0 && (reinterpret_cast<uintptr_t>(glog) & 15) == 0 && (reinterpret_cast<uintptr_t>(alpha_ws) & 15) == 0,
                  "ctc_beta_bwd: Lp %% 4 and 16-byte aligned glog / alpha_ws required");
  const int frame_bytes = Lp * 4 + 32 * ns * 4 + 4 + 8;
  const int CH = ctc_chunk_frames(frame_bytes);
  const size_t smem = 2 * static_cast<size_t>(CH) * frame_bytes;
  dim3 grid(B), block(32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* y = reinterpret_cast<const long long*>(ys);
  const long long* hl = reinterpret_cast<const long long*>(hlens);
  const long long* yl = reinterpret_cast<const long long*>(ylens);
  switch (ns) {
    case 2: ctc_beta_grad_kernel<2><<<grid, block, smem, st>>>(glog, lse, y, hl, yl, B, T, Lp, ys_ld, CH, alpha_ws, coff_ws, nll_raw, gout, dG, rowscale); break;
    case 4: ctc_beta_grad_kernel<4><<<grid, block, smem, st>>>(glog, lse, y, hl, yl, B, T, Lp, ys_ld, CH, alpha_ws, coff_ws, nll_raw, gout, dG, rowscale); break;
    case 8: ctc_beta_grad_kernel<8><<<grid, block, smem, st>>>(glog, lse, y, hl, yl, B, T, Lp, ys_ld, CH, alpha_ws, coff_ws, nll_raw, gout, dG, rowscale); break;
    default: ctc_beta_grad_kernel<16><<<grid, block, smem, st>>>(glog, lse, y, hl, yl, B, T, Lp, ys_ld, CH, alpha_ws, coff_ws, nll_raw, gout, dG, rowscale); break;
  }
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("ctc_beta_bwd");
  return MTASR_OK;
}

extern "C" int mtasr_lse_finalize(const float* part, int64_t rows, int32_t n_tiles, float* lse, int64_t* argmax,
                                  void* stream) {
  MTASR_CHECK_ARG(part && rows > 0 && n_tiles > 0 && (lse || argmax), "lse_finalize: bad arguments");
  const int wpb = 8;
  lse_finalize_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), 32 * wpb, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(part), static_cast<int>(rows), n_tiles, lse, reinterpret_cast<long long*>(argmax));
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("lse_finalize");
  return MTASR_OK;
}

extern "C" int mtasr_ctc_collapse(const int64_t* ids, int32_t B, int32_t T, int64_t blank_id, int64_t pad_id,
                                  int64_t* out, int32_t* lengths, void* stream) {
  MTASR_CHECK_ARG(ids && out && lengths && B > 0 && T > 0, "ctc_collapse: bad arguments");
  const int wpb = 4;
  ctc_collapse_kernel<<<(B + wpb - 1) / wpb, 32 * wpb, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), B, T, blank_id, pad_id, reinterpret_cast<long long*>(out), lengths);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("ctc_collapse");
  return MTASR_OK;
}

extern "C" int mtasr_ctc_gather_cols(const float* dense, const int64_t* ys, const int64_t* ylens, int32_t B, int32_t T,
                                     int32_t V, int32_t Lp, int32_t ys_ld, int64_t blank, float* out, void* stream) {
  MTASR_CHECK_ARG(dense && ylens && out && B > 0 && T > 0 && V > 0 && Lp > 0, "ctc_gather_cols: bad arguments");
  const long long total = static_cast<long long>(B) * T * Lp;
  ctc_gather_cols_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dense, reinterpret_cast<const long long*>(ys), reinterpret_cast<const long long*>(ylens), B, T, V, Lp, ys_ld, blank, out);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("ctc_gather_cols");
  return MTASR_OK;
}

extern "C" int mtasr_ctc_scatter_cols(const float* src, const int64_t* ys, const int64_t* ylens, int32_t B, int32_t T,
                                      int32_t V, int32_t Lp, int32_t ys_ld, int64_t blank, float* dense, void* stream) {
  MTASR_CHECK_ARG(src && ylens && dense && B > 0 && T > 0 && V > 0 && Lp > 0, "ctc_scatter_cols: bad arguments");
  const long long total = static_cast<long long>(B) * T * Lp;
  ctc_scatter_cols_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<const long long*>(ys), reinterpret_cast<const long long*>(ylens), B, T, V, Lp, ys_ld, blank, dense);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("ctc_scatter_cols");
  return MTASR_OK;
}

extern "C" int mtasr_ctc_gather_rows(const void* w_bf16, const float* bias, const int64_t* ys, const int64_t* ylens,
                                     int32_t B, int32_t Lp, int32_t D, int32_t ys_ld, int64_t blank, int64_t V, void* wg_bf16,
                                     float* bg, void* stream) {
  MTASR_CHECK_ARG(w_bf16 && ylens && wg_bf16 && B > 0 && Lp > 0 && D > 0 && D % 8 == 0 && V > 0 && blank >= 0 && blank < V,
                  "ctc_gather_rows: bad arguments");
  ctc_gather_rows_kernel<<<B * Lp, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(w_bf16), bias, reinterpret_cast<const long long*>(ys),
      reinterpret_cast<const long long*>(ylens), B, Lp, D, ys_ld, blank, V, reinterpret_cast<__nv_bfloat16*>(wg_bf16), bg);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("ctc_gather_rows");
  return MTASR_OK;
}

extern "C" int mtasr_ctc_scatter_rows(const float* dwg, const float* dbg, const int64_t* ys, const int64_t* ylens,
                                      int32_t B, int32_t Lp, int32_t D, int32_t ys_ld, int64_t blank, int64_t V, float* dw,
                                      float* db, void* stream) {
  MTASR_CHECK_ARG(dwg && ylens && dw && B > 0 && Lp > 0 && D > 0 && V > 0 && blank >= 0 && blank < V, "ctc_scatter_rows: bad arguments");
  ctc_scatter_rows_kernel<<<B * Lp, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dwg, dbg, reinterpret_cast<const long long*>(ys), reinterpret_cast<const long long*>(ylens), B, Lp, D, ys_ld, blank, V, dw, db);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("ctc_scatter_rows");
  return MTASR_OK;
}

extern "C" int mtasr_split_labels(const int64_t* labels, int32_t B, int32_t L, int64_t ld, int32_t K, int64_t sep_id, int64_t pad_id,
                                  int32_t has_pad, int64_t ignore_id, int32_t has_ignore, int64_t end_id, int32_t has_end,
                                  int32_t allow_empty, int64_t* out, int64_t* lens, int32_t* status, void* stream) {
  MTASR_CHECK_ARG(labels && out && lens && status && B > 0 && L >= 0 && K > 0 && ld >= L, "split_labels: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  split_labels_kernel<<<(B + 63) / 64, 64, 0, st>>>(reinterpret_cast<const long long*>(labels), B, L, ld, K, sep_id, pad_id, has_pad,
                                                    ignore_id, has_ignore, end_id, has_end, allow_empty,
                                                    reinterpret_cast<long long*>(out), reinterpret_cast<long long*>(lens), status);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("split_labels");
  return MTASR_OK;
}

extern "C" int mtasr_ctc_segments(const int64_t* path, const uint8_t* mask, int32_t B, int32_t T, int64_t blank, int32_t* seg_start,
                                  int32_t* seg_end, int32_t* nseg, void* stream) {
  MTASR_CHECK_ARG(path && mask && seg_start && seg_end && nseg && B > 0 && T > 0, "ctc_segments: bad arguments");
  ctc_segments_kernel<<<(B + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const long long*>(path), mask, B, T,
                                                                                   blank, seg_start, seg_end, nseg);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("ctc_segments");
  return MTASR_OK;
}

extern "C" int mtasr_segment_mean_fwd(const float* x, const float* pblank, const int32_t* seg_start, const int32_t* seg_end,
                                      const int32_t* nseg, int32_t B, int32_t T, int32_t D, int32_t Lmax, float* out, float* conf,
                                      void* stream) {
  MTASR_CHECK_ARG(x && seg_start && seg_end && nseg && out && B > 0 && T > 0 && D > 0 && Lmax > 0, "segment_mean_fwd: bad arguments");
  MTASR_CHECK_ARG(conf == nullptr || pblank != nullptr, "segment_mean_fwd: conf needs pblank");
  segment_mean_fwd_kernel<<<dim3(Lmax, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, pblank, seg_start, seg_end, nseg, T, D, Lmax,
                                                                                        out, conf);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("segment_mean_fwd");
  return MTASR_OK;
}

extern "C" int mtasr_segment_mean_bwd(const float* dout, const int32_t* seg_start, const int32_t* seg_end, const int32_t* nseg,
                                      int32_t B, int32_t T, int32_t D, int32_t Lmax, float* dx, void* stream) {
  MTASR_CHECK_ARG(dout && seg_start && seg_end && nseg && dx && B > 0 && T > 0 && D > 0 && Lmax > 0, "segment_mean_bwd: bad arguments");
  segment_mean_bwd_kernel<<<dim3(Lmax, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(dout, seg_start, seg_end, nseg, T, D, Lmax, dx);
  g_launches.fetch_add(1);
  MTASR_CHECK_LAUNCH("segment_mean_bwd");
  return MTASR_OK;
}
