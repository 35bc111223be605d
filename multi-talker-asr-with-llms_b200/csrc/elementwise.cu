// HBM-bound row kernels of the hot path: LayerNorm fwd/bwd (+GELU), casts, column sums, attention softmax with
// the gated relative-position bias folded in (fwd/bwd), padding copies, GLU.  All use 128-bit accesses, one warp
// per row, grid-stride over rows with grids sized to a multiple of the SM count.
#include <cuda_fp16.h>

#include "common.cuh"

namespace mtasr {

static constexpr int MAXV = 32;  // per-lane register budget for one row: D <= 32*32 = 1024

__device__ __forceinline__ void ld8(const void* base, int dtype, long long idx, float (&o)[8]) {
  if (dtype == MTASR_DT_BF16) {
    const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    float2 t;
    t = unpack_bf16x2(u.x); o[0] = t.x; o[1] = t.y;
    t = unpack_bf16x2(u.y); o[2] = t.x; o[3] = t.y;
    t = unpack_bf16x2(u.z); o[4] = t.x; o[5] = t.y;
    t = unpack_bf16x2(u.w); o[6] = t.x; o[7] = t.y;
  } else {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void st8_f32(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ------------------------------------------------------------------------------------------------ LayerNorm fwd
// y = LN(x) * gamma + beta, optional GELU after (conv feature layers, hf:726).  Outputs bf16 and/or fp32.
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, long long rows, int D, int post_gelu,
                     __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const int nchunk = D >> 3;
  for (long long row = warp0; row < rows; row += nwarps) {
    float r[MAXV];
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < MAXV / 8; ++n) {
      const int c = lane + 32 * n;
      float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (c < nchunk) ld8(x, x_dtype, row * D + c * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) { r[n * 8 + j] = t[j]; s += t[j]; }
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int n = 0; n < MAXV / 8; ++n) {
      if (lane + 32 * n < nchunk) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = r[n * 8 + j] - mean; q += d * d; }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / D + eps);
#pragma unroll
    for (int n = 0; n < MAXV / 8; ++n) {
      const int c = lane + 32 * n;
      if (c < nchunk) {
        float o[8], gm[8], bt[8];
        ld8(gamma, MTASR_DT_F32, c * 8, gm);
        ld8(beta, MTASR_DT_F32, c * 8, bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = (r[n * 8 + j] - mean) * rstd * gm[j] + bt[j];
          o[j] = post_gelu ? gelu_fast_f(t) : t;
        }
        if (y_bf16) st8_bf16(y_bf16 + row * D + c * 8, o);
        if (y_f32) st8_f32(y_f32 + row * D + c * 8, o);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm bwd
// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; optional + dres (residual-stream gradient).
// One warp per row; the parameter gradients are a separate column-reduction kernel (below) so that this one stays light
// on registers (two resident CTAs per SM) -- keeping 2 x D/32 running sums per lane here cost more in occupancy than
// re-reading dy and x costs in bandwidth.
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ x, int x_dtype,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                     const float* __restrict__ dres, long long rows, int D, float* __restrict__ dx_f32,
                     __nv_bfloat16* __restrict__ dx_bf16) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const int nchunk = D >> 3;
  for (long long row = warp0; row < rows; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float g[MAXV], xh[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int n = 0; n < MAXV / 8; ++n) {
      const int c = lane + 32 * n;
      if (c < nchunk) {
        float a[8], b[8], gm[8];
        ld8(dy, dy_dtype, row * D + c * 8, a);
        ld8(x, x_dtype, row * D + c * 8, b);
        ld8(gamma, MTASR_DT_F32, c * 8, gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xhat = (b[j] - mu) * rs;
          const float gg = a[j] * gm[j];
          xh[n * 8 + j] = xhat;
          g[n * 8 + j] = gg;
          s1 += gg;
          s2 += gg * xhat;
        }
      }
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int n = 0; n < MAXV / 8; ++n) {
      const int c = lane + 32 * n;
      if (c < nchunk) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rs * (g[n * 8 + j] - s1 - xh[n * 8 + j] * s2);
        if (dres) {
          float d[8];
          ld8(dres, MTASR_DT_F32, row * D + c * 8, d);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += d[j];
        }
        if (dx_f32) st8_f32(dx_f32 + row * D + c * 8, o);
        if (dx_bf16) st8_bf16(dx_bf16 + row * D + c * 8, o);
      }
    }
  }
}

// dgamma[c] += sum_rows dy * xhat, dbeta[c] += sum_rows dy (zero them first).  Column reduction: thread = 8 consecutive
// columns, 8 row-lanes per CTA, 2 rows in flight per thread; CTA partials combined in smem, one atomic per column.
__global__ void __launch_bounds__(256)
layernorm_bwd_param_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ x, int x_dtype,
                           const float* __restrict__ mean, const float* __restrict__ rstd, long long rows, int D,
                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_trigger();
  __shared__ float redg[8][32][9], redb[8][32][9];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * 8;
  float sg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, sb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < D) {
    const long long step = static_cast<long long>(gridDim.y) * 8;
    long long m = static_cast<long long>(blockIdx.y) * 8 + ry;
    for (; m + step < rows; m += 2 * step) {
      float a0[8], b0[8], a1[8], b1[8];
      ld8(dy, dy_dtype, m * D + col, a0);
      ld8(x, x_dtype, m * D + col, b0);
      ld8(dy, dy_dtype, (m + step) * D + col, a1);
      ld8(x, x_dtype, (m + step) * D + col, b1);
      const float mu0 = mean[m], rs0 = rstd[m], mu1 = mean[m + step], rs1 = rstd[m + step];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sg[j] += a0[j] * ((b0[j] - mu0) * rs0) + a1[j] * ((b1[j] - mu1) * rs1);
        sb[j] += a0[j] + a1[j];
      }
    }
    for (; m < rows; m += step) {
      float a0[8], b0[8];
      ld8(dy, dy_dtype, m * D + col, a0);
      ld8(x, x_dtype, m * D + col, b0);
      const float mu0 = mean[m], rs0 = rstd[m];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sg[j] += a0[j] * ((b0[j] - mu0) * rs0);
        sb[j] += a0[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { redg[ry][cx][j] = sg[j]; redb[ry][cx][j] = sb[j]; }
  __syncthreads();
  const int c = threadIdx.x;
  const int gcol = blockIdx.x * 256 + c;
  if (gcol < D) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { tg += redg[i][c >> 3][c & 7]; tb += redb[i][c >> 3][c & 7]; }
    if (dgamma) atomicAdd(dgamma + gcol, tg);
    if (dbeta) atomicAdd(dbeta + gcol, tb);
  }
}

// Fused LayerNorm backward: dx (+ dres) AND the three column reductions that used to be separate passes over the same
// rows -- dgamma = sum_rows dy * xhat, dbeta = sum_rows dy, and dxsum = sum_rows dx (the bias gradient of the Linear
// whose output gradient this dx is: out-proj / FFN2 of the neighbouring half layer).  A row is owned by a GROUP of NW
// warps, thread = 8 consecutive columns, so the running column sums are 24 registers per thread (the warp-per-row kernel
// above would need 3 x D/32); the two row scalars cross the group's warps through a parity-double-buffered smem slot and
// ONE named barrier per row.  The next row's loads are issued before the current row's reductions.
__device__ __forceinline__ void ld_raw8(const void* base, int dtype, long long idx, uint4& u0, uint4& u1) {
  if (dtype == MTASR_DT_BF16) {
    u0 = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
  } else {
    u0 = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(base) + idx);
    u1 = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(base) + idx + 4);
  }
}
__device__ __forceinline__ void cvt_raw8(int dtype, const uint4& u0, const uint4& u1, float (&o)[8]) {
  if (dtype == MTASR_DT_BF16) {
    float2 t;
    t = unpack_bf16x2(u0.x); o[0] = t.x; o[1] = t.y;
    t = unpack_bf16x2(u0.y); o[2] = t.x; o[3] = t.y;
    t = unpack_bf16x2(u0.z); o[4] = t.x; o[5] = t.y;
    t = unpack_bf16x2(u0.w); o[6] = t.x; o[7] = t.y;
  } else {
    o[0] = __uint_as_float(u0.x); o[1] = __uint_as_float(u0.y); o[2] = __uint_as_float(u0.z); o[3] = __uint_as_float(u0.w);
    o[4] = __uint_as_float(u1.x); o[5] = __uint_as_float(u1.y); o[6] = __uint_as_float(u1.z); o[7] = __uint_as_float(u1.w);
  }
}

// entry i of the summed projection: i < 64 -> rows 0..3 of the (8, 64) weight (gate a), else rows 4..7 (gate b), hf:170-173
__device__ __forceinline__ float gate_wsum(const float* __restrict__ w8, int i) {
  const float* p = w8 + (i >> 6) * 256 + (i & 63);
  return (p[0] + p[64]) + (p[128] + p[192]);
}

template <int NW>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_fused_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ x, int x_dtype,
                           const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                           const float* __restrict__ dres, long long rows, int D, float* __restrict__ dx_f32,
                           __nv_bfloat16* __restrict__ dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                           float* __restrict__ dxsum, const float* __restrict__ gate_ab, const float* __restrict__ gate_w8) {
  pdl_trigger();
  // gate_ab (rows, D/64, 2) / gate_w8 (8, 64): the gru_rel_pos gate's gradient path into the LayerNorm OUTPUT (hf:167-176) is
  // rank 2 per head -- dy[c] += da[row, head] * wa[c % 64] + db[row, head] * wb[c % 64] -- and is added here from the two
  // scalars per (row, head) instead of being materialised as a (rows, D) fp32 tensor and read back by a GEMM epilogue.
  constexpr int G = NW * 32;                       // threads per row group
  constexpr int NG = NW == 3 ? 2 : 256 / G;        // row groups per CTA (blockDim.x = NG * G)
  constexpr int GC = G * 8;                        // columns a group spans
  __shared__ float2 red[2][NG][NW];
  __shared__ float acc[NG][GC];
  const int g = threadIdx.x / G, tg = threadIdx.x % G, wg = tg >> 5, lane = tg & 31;
  const int c0 = tg * 8;
  const bool active = c0 < D;
  const bool has_res = dres != nullptr;
  float gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gm[j] = 0.f;
  if (active) ld8(gamma, MTASR_DT_F32, c0, gm);
  float ag[8], ab[8], ax[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ag[j] = ab[j] = ax[j] = 0.f;
  const bool has_gate = gate_ab != nullptr;
  const int gH = D >> 6, ghead = c0 >> 6;
  float gwa[8], gwb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gwa[j] = gwb[j] = 0.f;
  if (has_gate && active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gwa[j] = gate_wsum(gate_w8, (c0 & 63) + j);
      gwb[j] = gate_wsum(gate_w8, 64 + (c0 & 63) + j);
    }
  }
  float2 gab = make_float2(0.f, 0.f);
  const long long stride = static_cast<long long>(gridDim.x) * NG;
  long long row = static_cast<long long>(blockIdx.x) * NG + g;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  uint4 ra0 = z4, ra1 = z4, rb0 = z4, rb1 = z4, rd0 = z4, rd1 = z4;
  float mu = 0.f, rs = 0.f;
  if (row < rows) {
    if (active) {
      ld_raw8(dy, dy_dtype, row * D + c0, ra0, ra1);
      ld_raw8(x, x_dtype, row * D + c0, rb0, rb1);
      if (has_res) ld_raw8(dres, MTASR_DT_F32, row * D + c0, rd0, rd1);
      if (has_gate) gab = *reinterpret_cast<const float2*>(gate_ab + (row * gH + ghead) * 2);
    }
    mu = mean[row];
    rs = rstd[row];
  }
  const float inv_d = 1.0f / static_cast<float>(D);
  int par = 0;
  for (; row < rows; row += stride, par ^= 1) {
    float a[8], b[8], d[8];
    cvt_raw8(dy_dtype, ra0, ra1, a);
    cvt_raw8(x_dtype, rb0, rb1, b);
    cvt_raw8(MTASR_DT_F32, rd0, rd1, d);
    const float mu_c = mu, rs_c = rs;
    if (has_gate) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += fmaf(gab.x, gwa[j], gab.y * gwb[j]);
    }
    const long long nxt = row + stride;
    if (nxt < rows) {                               // next row's loads in flight under this row's reductions
      if (active) {
        ld_raw8(dy, dy_dtype, nxt * D + c0, ra0, ra1);
        ld_raw8(x, x_dtype, nxt * D + c0, rb0, rb1);
        if (has_res) ld_raw8(dres, MTASR_DT_F32, nxt * D + c0, rd0, rd1);
        if (has_gate) gab = *reinterpret_cast<const float2*>(gate_ab + (nxt * gH + ghead) * 2);
      }
      mu = mean[nxt];
      rs = rstd[nxt];
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xhat = (b[j] - mu_c) * rs_c;
      const float gg = a[j] * gm[j];
      s1 += gg;
      s2 = fmaf(gg, xhat, s2);
      ab[j] += a[j];
      ag[j] = fmaf(a[j], xhat, ag[j]);
      b[j] = xhat;
      a[j] = gg;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (NW > 1) {
      if (lane == 0) red[par][g][wg] = make_float2(s1, s2);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(G) : "memory");
      s1 = s2 = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const float2 t = red[par][g][w];
        s1 += t.x;
        s2 += t.y;
      }
    }
    s1 *= inv_d;
    s2 *= inv_d;
    if (active) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = rs_c * (a[j] - s1 - b[j] * s2) + d[j];
        ax[j] += o[j];
      }
      if (dx_f32) st8_f32(dx_f32 + row * D + c0, o);
      if (dx_bf16) st8_bf16(dx_bf16 + row * D + c0, o);
    }
  }
  // CTA-level combination of the NG groups' column sums, one atomic per column and quantity
  float* const outs[3] = {dgamma, dbeta, dxsum};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (outs[q] == nullptr) continue;             // uniform
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[g][c0 + j] = q == 0 ? ag[j] : (q == 1 ? ab[j] : ax[j]);
    __syncthreads();
    for (int c = threadIdx.x; c < GC; c += NG * G) {
      if (c < D) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < NG; ++i) t += acc[i][c];
        atomicAdd(outs[q] + c, t);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ cast / colsum
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n8) {
  pdl_trigger();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float t[8];
    ld8(x, MTASR_DT_F32, i * 8, t);
    st8_bf16(y + i * 8, t);
  }
}
__global__ void cast_tail_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long start,
                                 long long n) {
  const long long i = start + threadIdx.x;
  if (i < n) y[i] = f2bf(x[i]);
}

// fp32 -> bf16 operand split for the fp32-accurate GEMM mode.  terms = 3: x = hi + lo + O(2^-18 |x|), every chunk of `c`
// consecutive fp32 values of a row becomes [lo | hi | hi] (order 0, A side) / [hi | lo | hi] (order 1, B side), so ONE
// bf16 GEMM over the 3x longer contraction evaluates a_lo b_hi + a_hi b_lo + a_hi b_hi with fp32 accumulation.
// terms = 6: three-way split x = x1 + x2 + x3 + O(2^-27 |x|), A side [a3|a2|a1|a2|a1|a1], B side [b1|b2|b3|b1|b2|b1]
// = all products down to 2^-24.
__device__ __forceinline__ void split_parts(const float (&t)[8], float (&p1)[8], float (&p2)[8], float (&p3)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    p1[j] = bf2f(f2bf(t[j]));
    const float r1 = t[j] - p1[j];
    p2[j] = bf2f(f2bf(r1));
    p3[j] = r1 - p2[j];
  }
}
// PCGrad projection (ref:src/trainer_seq2seq.py:1116-1124) without host round trips: out2 = {<gi, gj>, <gj, gj>} in one pass, then
// gi -= (dot < 0 ? dot / (norm2 + 1e-12) : 0) * gj with the scalars read from device memory (the reference branches on the
// host: `if dot < 0`, one synchronisation per head pair).
__global__ void __launch_bounds__(256) pcgrad_dots_kernel(const float* __restrict__ gi, const float* __restrict__ gj, long long n,
                                                          float* __restrict__ out2) {
  __shared__ float red[2][33];
  double d = 0.0, q = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float a = gi[i], b = gj[i];
    d += static_cast<double>(a) * b;
    q += static_cast<double>(b) * b;
  }
  const float df = block_sum(static_cast<float>(d), red[0]);
  const float qf = block_sum(static_cast<float>(q), red[1]);
  if (threadIdx.x == 0) { atomicAdd(out2, df); atomicAdd(out2 + 1, qf); }
}
__global__ void __launch_bounds__(256) pcgrad_project_kernel(float* __restrict__ gi, const float* __restrict__ gj, long long n,
                                                             const float* __restrict__ dots2) {
  const float dot = dots2[0];
  if (!(dot < 0.f)) return;
  const float alpha = dot / (dots2[1] + 1e-12f);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    gi[i] -= alpha * gj[i];
}

// stand-alone dropout pass (one thread per element; the fused sites live in the GEMM / attention kernels)
__global__ void dropout_kernel(const void* __restrict__ x, int x_dtype, long long rows, long long cols, DropP d, void* __restrict__ y,
                               int y_dtype) {
  pdl_trigger();
  const uint32_t s0 = __ldg(d.seed), s1 = __ldg(d.seed + 1);
  const long long n = rows * cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    const float v = x_dtype == MTASR_DT_F32 ? reinterpret_cast<const float*>(x)[i] : bf2f(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
    const float o = v * drop_mult(s0, s1, d, static_cast<unsigned long long>(r) * d.ld + c);
    if (y_dtype == MTASR_DT_F32) reinterpret_cast<float*>(y)[i] = o;
    else reinterpret_cast<__nv_bfloat16*>(y)[i] = f2bf(o);
  }
}

__global__ void split_bf16_kernel(const float* __restrict__ x, long long n8, long long c8, int order, int terms,
                                  __nv_bfloat16* __restrict__ y) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float t[8], p1[8], p2[8], p3[8];
    ld8(x, MTASR_DT_F32, i * 8, t);
    split_parts(t, p1, p2, p3);
    const long long chunk = i / c8;
    const long long off = (i % c8) * 8;
    const long long cs = c8 * 8;
    __nv_bfloat16* o = y + chunk * (static_cast<long long>(terms) * cs) + off;
    // small products first: the tensor core adds every K=16 group into the fp32 accumulator with truncation, so the
    // error grows with (number of accumulation steps) x |accumulator|; the low-order terms are accumulated while the
    // accumulator is still ~2^-8 / 2^-16 of its final size and only the x1 y1 chain runs at full magnitude.
    if (terms == 3) {
      st8_bf16(o, order == 0 ? p2 : p1);
      st8_bf16(o + cs, order == 0 ? p1 : p2);
      st8_bf16(o + 2 * cs, p1);
    } else if (order == 0) {
      st8_bf16(o, p3); st8_bf16(o + cs, p2); st8_bf16(o + 2 * cs, p1);
      st8_bf16(o + 3 * cs, p2); st8_bf16(o + 4 * cs, p1); st8_bf16(o + 5 * cs, p1);
    } else {
      st8_bf16(o, p1); st8_bf16(o + cs, p2); st8_bf16(o + 2 * cs, p3);
      st8_bf16(o + 3 * cs, p1); st8_bf16(o + 4 * cs, p2); st8_bf16(o + 5 * cs, p1);
    }
  }
}

// P = exp(logit - lse[row]) * rowscale[row]: fp16 logits (vocab GEMM mode 1) -> bf16 softmax * upstream gradient, plus
// (optionally) its column sums = the bias gradient, so that P is not re-read by a separate reduction.  Pure streaming
// (2 B read + 2 B written per element): CTA (x, y) owns 2048 columns (one 128-bit piece per thread and row) of row group
// y; the 8 column sums of a thread live in registers over its rows and leave with one atomic per column at the end.
__global__ void __launch_bounds__(256)
softmax_from_logits_kernel(const __half* __restrict__ lg, const float* __restrict__ lse, const float* __restrict__ rowscale,
                           long long rows, int V, long long ld, long long rows_per_group, __nv_bfloat16* __restrict__ P,
                           float* __restrict__ colsum) {
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // 8-column piece
  if (i * 8 >= V) return;
  const long long r0 = blockIdx.y * rows_per_group;
  const long long r1 = r0 + rows_per_group < rows ? r0 + rows_per_group : rows;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  constexpr int RU = 4;   // rows in flight per thread: the loads of RU rows are issued before any of them is consumed
  for (long long rb = r0; rb < r1; rb += RU) {
    float rsc[RU], nl[RU];
    uint4 u[RU];
#pragma unroll
    for (int k = 0; k < RU; ++k) {
      const long long row = rb + k;
      rsc[k] = row < r1 ? rowscale[row] : 0.f;
      nl[k] = row < r1 ? -lse[row] * 1.4426950408889634f : 0.f;
    }
#pragma unroll
    for (int k = 0; k < RU; ++k) {
      const long long row = rb + k;
      u[k] = (row < r1 && rsc[k] != 0.f) ? reinterpret_cast<const uint4*>(lg + row * ld)[i] : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < RU; ++k) {
      const long long row = rb + k;
      if (row >= r1) break;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);   // padded frames / infeasible utterances (rowscale 0): zeros, logits not read
      if (rsc[k] != 0.f) {
        const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          const int col = i * 8 + j * 2;
          const float p0 = col < V ? ex2_approx(fmaf(f.x, 1.4426950408889634f, nl[k])) * rsc[k] : 0.f;   // tail columns inside ld: 0
          const float p1 = col + 1 < V ? ex2_approx(fmaf(f.y, 1.4426950408889634f, nl[k])) * rsc[k] : 0.f;
          ow[j] = pack_bf16x2(p0, p1);
          acc[j * 2] += p0;
          acc[j * 2 + 1] += p1;
        }
      }
      reinterpret_cast<uint4*>(P + row * ld)[i] = o;
    }
  }
  if (colsum) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (i * 8 + j < V) atomicAdd(colsum + i * 8 + j, acc[j]);
  }
}

// ------------------------------------------------------------------------------------------------ weight norm (last dim)
// torch.nn.utils.parametrizations.weight_norm(conv, dim=2) of the positional conv (hf:48-66): w[r][c] = g[c] v[r][c] / ||v[:, c]||
// with the norm over all rows r = (out, in) of tap c.  Column reductions over the (R, Kt) matrix + one elementwise pass.
static constexpr int WN_PARTS = 128;   // CTAs (= partial sums per column) of the weight-norm column reductions
template <bool DOT>
__global__ void __launch_bounds__(256)
wn_colreduce_kernel(const float* __restrict__ a, const float* __restrict__ b, long long R, int Kt, float* __restrict__ out) {
  __shared__ float red[256];
  const int c = threadIdx.x % Kt, lr = threadIdx.x / Kt, nr = blockDim.x / Kt;
  float acc = 0.f;
  for (long long r = static_cast<long long>(blockIdx.x) * nr + lr; r < R; r += static_cast<long long>(gridDim.x) * nr) {
    const float x = a[r * Kt + c];
    acc = fmaf(x, DOT ? b[r * Kt + c] : x, acc);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (lr == 0) {
    for (int i = 1; i < nr; ++i) acc += red[i * Kt + c];
    out[static_cast<long long>(blockIdx.x) * Kt + c] = acc;   // per-CTA partial: summed in a fixed order below (deterministic)
  }
}
// out[c] = sum over the CTA partials, fixed order: the normalised weight must not change from call to call (an atomic
// accumulation moved ||v|| by an ulp between calls, which flipped bf16 roundings of the kernel and made the forward
// non-reproducible at the 1e-4 level).
__global__ void wn_partial_sum_kernel(const float* __restrict__ part, int nparts, int Kt, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Kt) return;
  float s = 0.f;
  for (int i = 0; i < nparts; ++i) s += part[static_cast<long long>(i) * Kt + c];
  out[c] = s;
}
// fwd: w = v * g / sqrt(sumsq);  bwd: dv = (g / n) (dw - v dot / n^2), n = sqrt(sumsq), dot = sum_r dw v
__global__ void wn_apply_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ sumsq,
                                const float* __restrict__ dw, const float* __restrict__ dot, long long n, int Kt,
                                float* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Kt);
    const float inv = rsqrtf(sumsq[c]);
    if (dw == nullptr) out[i] = v[i] * g[c] * inv;
    else out[i] = g[c] * inv * (dw[i] - v[i] * dot[c] * inv * inv);
  }
}
__global__ void wn_dg_kernel(const float* __restrict__ dot, const float* __restrict__ sumsq, int Kt, float* __restrict__ dg) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < Kt) dg[c] = dot[c] * rsqrtf(sumsq[c]);
}

// out[n] += sum_m x[m][n]  (bias gradients).  Vector path (N % 8 == 0, 16-byte aligned rows): each thread owns 8
// consecutive columns (one 128-bit load per row for bf16, two for fp32), 8 row-lanes per CTA, 4 rows in flight per
// thread; CTA partials are combined in smem and added with one atomic per column.
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const void* __restrict__ x, int dtype, long long M, int N, long long ld, float* __restrict__ out) {
  pdl_trigger();
  __shared__ float red[8][32][9];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + cx) * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    const long long step = static_cast<long long>(gridDim.y) * 8;
    long long m = static_cast<long long>(blockIdx.y) * 8 + ry;
    for (; m + 3 * step < M; m += 4 * step) {
      float a[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) ld8(x, dtype, (m + u * step) * ld + col, a[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += a[u][j];
    }
    for (; m < M; m += step) {
      float a[8];
      ld8(x, dtype, m * ld + col, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += a[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ry][cx][j] = s[j];
  __syncthreads();
  // 256 threads -> 256 columns of this CTA
  const int c = threadIdx.x;
  const int gcol = blockIdx.x * 256 + c;
  if (gcol < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][c >> 3][c & 7];
    atomicAdd(out + gcol, t);
  }
}

// scalar fallback for odd N / unaligned rows
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, int dtype, long long M, int N, long long ld, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (col < N) {
    for (long long m = static_cast<long long>(blockIdx.y) * 8 + ry; m < M; m += static_cast<long long>(gridDim.y) * 8) {
      s += dtype == MTASR_DT_BF16 ? bf2f(reinterpret_cast<const __nv_bfloat16*>(x)[m * ld + col])
                                  : reinterpret_cast<const float*>(x)[m * ld + col];
    }
  }
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cx];
    atomicAdd(out + col, t);
  }
}

// ------------------------------------------------------------------------------------------------ rel-pos gate
// gru_rel_pos gate of hf:167-176 for every (b, t, head): with x_h the 64-wide head slice of the layer input,
//   [a, b] = view(Linear_{64->8}(x_h), 2, 4).sum(-1) = [wa . x_h + ba, wb . x_h + bb]   (wa/wb = sums of 4 weight rows)
//   gate = sigmoid(a) * (sigmoid(b) * const_h - 1) + 2
// One warp per (b, t) row; lane l holds elements l and l+32 of each head slice.  wab = [wa | wb] (128 floats),
// bab = [ba, bb].  Output gate (B,H,T) fp32.
// One warp per (b, t) row, two lanes per head: lane l owns the contiguous 32 elements [32 (l & 1), +32) of head l >> 1
// (four 128-bit loads for bf16, eight for fp32), so a head's dot products are one lane-local accumulation plus a single
// shuffle with the partner lane.  H <= 16.  The summed weights live in shared memory (broadcast reads).
__device__ __forceinline__ void gate_load32(const void* x, int x_dtype, long long off, float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float t[8];
    ld8(x, x_dtype, off + c * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[c * 8 + j] = t[j];
  }
}

__global__ void __launch_bounds__(256)
relpos_gate_fwd_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ w8, const float* __restrict__ b8,
                       const float* __restrict__ cst, int B, int T, int H, float* __restrict__ gate) {
  pdl_trigger();
  __shared__ __align__(16) float w_s[128];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) w_s[i] = gate_wsum(w8, i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const int h = lane >> 1, half = lane & 1;
  const bool active = h < H;
  const float* wa = w_s + half * 32;
  const float* wb = w_s + 64 + half * 32;
  const float ba = (b8[0] + b8[1]) + (b8[2] + b8[3]), bb = (b8[4] + b8[5]) + (b8[6] + b8[7]);
  const float c_h = active ? cst[h] : 0.f;
  const long long rows = static_cast<long long>(B) * T;
  const int D = H * 64;
  for (long long row = warp0; row < rows; row += nwarps) {
    float a = 0.f, bv = 0.f;
    if (active) {
      float v[32];
      gate_load32(x, x_dtype, row * D + h * 64 + half * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        a = fmaf(v[i], wa[i], a);
        bv = fmaf(v[i], wb[i], bv);
      }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    bv += __shfl_xor_sync(0xffffffffu, bv, 1);
    if (active && half == 0) {
      const float ga = 1.f / (1.f + __expf(-(a + ba))), gb = 1.f / (1.f + __expf(-(bv + bb)));
      const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
      gate[(static_cast<long long>(b) * H + h) * T + t] = ga * (gb * c_h - 1.f) + 2.f;
    }
  }
}

// Backward: dx (B,T,D) fp32 = da*wa + db*wb per head slice; dw8 (8,64), db8 (8), dcst (H) accumulated with atomics
// (zero them first).
__global__ void __launch_bounds__(128, 3)
relpos_gate_bwd_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ w8, const float* __restrict__ b8,
                       const float* __restrict__ cst, const float* __restrict__ dgate, int B, int T, int H,
                       float* __restrict__ dx, float* __restrict__ dab, float* __restrict__ dw8, float* __restrict__ db8,
                       float* __restrict__ dcst) {
  pdl_trigger();
  __shared__ __align__(16) float w_s[128];
  __shared__ float red_s[132];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) w_s[i] = gate_wsum(w8, i);
  for (int i = threadIdx.x; i < 132; i += blockDim.x) red_s[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const int h = lane >> 1, half = lane & 1;
  const bool active = h < H;
  const float* wa = w_s + half * 32;
  const float* wb = w_s + 64 + half * 32;
  const float ba = (b8[0] + b8[1]) + (b8[2] + b8[3]), bb = (b8[4] + b8[5]) + (b8[6] + b8[7]);
  const float c_h = active ? cst[h] : 0.f;
  const long long rows = static_cast<long long>(B) * T;
  const int D = H * 64;
  float dwa[32], dwb[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { dwa[i] = 0.f; dwb[i] = 0.f; }
  float dba = 0.f, dbb = 0.f, dc = 0.f;
  for (long long row = warp0; row < rows; row += nwarps) {
    float v[32];
    float a = 0.f, bv = 0.f;
    if (active) {
      gate_load32(x, x_dtype, row * D + h * 64 + half * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        a = fmaf(v[i], wa[i], a);
        bv = fmaf(v[i], wb[i], bv);
      }
    }
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    bv += __shfl_xor_sync(0xffffffffu, bv, 1);
    if (active) {
      const float ga = 1.f / (1.f + __expf(-(a + ba))), gb = 1.f / (1.f + __expf(-(bv + bb)));
      const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
      const float dg = dgate[(static_cast<long long>(b) * H + h) * T + t];
      const float da = dg * (gb * c_h - 1.f) * ga * (1.f - ga);
      const float db = dg * ga * c_h * gb * (1.f - gb);
      if (half == 0) { dc += dg * ga * gb; dba += da; dbb += db; }
      if (dab != nullptr && half == 0) *reinterpret_cast<float2*>(dab + (row * H + h) * 2) = make_float2(da, db);
      if (dx != nullptr) {
        float* dxp = dx + row * D + h * 64 + half * 32;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 o;
          o.x = fmaf(da, wa[c * 4 + 0], db * wb[c * 4 + 0]);
          o.y = fmaf(da, wa[c * 4 + 1], db * wb[c * 4 + 1]);
          o.z = fmaf(da, wa[c * 4 + 2], db * wb[c * 4 + 2]);
          o.w = fmaf(da, wa[c * 4 + 3], db * wb[c * 4 + 3]);
          reinterpret_cast<float4*>(dxp)[c] = o;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        dwa[i] = fmaf(da, v[i], dwa[i]);
        dwb[i] = fmaf(db, v[i], dwb[i]);
      }
    }
  }
  // lanes of equal parity hold partial sums for the same 32 weight entries: fold them, then one smem atomic per entry
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float sa = dwa[i], sb = dwb[i];
#pragma unroll
    for (int o = 2; o < 32; o <<= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sb += __shfl_xor_sync(0xffffffffu, sb, o);
    }
    if (lane < 2) {
      atomicAdd(red_s + half * 32 + i, sa);
      atomicAdd(red_s + 64 + half * 32 + i, sb);
    }
  }
  dba = warp_sum(dba);
  dbb = warp_sum(dbb);
  if (lane == 0) { atomicAdd(red_s + 128, dba); atomicAdd(red_s + 129, dbb); }
  if (active && half == 0 && dc != 0.f) atomicAdd(dcst + h, dc);
  __syncthreads();
  // the four weight rows (bias entries) that were summed into one share its gradient
  for (int i = threadIdx.x; i < 130; i += blockDim.x) {
    const float sv = red_s[i];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (i < 128) atomicAdd(dw8 + ((i >> 6) * 4 + r) * 64 + (i & 63), sv);
      else atomicAdd(db8 + (i - 128) * 4 + r, sv);
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention softmax
// P[b,h,q,k] = softmax_k( S[b,h,q,k]*scale + gate[b,h,q] * table[h, k-q+T-1] )  for k < klen[b], else 0.
// Replaces the materialised (B*H,T,T) gated position bias of hf:167-180 + the masked softmax inside
// F.multi_head_attention_forward (hf:206-228): the bias is Toeplitz (hf:243-271), so only table (H, 2T-1) and
// gate (B,H,T) are read.  One warp per (b,h,q) row.
__global__ void __launch_bounds__(256)
attn_softmax_fwd_kernel(const float* __restrict__ S, const float* __restrict__ gate, const float* __restrict__ table,
                        const int* __restrict__ klen, int B, int H, int T, int Tp, float scale,
                        __nv_bfloat16* __restrict__ P, int terms) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const long long rows = static_cast<long long>(B) * H * T;
  for (long long row = warp0; row < rows; row += nwarps) {
    const int q = static_cast<int>(row % T);
    const int h = static_cast<int>((row / T) % H);
    const int b = static_cast<int>(row / (static_cast<long long>(T) * H));
    const int kl = klen ? min(klen[b], T) : T;
    const float g = gate[row];
    const float* srow = S + row * Tp;
    const float* trow = table + static_cast<long long>(h) * (2 * T - 1) + (T - 1 - q);
    __nv_bfloat16* prow = P + row * Tp * terms;   // terms 1: plain bf16 P; 3 / 6: split A-side operand (split_bf16_kernel)
    float m = -INFINITY;
    for (int k = lane; k < kl; k += 32) m = fmaxf(m, srow[k] * scale + g * trow[k]);
    m = warp_max(m);
    float s = 0.f;
    for (int k = lane; k < kl; k += 32) s += __expf(srow[k] * scale + g * trow[k] - m);
    s = warp_sum(s);
    const float inv = kl > 0 ? 1.f / s : 0.f;
    for (int k = lane; k < Tp; k += 32) {
      float p = 0.f;
      if (k < kl) p = __expf(srow[k] * scale + g * trow[k] - m) * inv;
      const __nv_bfloat16 p1 = f2bf(p);
      prow[k] = p1;
      if (terms > 1) {   // A-side operand of the fp32-accurate P V contraction
        const float r1 = p - bf2f(p1);
        const __nv_bfloat16 p2 = f2bf(r1);
        if (terms == 3) {            // [lo | hi | hi]
          prow[k] = p2;
          prow[Tp + k] = p1;
          prow[2 * Tp + k] = p1;
        } else {                     // [p3 | p2 | p1 | p2 | p1 | p1]
          prow[k] = f2bf(r1 - bf2f(p2));
          prow[Tp + k] = p2;
          prow[2 * Tp + k] = p1;
          prow[3 * Tp + k] = p2;
          prow[4 * Tp + k] = p1;
          prow[5 * Tp + k] = p1;
        }
      }
    }
  }
}

// dS = P * (dP - sum_k P dP) (bf16 out, already multiplied by `scale` for the dQ/dK contractions);
// dgate[b,h,q] = sum_k dZ * table; dtable[h, k-q+T-1] += dZ * gate (smem accumulation per CTA, CTAs are per-head).
__global__ void __launch_bounds__(256)
attn_softmax_bwd_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dP,
                        const float* __restrict__ gate, const float* __restrict__ table, int B, int H, int T, int Tp,
                        float scale, int rows_per_cta, __nv_bfloat16* __restrict__ dS, float* __restrict__ dgate,
                        float* __restrict__ dtable) {
  extern __shared__ float acc[];  // 2T-1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int h = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * T - 1; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const long long bq0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long bq_end = min(bq0 + rows_per_cta, static_cast<long long>(B) * T);
  const float* trow_h = table + static_cast<long long>(h) * (2 * T - 1);
  for (long long bq = bq0 + warp; bq < bq_end; bq += nw) {
    const int b = static_cast<int>(bq / T), q = static_cast<int>(bq % T);
    const long long row = (static_cast<long long>(b) * H + h) * T + q;
    const __nv_bfloat16* prow = P + row * Tp;
    const float* dprow = dP + row * Tp;
    __nv_bfloat16* dsrow = dS + row * Tp;
    const float g = gate[row];
    // P was rounded to bf16, so its row no longer sums to 1 exactly: normalising the row dot by sum(P) keeps
    // sum_k dZ[k] == 0 (the exact softmax backward of slightly perturbed logits) instead of leaking a rounding bias
    // into the heavily cancelling dgate / dtable reductions.
    float dot = 0.f, psum = 0.f;
    for (int k = lane; k < T; k += 32) {
      const float pk = bf2f(prow[k]);
      dot += pk * dprow[k];
      psum += pk;
    }
    dot = warp_sum(dot);
    psum = warp_sum(psum);
    dot = psum > 0.f ? dot / psum : 0.f;
    float dg = 0.f;
    const float* trow = trow_h + (T - 1 - q);
    float* arow = acc + (T - 1 - q);
    for (int k = lane; k < Tp; k += 32) {
      float dz = 0.f;
      if (k < T) {
        dz = bf2f(prow[k]) * (dprow[k] - dot);
        dg += dz * trow[k];
        atomicAdd(arow + k, dz * g);
      }
      dsrow[k] = f2bf(dz * scale);
    }
    dg = warp_sum(dg);
    if (lane == 0) dgate[row] = dg;
  }
  __syncthreads();
  float* dt = dtable + static_cast<long long>(h) * (2 * T - 1);
  for (int i = threadIdx.x; i < 2 * T - 1; i += blockDim.x)
    if (acc[i] != 0.f) atomicAdd(dt + i, acc[i]);
}

// ------------------------------------------------------------------------------------------------ pad / GLU / mask
// y[b][pad_l + t][:] = bf16(x[b][t][:]) with zero rows on both sides (pos-conv / adapter "same" padding).
__global__ void pad_cast_kernel(const void* __restrict__ x, int x_dtype, int B, int T, int D, int pad_l, int Tpad,
                                const int* __restrict__ vlen, __nv_bfloat16* __restrict__ y) {
  pdl_trigger();
  const long long n8 = static_cast<long long>(B) * Tpad * (D >> 3);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (D >> 3));
    const long long bt = i / (D >> 3);
    const int tp = static_cast<int>(bt % Tpad), b = static_cast<int>(bt / Tpad);
    const int t = tp - pad_l;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (t >= 0 && t < T && (!vlen || t < vlen[b])) ld8(x, x_dtype, (static_cast<long long>(b) * T + t) * D + c * 8, v);
    st8_bf16(y + i * 8, v);
  }
}

// GLU over the channel halves of a channels-last row: y[r][c] = a[r][c] * sigmoid(a[r][C + c]).
__global__ void glu_fwd_kernel(const void* __restrict__ x, int x_dtype, long long rows, int C,
                               __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32) {
  pdl_trigger();
  const long long n8 = rows * (C >> 3);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (C >> 3));
    const long long r = i / (C >> 3);
    float a[8], g[8], o[8];
    ld8(x, x_dtype, r * 2 * C + c * 8, a);
    ld8(x, x_dtype, r * 2 * C + C + c * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a[j] * sigmoid_f(g[j]);
    if (y_bf16) st8_bf16(y_bf16 + r * C + c * 8, o);
    if (y_f32) st8_f32(y_f32 + r * C + c * 8, o);
  }
}
// dx[r][c] = dy * sig(g);  dx[r][C+c] = dy * a * sig(g) * (1 - sig(g))
__global__ void glu_bwd_kernel(const void* __restrict__ x, int x_dtype, const void* __restrict__ dy, int dy_dtype,
                               long long rows, int C, __nv_bfloat16* __restrict__ dx) {
  const long long n8 = rows * (C >> 3);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (C >> 3));
    const long long r = i / (C >> 3);
    float a[8], g[8], d[8], o1[8], o2[8];
    ld8(x, x_dtype, r * 2 * C + c * 8, a);
    ld8(x, x_dtype, r * 2 * C + C + c * 8, g);
    ld8(dy, dy_dtype, r * C + c * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = sigmoid_f(g[j]);
      o1[j] = d[j] * s;
      o2[j] = d[j] * a[j] * s * (1.f - s);
    }
    st8_bf16(dx + r * 2 * C + c * 8, o1);
    st8_bf16(dx + r * 2 * C + C + c * 8, o2);
  }
}

// du = dy * act'(src): act 3 = GELU'(pre-activation), act 4 = ReLU (src = activation output).
__global__ void act_bwd_kernel(const void* __restrict__ dy, int dy_dtype, const __nv_bfloat16* __restrict__ src, int act,
                               long long n8, __nv_bfloat16* __restrict__ du) {
  pdl_trigger();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float d[8], u[8], o[8];
    ld8(dy, dy_dtype, i * 8, d);
    ld8(src, MTASR_DT_BF16, i * 8, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = act == 3 ? d[j] * gelu_grad_f(u[j]) : (u[j] > 0.f ? d[j] : 0.f);
    st8_bf16(du + i * 8, o);
  }
}

static int grid_for(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace mtasr

using namespace mtasr;

extern "C" int mtasr_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, float eps,
                                   int64_t rows, int32_t D, int32_t post_gelu, void* y_bf16, float* y_f32, float* mean,
                                   float* rstd, void* stream) {
  MTASR_CHECK_ARG(x && gamma && beta && rows > 0 && (y_bf16 || y_f32), "layernorm_fwd: bad arguments");
  MTASR_CHECK_ARG(D % 8 == 0 && D <= 32 * MAXV, "layernorm_fwd: D=%d must be a multiple of 8 and <= 1024", D);
  layernorm_fwd_kernel<<<grid_for(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, gamma, beta, eps, rows, D, post_gelu, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32, mean, rstd);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("layernorm_fwd");
  return MTASR_OK;
}

static int layernorm_bwd_launch(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* mean,
                                const float* rstd, const float* gamma, const float* dres, int64_t rows, int32_t D,
                                float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, const float* gate_ab,
                                const float* gate_w8, cudaStream_t st) {
  MTASR_CHECK_ARG(dy && x && mean && rstd && gamma && rows > 0 && (dx_f32 || dx_bf16 || dgamma || dbeta), "layernorm_bwd: bad arguments");
  MTASR_CHECK_ARG((gate_ab == nullptr) == (gate_w8 == nullptr), "layernorm_bwd: gate_ab and gate_w8 come together");
  MTASR_CHECK_ARG(!gate_ab || (D % 64 == 0 && (dx_f32 || dx_bf16)), "layernorm_bwd: the gate path needs D %% 64 == 0 and a dx output");
  MTASR_CHECK_ARG(D % 8 == 0 && D <= 32 * MAXV, "layernorm_bwd: D=%d must be a multiple of 8 and <= 1024", D);
  MTASR_CHECK_ARG(!dxsum || dx_f32 || dx_bf16, "layernorm_bwd: dxsum needs a dx output");
  if (dx_f32 || dx_bf16) {
    // one pass: dx, dgamma, dbeta and the column sums of dx (all three reductions accumulate into zeroed buffers)
    const int nw = (D / 8 + 31) / 32;
    const int ng = nw == 3 ? 2 : 8 / nw;
    long long g = (rows + ng - 1) / ng;
    if (g > num_sms() * 2) g = num_sms() * 2;
    const unsigned grid = static_cast<unsigned>(g), block = static_cast<unsigned>(ng * nw * 32);
    __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
#define MTASR_LNB(NW) layernorm_bwd_fused_kernel<NW><<<grid, block, 0, st>>>(dy, dy_dtype, x, x_dtype, mean, rstd, gamma, dres, rows, D, dx_f32, dxb, dgamma, dbeta, dxsum, gate_ab, gate_w8)
    switch (nw) {
      case 1: MTASR_LNB(1); break;
      case 2: MTASR_LNB(2); break;
      case 3: MTASR_LNB(3); break;
      default: MTASR_LNB(4); break;
    }
#undef MTASR_LNB
    MTASR_COUNT_LAUNCH();
  } else {
    const int gx = (D + 255) / 256;
    int gy = static_cast<int>((rows + 63) / 64);
    const int cap = (num_sms() * 4 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    layernorm_bwd_param_kernel<<<dim3(gx, gy), 256, 0, st>>>(dy, dy_dtype, x, x_dtype, mean, rstd, rows, D, dgamma, dbeta);
    MTASR_COUNT_LAUNCH();
  }
  MTASR_CHECK_LAUNCH("layernorm_bwd");
  return MTASR_OK;
}

extern "C" int mtasr_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* mean,
                                   const float* rstd, const float* gamma, const float* dres, int64_t rows, int32_t D,
                                   float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, void* stream) {
  return layernorm_bwd_launch(dy, dy_dtype, x, x_dtype, mean, rstd, gamma, dres, rows, D, dx_f32, dx_bf16, dgamma, dbeta, nullptr,
                              nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int mtasr_layernorm_bwd_sums(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* mean,
                                        const float* rstd, const float* gamma, const float* dres, int64_t rows, int32_t D,
                                        float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, const float* gate_ab,
                                        const float* gate_w8, void* stream) {
  return layernorm_bwd_launch(dy, dy_dtype, x, x_dtype, mean, rstd, gamma, dres, rows, D, dx_f32, dx_bf16, dgamma, dbeta, dxsum,
                              gate_ab, gate_w8, static_cast<cudaStream_t>(stream));
}

extern "C" int mtasr_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
  MTASR_CHECK_ARG(x && y && n > 0, "cast_f32_bf16: bad arguments");
  MTASR_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "cast: unaligned pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n8 = n / 8;
  if (n8 > 0) {
    cast_f32_bf16_kernel<<<grid_for(n8, 256), 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n8);
    MTASR_COUNT_LAUNCH();
  }
  if (n8 * 8 < n) {
    cast_tail_kernel<<<1, 32, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n8 * 8, n);
    MTASR_COUNT_LAUNCH();
  }
  MTASR_CHECK_LAUNCH("cast_f32_bf16");
  return MTASR_OK;
}

extern "C" int mtasr_softmax_from_logits(const void* logits_f16, const float* lse, const float* rowscale, int64_t rows, int32_t V,
                                         int64_t ld, void* P_bf16, float* colsum, void* stream) {
  MTASR_CHECK_ARG(logits_f16 && lse && rowscale && P_bf16 && rows > 0 && V > 0 && ld >= V && ld % 8 == 0, "softmax_from_logits: bad arguments");
  MTASR_CHECK_ARG((reinterpret_cast<uintptr_t>(logits_f16) & 15) == 0 && (reinterpret_cast<uintptr_t>(P_bf16) & 15) == 0,
                  "softmax_from_logits: unaligned pointer");
  const int col_blocks = ((V + 7) / 8 + 255) / 256;
  long long groups = (static_cast<long long>(num_sms()) * 8 + col_blocks - 1) / col_blocks;
  if (groups > rows) groups = rows;
  if (groups > 65535) groups = 65535;
  const long long rpg = (rows + groups - 1) / groups;
  groups = (rows + rpg - 1) / rpg;
  softmax_from_logits_kernel<<<dim3(col_blocks, static_cast<unsigned>(groups)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(logits_f16), lse, rowscale, rows, V, ld, rpg, reinterpret_cast<__nv_bfloat16*>(P_bf16), colsum);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("softmax_from_logits");
  return MTASR_OK;
}

extern "C" int mtasr_weightnorm_fwd(const float* v, const float* g, int64_t R, int32_t Kt, float* w, float* sumsq, void* stream) {
  MTASR_CHECK_ARG(v && g && w && sumsq && R > 0 && Kt > 0 && Kt <= 256 && 256 % Kt == 0, "weightnorm_fwd: need Kt | 256 (Kt=%d)", Kt);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // sumsq: (1 + WN_PARTS) * Kt floats; [0, Kt) receives the result, the rest holds the per-CTA partials
  wn_colreduce_kernel<false><<<WN_PARTS, 256, 0, st>>>(v, nullptr, R, Kt, sumsq + Kt);
  MTASR_COUNT_LAUNCH();
  wn_partial_sum_kernel<<<(Kt + 127) / 128, 128, 0, st>>>(sumsq + Kt, WN_PARTS, Kt, sumsq);
  MTASR_COUNT_LAUNCH();
  wn_apply_kernel<<<grid_for(R * Kt, 256), 256, 0, st>>>(v, g, sumsq, nullptr, nullptr, R * Kt, Kt, w);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("weightnorm_fwd");
  return MTASR_OK;
}

extern "C" int mtasr_weightnorm_bwd(const float* dw, const float* v, const float* g, const float* sumsq, int64_t R, int32_t Kt,
                                    float* dv, float* dg, float* dot, void* stream) {
  MTASR_CHECK_ARG(dw && v && g && sumsq && dv && dg && dot && R > 0 && Kt > 0 && Kt <= 256 && 256 % Kt == 0, "weightnorm_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  wn_colreduce_kernel<true><<<WN_PARTS, 256, 0, st>>>(dw, v, R, Kt, dot + Kt);
  MTASR_COUNT_LAUNCH();
  wn_partial_sum_kernel<<<(Kt + 127) / 128, 128, 0, st>>>(dot + Kt, WN_PARTS, Kt, dot);
  MTASR_COUNT_LAUNCH();
  wn_apply_kernel<<<grid_for(R * Kt, 256), 256, 0, st>>>(v, g, sumsq, dw, dot, R * Kt, Kt, dv);
  MTASR_COUNT_LAUNCH();
  wn_dg_kernel<<<(Kt + 127) / 128, 128, 0, st>>>(dot, sumsq, Kt, dg);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("weightnorm_bwd");
  return MTASR_OK;
}

extern "C" int mtasr_colsum(const void* x, int32_t dtype, int64_t M, int32_t N, int64_t ld, float* out, void* stream) {
  MTASR_CHECK_ARG(x && out && M > 0 && N > 0 && ld >= N, "colsum: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(out, 0, sizeof(float) * N, st) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "colsum: memset failed");
  const int esz = dtype == MTASR_DT_BF16 ? 2 : 4;
  const bool vec = N % 8 == 0 && (ld * esz) % 16 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
  const int gx = vec ? (N + 255) / 256 : (N + 31) / 32;
  int gy = static_cast<int>((M + 63) / 64);
  const int cap = (num_sms() * 4 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  if (vec) colsum_vec_kernel<<<dim3(gx, gy), 256, 0, st>>>(x, dtype, M, N, ld, out);
  else colsum_kernel<<<dim3(gx, gy), 256, 0, st>>>(x, dtype, M, N, ld, out);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("colsum");
  return MTASR_OK;
}

extern "C" int mtasr_relpos_gate_fwd(const void* x, int32_t x_dtype, const float* w8, const float* b8, const float* cst,
                                     int32_t B, int32_t T, int32_t H, float* gate, void* stream) {
  const float *wab = w8, *bab = b8;
  MTASR_CHECK_ARG(x && wab && bab && cst && gate && B > 0 && T > 0 && H > 0 && H <= 32, "relpos_gate_fwd: bad arguments");
  MTASR_CHECK_ARG(H <= 16, "relpos_gate_fwd: H=%d > 16 heads not supported", H);
  relpos_gate_fwd_kernel<<<grid_for(static_cast<long long>(B) * T, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, wab, bab, cst, B, T, H, gate);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("relpos_gate_fwd");
  return MTASR_OK;
}

extern "C" int mtasr_relpos_gate_bwd(const void* x, int32_t x_dtype, const float* w8, const float* b8, const float* cst,
                                     const float* dgate, int32_t B, int32_t T, int32_t H, float* dx, float* dab, float* dw8,
                                     float* db8, float* dcst, void* stream) {
  const float *wab = w8, *bab = b8;
  float *dwab = dw8, *dbab = db8;
  MTASR_CHECK_ARG(x && wab && bab && cst && dgate && (dx || dab) && dwab && dbab && dcst && B > 0 && T > 0 && H > 0 && H <= 32,
                  "relpos_gate_bwd: bad arguments");
  MTASR_CHECK_ARG(H <= 16, "relpos_gate_bwd: H=%d > 16 heads not supported", H);
  long long g = (static_cast<long long>(B) * T + 15) / 16;
  if (g > num_sms() * 3) g = num_sms() * 3;
  relpos_gate_bwd_kernel<<<static_cast<unsigned>(g), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, wab, bab, cst, dgate, B, T, H, dx, dab, dwab, dbab, dcst);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("relpos_gate_bwd");
  return MTASR_OK;
}

extern "C" int mtasr_attn_softmax_fwd(const float* S, const float* gate, const float* table, const int32_t* klen,
                                      int32_t B, int32_t H, int32_t T, int32_t Tp, float scale, void* P, void* stream) {
  MTASR_CHECK_ARG(S && gate && table && P && B > 0 && H > 0 && T > 0 && Tp >= T, "attn_softmax_fwd: bad arguments");
  const long long rows = static_cast<long long>(B) * H * T;
  attn_softmax_fwd_kernel<<<grid_for(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      S, gate, table, klen, B, H, T, Tp, scale, reinterpret_cast<__nv_bfloat16*>(P), 1);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("attn_softmax_fwd");
  return MTASR_OK;
}

extern "C" int mtasr_attn_softmax_fwd_split(const float* S, const float* gate, const float* table, const int32_t* klen,
                                            int32_t B, int32_t H, int32_t T, int32_t Tp, float scale, int32_t terms,
                                            void* Ps, void* stream) {
  MTASR_CHECK_ARG(S && gate && table && Ps && B > 0 && H > 0 && T > 0 && Tp >= T && (terms == 3 || terms == 6),
                  "attn_softmax_fwd_split: bad arguments");
  const long long rows = static_cast<long long>(B) * H * T;
  attn_softmax_fwd_kernel<<<grid_for(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      S, gate, table, klen, B, H, T, Tp, scale, reinterpret_cast<__nv_bfloat16*>(Ps), terms);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("attn_softmax_fwd_split");
  return MTASR_OK;
}

extern "C" int mtasr_pcgrad_dots(const float* gi, const float* gj, int64_t n, float* out2, void* stream) {
  MTASR_CHECK_ARG(gi && gj && out2 && n > 0, "pcgrad_dots: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(out2, 0, 2 * sizeof(float), st) != cudaSuccess) return set_error(MTASR_ERR_LAUNCH, "pcgrad_dots: memset failed");
  pcgrad_dots_kernel<<<num_sms() * 4, 256, 0, st>>>(gi, gj, n, out2);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("pcgrad_dots");
  return MTASR_OK;
}

extern "C" int mtasr_pcgrad_project(float* gi, const float* gj, int64_t n, const float* dots2, void* stream) {
  MTASR_CHECK_ARG(gi && gj && dots2 && n > 0, "pcgrad_project: bad arguments");
  pcgrad_project_kernel<<<num_sms() * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(gi, gj, n, dots2);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("pcgrad_project");
  return MTASR_OK;
}

extern "C" int mtasr_dropout(const void* x, int32_t x_dtype, int64_t rows, int64_t cols, const void* seed, uint32_t site,
                             uint32_t keep16, void* y, int32_t y_dtype, void* stream) {
  MTASR_CHECK_ARG(x && y && seed && rows > 0 && cols > 0, "dropout: bad arguments");
  MTASR_CHECK_ARG(keep16 > 0 && keep16 <= 65536, "dropout: keep16 must be in (0, 65536]");
  MTASR_CHECK_ARG((x_dtype == MTASR_DT_F32 || x_dtype == MTASR_DT_BF16) && (y_dtype == MTASR_DT_F32 || y_dtype == MTASR_DT_BF16),
                  "dropout: dtypes must be f32 / bf16");
  DropP d;
  d.seed = reinterpret_cast<const uint32_t*>(seed);
  d.site = site;
  d.thresh = keep16;
  d.scale = 65536.0f / static_cast<float>(keep16);
  d.ld = (cols + 1) & ~1LL;
  const long long n = rows * cols;
  dropout_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x_dtype, rows, cols, d, y, y_dtype);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("dropout");
  return MTASR_OK;
}

extern "C" int mtasr_split_bf16(const float* x, int64_t n, int64_t c, int32_t order, int32_t terms, void* y_bf16, void* stream) {
  MTASR_CHECK_ARG(x && y_bf16 && n > 0 && c > 0 && c % 8 == 0 && n % c == 0 && (order == 0 || order == 1) && (terms == 3 || terms == 6),
                  "split_bf16: n=%lld must be a multiple of the chunk c=%lld, c a multiple of 8, terms 3 or 6", static_cast<long long>(n),
                  static_cast<long long>(c));
  MTASR_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_bf16) & 15) == 0, "split_bf16: unaligned pointer");
  const long long n8 = n / 8;
  split_bf16_kernel<<<grid_for(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n8, c / 8, order, terms,
                                                                                       reinterpret_cast<__nv_bfloat16*>(y_bf16));
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("split_bf16");
  return MTASR_OK;
}

extern "C" int mtasr_attn_softmax_bwd(const void* P, const float* dP, const float* gate, const float* table, int32_t B,
                                      int32_t H, int32_t T, int32_t Tp, float scale, void* dS, float* dgate,
                                      float* dtable, void* stream) {
  MTASR_CHECK_ARG(P && dP && gate && table && dS && dgate && dtable && B > 0 && H > 0 && T > 0 && Tp >= T, "attn_softmax_bwd: bad arguments");
  const int rows_per_cta = 64;
  const long long bq = static_cast<long long>(B) * T;
  dim3 grid(static_cast<unsigned>((bq + rows_per_cta - 1) / rows_per_cta), H);
  const size_t smem = sizeof(float) * (2 * T - 1);
  MTASR_CHECK_ARG(smem <= 48 * 1024, "attn_softmax_bwd: T=%d too long for the smem table accumulator", T);
  attn_softmax_bwd_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(P), dP, gate, table, B, H, T, Tp, scale, rows_per_cta,
      reinterpret_cast<__nv_bfloat16*>(dS), dgate, dtable);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("attn_softmax_bwd");
  return MTASR_OK;
}

extern "C" int mtasr_pad_cast(const void* x, int32_t x_dtype, int32_t B, int32_t T, int32_t D, int32_t pad_l,
                              int32_t Tpad, const int32_t* vlen, void* y, void* stream) {
  MTASR_CHECK_ARG(x && y && B > 0 && T > 0 && D % 8 == 0 && Tpad >= T + pad_l, "pad_cast: bad arguments");
  const long long n8 = static_cast<long long>(B) * Tpad * (D >> 3);
  pad_cast_kernel<<<grid_for(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, B, T, D, pad_l, Tpad, vlen, reinterpret_cast<__nv_bfloat16*>(y));
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("pad_cast");
  return MTASR_OK;
}

extern "C" int mtasr_glu_fwd(const void* x, int32_t x_dtype, int64_t rows, int32_t C, void* y_bf16, float* y_f32,
                             void* stream) {
  MTASR_CHECK_ARG(x && (y_bf16 || y_f32) && rows > 0 && C % 8 == 0, "glu_fwd: bad arguments");
  glu_fwd_kernel<<<grid_for(rows * (C >> 3), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, rows, C, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("glu_fwd");
  return MTASR_OK;
}

extern "C" int mtasr_glu_bwd(const void* x, int32_t x_dtype, const void* dy, int32_t dy_dtype, int64_t rows, int32_t C,
                             void* dx_bf16, void* stream) {
  MTASR_CHECK_ARG(x && dy && dx_bf16 && rows > 0 && C % 8 == 0, "glu_bwd: bad arguments");
  glu_bwd_kernel<<<grid_for(rows * (C >> 3), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_dtype, dy, dy_dtype, rows, C, reinterpret_cast<__nv_bfloat16*>(dx_bf16));
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("glu_bwd");
  return MTASR_OK;
}

extern "C" int mtasr_act_bwd(const void* dy, int32_t dy_dtype, const void* src_bf16, int32_t act, int64_t n, void* du_bf16,
                             void* stream) {
  MTASR_CHECK_ARG(dy && src_bf16 && du_bf16 && n > 0 && n % 8 == 0 && (act == 3 || act == 4), "act_bwd: bad arguments");
  act_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, dy_dtype, reinterpret_cast<const __nv_bfloat16*>(src_bf16), act, n / 8, reinterpret_cast<__nv_bfloat16*>(du_bf16));
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("act_bwd");
  return MTASR_OK;
}
