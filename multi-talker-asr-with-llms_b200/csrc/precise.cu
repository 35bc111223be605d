// fp32 parity mode ("precise") -- the pieces that are NOT contractions on split bf16 operands (those run on the tcgen05
// GEMM of gemm.cu, see precise.py): the LSTM recurrence of the separator in plain fp32 FMA arithmetic, the ReLU backward
// and the dense softmax term of the CTC-head gradient in fp32.
//
// The reference keeps the separator / CTC head / loss in fp32 (ref:models/losses.py:265-268, ref:models/ctc.py:53,
// ref:inference_asr.py:120).  The throughput path runs the recurrence with bf16 weights on mma.sync (lstm.cu); this file is
// the validation path behind `mtasr_b200.precise.set_precision("fp32")`: one launch per time step (the grid boundary is
// the step barrier), CUDA-core fp32 dot products, expf / tanhf in full precision.  It is not tuned: ~30 us per step.
#include "common.cuh"

namespace mtasr {

// ---------------------------------------------------------------------------------------------- LSTM, forward step
// gates = xg[b][t] + Whh h_{t-1}[b];  i,f,o = sigmoid, g = tanh;  c = f c_prev + i g;  h = o tanh(c)
// (ref:models/separator.py:6-24; gate order i, f, g, o).  grid (Hs / 8, B), 8 warps: warp w owns hidden unit 8 bx + w.
__global__ void __launch_bounds__(256)
lstm_step_fwd_f32_kernel(const float* __restrict__ xg, const float* __restrict__ whh, int ldw, int T, int Hs, int t,
                         float* __restrict__ h_all, float* __restrict__ c_all, float* __restrict__ gates_act) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int u = blockIdx.x * 8 + warp;
  if (u >= Hs) return;
  const long long bt = static_cast<long long>(b) * T + t;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (t > 0) {
    const float* hp = h_all + (bt - 1) * Hs;
    for (int k = lane; k < Hs; k += 32) {
      const float hv = hp[k];
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[g] = fmaf(whh[static_cast<long long>(g * Hs + u) * ldw + k], hv, acc[g]);
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) acc[g] = warp_sum(acc[g]);
  if (lane == 0) {
    const float* x = xg + bt * 4 * Hs;
    const float gi = 1.f / (1.f + expf(-(acc[0] + x[u])));
    const float gf = 1.f / (1.f + expf(-(acc[1] + x[Hs + u])));
    const float gg = tanhf(acc[2] + x[2 * Hs + u]);
    const float go = 1.f / (1.f + expf(-(acc[3] + x[3 * Hs + u])));
    const float cp = t > 0 ? c_all[(bt - 1) * Hs + u] : 0.f;
    const float c = gf * cp + gi * gg;
    c_all[bt * Hs + u] = c;
    h_all[bt * Hs + u] = go * tanhf(c);
    float* ga = gates_act + bt * 4 * Hs;
    ga[u] = gi; ga[Hs + u] = gf; ga[2 * Hs + u] = gg; ga[3 * Hs + u] = go;
  }
}

// ---------------------------------------------------------------------------------------------- LSTM, backward step
// dh = dh_out[b][t] + Whh^T dgates[b][t+1];  dc = dh o (1 - tanh(c)^2) + dc_carry;  pre-activation gate gradients -> dgates[b][t];
// dc_carry <- dc f.  grid (ceil(Hs / 64), B), block (64 units, 4 gate blocks of the reduction).
__global__ void __launch_bounds__(256)
lstm_step_bwd_f32_kernel(const float* __restrict__ dh_out, const float* __restrict__ whh, int ldw, int T, int Hs, int t,
                         const float* __restrict__ c_all, const float* __restrict__ gates_act, float* __restrict__ dgates,
                         float* __restrict__ dc_carry) {
  __shared__ float red[4][64];
  const int b = blockIdx.y;
  const int kk = threadIdx.x & 63, jp = threadIdx.x >> 6;
  const int u = blockIdx.x * 64 + kk;
  const long long bt = static_cast<long long>(b) * T + t;
  float acc = 0.f;
  if (t + 1 < T && u < Hs) {
    const float* dgn = dgates + (bt + 1) * 4 * Hs + jp * Hs;
    const float* w = whh + static_cast<long long>(jp) * Hs * ldw + u;
#pragma unroll 8
    for (int j = 0; j < Hs; ++j) acc = fmaf(dgn[j], w[static_cast<long long>(j) * ldw], acc);
  }
  red[jp][kk] = acc;
  __syncthreads();
  if (jp != 0 || u >= Hs) return;
  const float dh = dh_out[bt * Hs + u] + red[0][kk] + red[1][kk] + red[2][kk] + red[3][kk];
  const float* ga = gates_act + bt * 4 * Hs;
  const float gi = ga[u], gf = ga[Hs + u], gg = ga[2 * Hs + u], go = ga[3 * Hs + u];
  const float c = c_all[bt * Hs + u];
  const float cp = t > 0 ? c_all[(bt - 1) * Hs + u] : 0.f;
  const float tc = tanhf(c);
  const float dc = fmaf(dh * go, 1.f - tc * tc, (t + 1 < T) ? dc_carry[static_cast<long long>(b) * Hs + u] : 0.f);
  float* dg = dgates + bt * 4 * Hs;
  dg[u] = dc * gg * gi * (1.f - gi);
  dg[Hs + u] = dc * cp * gf * (1.f - gf);
  dg[2 * Hs + u] = dc * gi * (1.f - gg * gg);
  dg[3 * Hs + u] = dh * tc * go * (1.f - go);
  dc_carry[static_cast<long long>(b) * Hs + u] = dc * gf;
}

__global__ void relu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, float* __restrict__ du) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    du[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// out[r][c] = exp(logits[r][c] - lse[r]) * rowscale[r] for c < V, 0 for V <= c < ld.  One CTA per row.
__global__ void __launch_bounds__(256)
softmax_scale_f32_kernel(const float* __restrict__ logits, const float* __restrict__ lse, const float* __restrict__ rowscale,
                         int V, long long ld, float* __restrict__ out) {
  const long long r = blockIdx.x;
  const float l = lse[r], s = rowscale[r];
  const float* src = logits + r * ld;
  float* dst = out + r * ld;
  for (long long c = threadIdx.x; c < ld; c += blockDim.x) dst[c] = c < V ? expf(src[c] - l) * s : 0.f;
}

}  // namespace mtasr

using namespace mtasr;

extern "C" int mtasr_lstm_fwd_f32(const float* xg, const float* whh, int32_t ldw, int32_t B, int32_t T, int32_t Hs, float* h_all,
                                  float* c_all, float* gates_act, void* stream) {
  MTASR_CHECK_ARG(xg && whh && h_all && c_all && gates_act && B > 0 && T > 0 && Hs > 0 && ldw >= Hs, "lstm_fwd_f32: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int t = 0; t < T; ++t) {
    lstm_step_fwd_f32_kernel<<<dim3((Hs + 7) / 8, B), 256, 0, st>>>(xg, whh, ldw, T, Hs, t, h_all, c_all, gates_act);
    MTASR_COUNT_LAUNCH();
  }
  MTASR_CHECK_LAUNCH("lstm_fwd_f32");
  return MTASR_OK;
}

extern "C" int mtasr_lstm_bwd_f32(const float* dh_out, const float* whh, int32_t ldw, int32_t B, int32_t T, int32_t Hs,
                                  const float* c_all, const float* gates_act, float* dgates, float* dc_carry, void* stream) {
  MTASR_CHECK_ARG(dh_out && whh && c_all && gates_act && dgates && dc_carry && B > 0 && T > 0 && Hs > 0 && ldw >= Hs,
                  "lstm_bwd_f32: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int t = T - 1; t >= 0; --t) {
    lstm_step_bwd_f32_kernel<<<dim3((Hs + 63) / 64, B), 256, 0, st>>>(dh_out, whh, ldw, T, Hs, t, c_all, gates_act, dgates, dc_carry);
    MTASR_COUNT_LAUNCH();
  }
  MTASR_CHECK_LAUNCH("lstm_bwd_f32");
  return MTASR_OK;
}

extern "C" int mtasr_relu_bwd_f32(const float* dy, const float* y, int64_t n, float* du, void* stream) {
  MTASR_CHECK_ARG(dy && y && du && n > 0, "relu_bwd_f32: bad arguments");
  const long long blocks = (n + 255) / 256;
  relu_bwd_f32_kernel<<<static_cast<unsigned>(blocks < 4096 ? blocks : 4096), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, y, n, du);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("relu_bwd_f32");
  return MTASR_OK;
}

extern "C" int mtasr_softmax_scale_f32(const float* logits, const float* lse, const float* rowscale, int64_t rows, int32_t V,
                                       int64_t ld, float* out, void* stream) {
  MTASR_CHECK_ARG(logits && lse && rowscale && out && rows > 0 && V > 0 && ld >= V, "softmax_scale_f32: bad arguments");
  MTASR_CHECK_ARG(rows < (1LL << 31), "softmax_scale_f32: too many rows");
  softmax_scale_f32_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, lse, rowscale, V, ld, out);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("softmax_scale_f32");
  return MTASR_OK;
}
