// First feature-extractor layer (hf:709-751 layer 0): conv1d(1 -> C0, k, stride) on the raw waveform.
// K = k (10) is far too small for tensor cores: this layer is bound by writing its (B, L0, C0) output, so it is a
// direct convolution on CUDA cores with the norm and GELU fused ("layer" norm variant, WavLM-Large) and
// channels-last bf16 output written as 128-byte warp rows.  The "group" variant (WavLM-Base+: GroupNorm with one
// group per channel, i.e. statistics over TIME) needs a full-length reduction first: raw fp32 output here, then
// groupnorm_stats + groupnorm_gelu below.
#include "common.cuh"

namespace mtasr {

static constexpr int C0_MAXI = 8;  // C0 <= 64 * 8 = 512
static constexpr int C0_MAXK = 16;

// mode 1: LayerNorm over channels + GELU -> bf16;  mode 0: raw (+bias) -> fp32
__global__ void __launch_bounds__(128)
conv0_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
             const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int B, int S, int L0, int C0,
             int k, int stride, int mode, __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32) {
  pdl_trigger();
  extern __shared__ float wsm[];  // [k][C0] transposed weights, then bias/gamma/beta [3][C0]
  float* bsm = wsm + k * C0;
  float* gsm = bsm + C0;
  float* besm = gsm + C0;
  for (int i = threadIdx.x; i < k * C0; i += blockDim.x) {
    const int c = i / k, j = i % k;
    wsm[j * C0 + c] = w[i];  // torch layout (C0, 1, k)
  }
  for (int i = threadIdx.x; i < C0; i += blockDim.x) {
    bsm[i] = bias ? bias[i] : 0.f;
    gsm[i] = gamma ? gamma[i] : 1.f;
    besm[i] = beta ? beta[i] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int ni = C0 / 64;
  const long long frames = static_cast<long long>(B) * L0;
  const long long warp0 = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long f0 = warp0 * 2; f0 < frames; f0 += nwarps * 2) {
    float acc[2][C0_MAXI][2];
    const float* xp[2];
    bool ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long f = f0 + r;
      ok[r] = f < frames;
      const long long ff = ok[r] ? f : f0;
      const int b = static_cast<int>(ff / L0), t = static_cast<int>(ff % L0);
      xp[r] = x + static_cast<long long>(b) * S + static_cast<long long>(t) * stride;
#pragma unroll
      for (int i = 0; i < C0_MAXI; ++i) {
        acc[r][i][0] = i < ni ? bsm[64 * i + 2 * lane] : 0.f;
        acc[r][i][1] = i < ni ? bsm[64 * i + 2 * lane + 1] : 0.f;
      }
    }
    for (int j = 0; j < k; ++j) {
      const float x0 = xp[0][j], x1 = xp[1][j];
#pragma unroll
      for (int i = 0; i < C0_MAXI; ++i) {
        if (i < ni) {
          const float2 wv = *reinterpret_cast<const float2*>(wsm + j * C0 + 64 * i + 2 * lane);
          acc[0][i][0] += wv.x * x0; acc[0][i][1] += wv.y * x0;
          acc[1][i][0] += wv.x * x1; acc[1][i][1] += wv.y * x1;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (!ok[r]) continue;  // warp-uniform
      const long long f = f0 + r;
      if (mode == 1) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < C0_MAXI; ++i) if (i < ni) s += acc[r][i][0] + acc[r][i][1];
        const float mean = warp_sum(s) / C0;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < C0_MAXI; ++i)
          if (i < ni) {
            const float d0 = acc[r][i][0] - mean, d1 = acc[r][i][1] - mean;
            q += d0 * d0 + d1 * d1;
          }
        const float rstd = rsqrtf(warp_sum(q) / C0 + eps);
#pragma unroll
        for (int i = 0; i < C0_MAXI; ++i)
          if (i < ni) {
            const int c = 64 * i + 2 * lane;
            const float v0 = gelu_fast_f((acc[r][i][0] - mean) * rstd * gsm[c] + besm[c]);
            const float v1 = gelu_fast_f((acc[r][i][1] - mean) * rstd * gsm[c + 1] + besm[c + 1]);
            *reinterpret_cast<uint32_t*>(y_bf16 + f * C0 + c) = pack_bf16x2(v0, v1);
          }
      } else {
#pragma unroll
        for (int i = 0; i < C0_MAXI; ++i)
          if (i < ni) {
            const int c = 64 * i + 2 * lane;
            *reinterpret_cast<float2*>(y_f32 + f * C0 + c) = make_float2(acc[r][i][0], acc[r][i][1]);
          }
      }
    }
  }
}

// Per-(b, channel) mean / rstd over time of a channels-last (B, L, C) fp32 tensor (GroupNorm with C groups, hf:736-751).
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const float* __restrict__ x, int L, int C, float eps, float* __restrict__ mean,
                       float* __restrict__ rstd) {
  __shared__ double s1[8][33], s2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx, b = blockIdx.y;
  double a = 0.0, q = 0.0;
  if (c < C) {
    const float* xb = x + static_cast<long long>(b) * L * C + c;
    for (int t = ry; t < L; t += 8) {
      const double v = xb[static_cast<long long>(t) * C];
      a += v;
      q += v * v;
    }
  }
  s1[ry][cx] = a;
  s2[ry][cx] = q;
  __syncthreads();
  if (ry == 0 && c < C) {
    double ta = 0.0, tq = 0.0;
    for (int i = 0; i < 8; ++i) { ta += s1[i][cx]; tq += s2[i][cx]; }
    const double m = ta / L;
    const double var = tq / L - m * m;
    mean[b * C + c] = static_cast<float>(m);
    rstd[b * C + c] = static_cast<float>(1.0 / sqrt((var > 0 ? var : 0) + static_cast<double>(eps)));
  }
}

__global__ void groupnorm_gelu_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int B, int L, int C,
                                      __nv_bfloat16* __restrict__ y, float* __restrict__ y32) {
  const long long n2 = static_cast<long long>(B) * L * (C / 2);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (C / 2)) * 2;
    const long long bl = i / (C / 2);
    const int b = static_cast<int>(bl / L);
    const float2 v = *reinterpret_cast<const float2*>(x + bl * C + c);
    const float o0 = gelu_fast_f((v.x - mean[b * C + c]) * rstd[b * C + c] * gamma[c] + beta[c]);
    const float o1 = gelu_fast_f((v.y - mean[b * C + c + 1]) * rstd[b * C + c + 1] * gamma[c + 1] + beta[c + 1]);
    if (y) *reinterpret_cast<uint32_t*>(y + bl * C + c) = pack_bf16x2(o0, o1);
    if (y32) *reinterpret_cast<float2*>(y32 + bl * C + c) = make_float2(o0, o1);
  }
}

}  // namespace mtasr

using namespace mtasr;

extern "C" int mtasr_conv0_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                               float eps, int32_t B, int32_t S, int32_t C0, int32_t k, int32_t stride, int32_t mode,
                               void* y_bf16, float* y_f32, void* stream) {
  MTASR_CHECK_ARG(x && w && B > 0 && S >= k && k > 0 && k <= C0_MAXK && stride > 0, "conv0_fwd: bad arguments");
  MTASR_CHECK_ARG(C0 % 64 == 0 && C0 <= 64 * C0_MAXI, "conv0_fwd: C0=%d must be a multiple of 64 and <= 512", C0);
  MTASR_CHECK_ARG(mode == 1 ? (y_bf16 && gamma && beta) : (y_f32 != nullptr), "conv0_fwd: missing output / norm params");
  const int L0 = (S - k) / stride + 1;
  const size_t smem = sizeof(float) * (static_cast<size_t>(k) * C0 + 3 * C0);
  const long long frames = static_cast<long long>(B) * L0;
  long long grid = (frames + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (grid > cap) grid = cap;
  conv0_kernel<<<static_cast<unsigned>(grid), 128, smem, static_cast<cudaStream_t>(stream)>>>(
      x, w, bias, gamma, beta, eps, B, S, L0, C0, k, stride, mode, reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("conv0_fwd");
  return MTASR_OK;
}

static int groupnorm_gelu_launch(const float* x, const float* gamma, const float* beta, float eps, int32_t B, int32_t L,
                                 int32_t C, float* mean_ws, float* rstd_ws, void* y_bf16, float* y_f32, void* stream) {
  MTASR_CHECK_ARG(x && gamma && beta && mean_ws && rstd_ws && (y_bf16 || y_f32) && B > 0 && L > 0 && C > 0 && C % 2 == 0,
                  "groupnorm_gelu: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  groupnorm_stats_kernel<<<dim3((C + 31) / 32, B), 256, 0, st>>>(x, L, C, eps, mean_ws, rstd_ws);
  MTASR_COUNT_LAUNCH();
  const long long n2 = static_cast<long long>(B) * L * (C / 2);
  long long grid = (n2 + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (grid > cap) grid = cap;
  groupnorm_gelu_kernel<<<static_cast<unsigned>(grid), 256, 0, st>>>(x, mean_ws, rstd_ws, gamma, beta, B, L, C,
                                                                     reinterpret_cast<__nv_bfloat16*>(y_bf16), y_f32);
  MTASR_COUNT_LAUNCH();
  MTASR_CHECK_LAUNCH("groupnorm_gelu");
  return MTASR_OK;
}

extern "C" int mtasr_groupnorm_gelu(const float* x, const float* gamma, const float* beta, float eps, int32_t B,
                                    int32_t L, int32_t C, float* mean_ws, float* rstd_ws, void* y_bf16, void* stream) {
  MTASR_CHECK_ARG(y_bf16 != nullptr, "groupnorm_gelu: bad arguments");
  return groupnorm_gelu_launch(x, gamma, beta, eps, B, L, C, mean_ws, rstd_ws, y_bf16, nullptr, stream);
}

extern "C" int mtasr_groupnorm_gelu_f32(const float* x, const float* gamma, const float* beta, float eps, int32_t B,
                                        int32_t L, int32_t C, float* mean_ws, float* rstd_ws, float* y_f32, void* stream) {
  MTASR_CHECK_ARG(y_f32 != nullptr, "groupnorm_gelu_f32: bad arguments");
  return groupnorm_gelu_launch(x, gamma, beta, eps, B, L, C, mean_ws, rstd_ws, nullptr, y_f32, stream);
}
