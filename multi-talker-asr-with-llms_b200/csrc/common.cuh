// Shared helpers for the mtasr C-ABI library (sm_100a only).
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/mtasr.h"

namespace mtasr {

// Thread-local last-error string, returned through mtasr_last_error_string().
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define MTASR_CHECK_ARG(cond, ...)                                  \
  do {                                                              \
    if (!(cond)) return ::mtasr::set_error(MTASR_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define MTASR_CHECK_LAUNCH(name)                                                                   \
  do {                                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return ::mtasr::set_error(MTASR_ERR_LAUNCH, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

int num_sms();
int num_units(int ncta);   // CTAs or TPC-safe CTA pairs available to a persistent kernel (api.cu)
extern std::atomic<long long> g_launches;
#define MTASR_COUNT_LAUNCH() ::mtasr::g_launches.fetch_add(1)

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------
// Kernel-to-kernel latency inside the step's CUDA graph is ~1 us and PDL does not shorten it (tools/micro/pdl_gap.cu, B200:
// 1.10 -> 0.96 us per empty launch); what PDL hides is the PROLOGUE of the dependent kernel.  The tcgen05 kernels (GEMM,
// fused attention: 386 launches per cfg2 step) spend 1.5-2 us per launch on tensor-map prefetch, mbarrier init, the TMEM
// allocation and a cluster / CTA barrier before they touch global memory; launched with
// cudaLaunchAttributeProgrammaticStreamSerialization they run that part while the predecessor drains and block in
// `griddepcontrol.wait` (pdl_wait) until the predecessor grid has completed and flushed.  Every kernel of the library
// issues `griddepcontrol.launch_dependents` (pdl_trigger) as its first instruction so that a PDL successor may be
// scheduled as early as its resources allow; that is harmless without such a successor, and correctness never depends on
// it: only kernels whose prologue ends in pdl_wait() are launched with the attribute.  MTASR_PDL=0 switches it off.
bool pdl_enabled();   // api.cu
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Launch of a kernel whose prologue ends in pdl_wait(): with the programmatic-serialization attribute when PDL is on.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// erf-based GELU (HF "gelu" == torch.nn.functional.gelu default) and its derivative.
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// Raw SFU approximations (single MUFU, no range fix-up branches: the fix-ups of exp2f()/__frcp_rn() put a
// BSSY/BSYNC region around every element and serialise the otherwise independent per-element chains).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Epilogue-speed GELU: erf by the Abramowitz-Stegun 7.1.26 rational form (|abs err| < 1.5e-7 plus ~1e-7 from the
// approximate SFU ops, i.e. at fp32 rounding level of the GELU output) with one MUFU.RCP + one MUFU.EX2.
__device__ __forceinline__ float erf_fast_pos(float ax, float e /* = exp(-ax*ax) */) {
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  return fmaf(-poly * t, e, 1.0f);
}
__device__ __forceinline__ float gelu_fast_f(float x) {
  const float ax = fabsf(x);
  const float z = ax * 0.70710678118654752f;
  const float e = ex2_approx(z * z * -1.4426950408889634f);
  return 0.5f * fmaf(ax, erf_fast_pos(z, e), x);             // 0.5 x (1 + sign(x) erf|z|)
}
__device__ __forceinline__ float gelu_grad_fast_f(float x) {
  const float ax = fabsf(x);
  const float z = ax * 0.70710678118654752f;
  const float e = ex2_approx(z * z * -1.4426950408889634f);      // exp(-x^2/2)
  const float er = copysignf(erf_fast_pos(z, e), x);
  return fmaf(x * 0.39894228040143268f, e, fmaf(0.5f, er, 0.5f));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---- dropout: counter-based, stateless RNG (hf:291,294,323,364,483 hidden / activation dropout; hf:217 attention dropout).
// The keep decision of element `idx` of dropout site `site` is a pure function of (seed[0], seed[1], site, idx): the forward,
// the backward and a gradient-checkpoint replay regenerate identical masks without storing them.  `seed` is two 32-bit words
// in DEVICE memory drawn by a torch op from torch's CUDA generator (no host sync; torch.manual_seed reproduces a run).
// One 32-bit hash (multiplicative mix + lowbias32 finaliser) yields the decisions of TWO neighbouring elements (16 bits
// each): keep iff u16 < thresh, thresh = round((1 - p) * 65536); kept values are scaled by 65536 / thresh (exactly unbiased
// for the quantised keep probability).
struct DropP {
  const uint32_t* seed;   // nullptr = no dropout
  uint32_t site;
  uint32_t thresh;
  float scale;
  long long ld;           // row pitch of the index space (even): idx = row * ld + col
};
__device__ __forceinline__ uint32_t drop_hash(uint32_t s0, uint32_t s1, uint32_t site, unsigned long long pair) {
  const uint32_t lo = static_cast<uint32_t>(pair), hi = static_cast<uint32_t>(pair >> 32);
  uint32_t h = (lo * 0x9E3779B1u) ^ s0;
  h ^= ((hi + site * 0x85EBCA77u) * 0xC2B2AE3Du) + s1;
  h ^= h >> 16; h *= 0x7feb352du;
  h ^= h >> 15; h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}
// multiplier (0 or scale) of element idx
__device__ __forceinline__ float drop_mult(uint32_t s0, uint32_t s1, const DropP& d, unsigned long long idx) {
  const uint32_t h = drop_hash(s0, s1, d.site, idx >> 1);
  const uint32_t u = (idx & 1ull) ? (h >> 16) : (h & 0xffffu);
  return u < d.thresh ? d.scale : 0.f;
}
// multipliers of the 8 consecutive elements idx .. idx+7, idx even: 4 hashes
__device__ __forceinline__ void drop_mult8(uint32_t s0, uint32_t s1, const DropP& d, unsigned long long idx, float (&m)[8]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = drop_hash(s0, s1, d.site, (idx >> 1) + j);
    m[2 * j] = (h & 0xffffu) < d.thresh ? d.scale : 0.f;
    m[2 * j + 1] = (h >> 16) < d.thresh ? d.scale : 0.f;
  }
}

// log(exp(a)+exp(b)) that tolerates -inf on either side.
__device__ __forceinline__ float logaddexp_f(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(fminf(a, b) - m));
}

// Block-wide sum for blockDim.x <= 1024 (result broadcast to all threads).
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

}  // namespace mtasr
