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
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace mtasr {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int STAGES = 4;
static constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
static constexpr int B_STAGE_BYTES = 256 * BK * 2;  // 32 KB (block_n <= 256)
static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
static constexpr int GEMM_THREADS = 256;
static constexpr int GEMM_SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static constexpr int TMEM_COLS = 512;

struct GemmKP {
  int M, N, K, batch0, batch1;
  int a_major, b_major, block_n;
  int a_inner, a_phase;
  int a_use0, a_use1, b_use0, b_use1;  // 0 => operand broadcast over that batch dim (coordinate forced to 0)
  int m_tiles, n_tiles, num_tiles, num_kb;
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

struct TileCoord {
  int m_tile, n_tile, b0, b1;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmKP& p, int tile) {
  TileCoord t;
  t.m_tile = tile % p.m_tiles;
  int rest = tile / p.m_tiles;
  t.n_tile = rest % p.n_tiles;
  int batch = rest / p.n_tiles;
  t.b0 = batch % p.batch0;
  t.b1 = batch / p.batch0;
  return t;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmKP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tmap(&tmap_a);
      prefetch_tmap(&tmap_b);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
      }
      fence_barrier_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = A_STAGE_BYTES + p.block_n * BK * 2;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile);
        const int ab0 = p.a_use0 ? t.b0 : 0, ab1 = p.a_use1 ? t.b1 : 0;
        const int bb0 = p.b_use0 ? t.b0 : 0, bb1 = p.b_use1 ? t.b1 : 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          uint64_t* bar = &full_bar[stage];
          mbar_arrive_expect_tx(bar, stage_tx);
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
            tma_load_4d(sb, &tmap_b, bar, kb * BK, t.n_tile * p.block_n, bb0, bb1);
          } else {
            for (int j = 0; j < p.block_n / 64; ++j)
              tma_load_4d(sb + j * 8192, &tmap_b, bar, t.n_tile * p.block_n + j * 64, kb * BK, bb0, bb1);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc_bf16(BM, p.block_n, p.a_major, p.b_major);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major SW128: +32 B per UMMA_K inside the 128 B swizzle row; 8-row groups 1024 B apart (SBO).
            // MN-major SW128: 64(mn) x 8(k) atoms of 1024 B; next 8 k = +1024 B (SBO), next 64 mn = +8192 B (LBO).
            const uint64_t da = p.a_major == 0 ? make_smem_desc(sa + k * 32, 16, 1024)
                                               : make_smem_desc(sa + k * 2048, 8192, 1024);
            const uint64_t db = p.b_major == 0 ? make_smem_desc(sb + k * 32, 16, 1024)
                                               : make_smem_desc(sb + k * 2048, 8192, 1024);
            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const long long batch = static_cast<long long>(t.b1) * p.batch0 + t.b0;
      const int row = t.m_tile * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const int col_base = t.n_tile * p.block_n;
      const int ncols = min(p.block_n, p.N - col_base);
      const long long c_off = t.b0 * p.c_sb0 + t.b1 * p.c_sb1 + static_cast<long long>(row) * p.c_ld;
      const long long r_off = t.b0 * p.r_sb0 + t.b1 * p.r_sb1 + static_cast<long long>(row) * p.r_ld;
      const float* bias = p.bias ? p.bias + t.b0 * p.bias_sb0 : nullptr;
      float rvec = 0.f, rscale = 1.f;
      if (p.mode == 2 && row_ok) {
        rvec = p.row_vec[batch * p.M + row];
        rscale = p.row_scale ? p.row_scale[batch * p.M + row] : 1.f;
      }
      float run_max = -INFINITY, run_sum = 0.f;
      int run_idx = 0;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;
      for (int chunk = 0; chunk * 32 < ncols; ++chunk) {
        uint32_t r[32];
        tmem_ld32(taddr + chunk * 32, r);
        tmem_ld_wait();
        const int c0 = col_base + chunk * 32;
        if (p.mode == 1) {
          float v[32];
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = c0 + j;
            float x = p.alpha * __uint_as_float(r[j]);
            if (col < p.N) {
              if (bias) x += __ldg(bias + col);
            } else {
              x = -INFINITY;
            }
            v[j] = x;
            cm = fmaxf(cm, x);
          }
          if (cm > run_max) {
            run_sum *= __expf(run_max - cm);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (v[j] == cm) { run_idx = c0 + j; break; }
            run_max = cm;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) run_sum += __expf(v[j] - run_max);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c0 + g * 8;
            const int nv = min(8, p.N - col);
            if (nv <= 0) break;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = p.alpha * __uint_as_float(r[g * 8 + j]);
              if (bias && j < nv) x += __ldg(bias + col + j);
              v[j] = x;
            }
            if (p.mode == 2) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __expf(v[j] - rvec) * rscale;
            } else {
              if (p.aux && row_ok) store8(p.aux, MTASR_DT_BF16, c_off + col, nv, v);
              if (p.act == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_f(v[j]);
              } else if (p.act == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              if (p.residual && row_ok) {
                float rr[8];
                load8(p.residual, p.res_dtype, r_off + col, nv, rr);
                if (p.act == 3) {        // backward through GELU: residual holds the saved pre-activation
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] *= gelu_grad_f(rr[j]);
                } else if (p.act == 4) { // backward through ReLU: residual holds the saved activation output
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = rr[j] > 0.f ? v[j] : 0.f;
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] += rr[j];
                }
              }
              if (p.accumulate && row_ok) {
                float cc[8];
                load8(p.c, p.c_dtype, c_off + col, nv, cc);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += cc[j];
              }
            }
            if (row_ok) store8(p.c, p.c_dtype, c_off + col, nv, v);
          }
        }
      }
      if (p.mode == 1 && row_ok) {
        p.lse_part[(batch * p.M + row) * p.n_tiles + t.n_tile] =
            make_float4(run_max, run_sum, __int_as_float(run_idx), 0.f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

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
                      const uint32_t* box, const char* what) {
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
    gstr[i - 1] = strides_elems[i] * 2;
    if (gstr[i - 1] % 16 != 0 || gstr[i - 1] == 0)
      return set_error(MTASR_ERR_INVALID_ARG, "gemm: %s stride[%d]=%llu bytes is not a positive multiple of 16", what, i,
                       (unsigned long long)gstr[i - 1]);
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0)
    return set_error(MTASR_ERR_INVALID_ARG, "gemm: %s base pointer not 16-byte aligned", what);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(MTASR_ERR_DRIVER,
                     "gemm: cuTensorMapEncodeTiled(%s) failed with %d (rank %d dims %llu,%llu,%llu,%llu strides %llu,%llu,%llu box %u,%u,%u)",
                     what, (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1],
                     (unsigned long long)gdim[2], (unsigned long long)(rank > 3 ? gdim[3] : 0),
                     (unsigned long long)gstr[0], (unsigned long long)gstr[1], (unsigned long long)(rank > 3 ? gstr[2] : 0),
                     bx[0], bx[1], bx[2]);
  }
  return 0;
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

extern "C" int mtasr_gemm_n_tiles(int32_t N, int32_t block_n) {
  if (block_n != 64 && block_n != 128 && block_n != 256) block_n = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  return (N + block_n - 1) / block_n;
}

extern "C" int mtasr_gemm_bf16(const mtasr_gemm_desc* d, void* stream) {
  MTASR_CHECK_ARG(d != nullptr, "gemm: null descriptor");
  MTASR_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "gemm: M,N,K must be positive (got %d,%d,%d)", d->M, d->N, d->K);
  MTASR_CHECK_ARG(d->batch0 > 0 && d->batch1 > 0, "gemm: batch dims must be positive");
  MTASR_CHECK_ARG(d->a && d->b, "gemm: null operand");
  MTASR_CHECK_ARG(d->mode >= 0 && d->mode <= 2, "gemm: bad mode %d", d->mode);
  MTASR_CHECK_ARG(d->mode == 1 ? d->lse_part != nullptr : d->c != nullptr, "gemm: missing output buffer");
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
  p.m_tiles = (d->M + BM - 1) / BM;
  p.n_tiles = (d->N + bn - 1) / bn;
  const long long nt = static_cast<long long>(p.m_tiles) * p.n_tiles * d->batch0 * d->batch1;
  MTASR_CHECK_ARG(nt < (1LL << 31), "gemm: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.num_kb = (d->K + BK - 1) / BK;
  p.c = d->c; p.c_dtype = d->c_dtype; p.c_ld = d->c_ld; p.c_sb0 = d->c_sb0; p.c_sb1 = d->c_sb1;
  p.aux = reinterpret_cast<__nv_bfloat16*>(d->aux);
  p.bias = d->bias; p.bias_sb0 = d->bias_sb0;
  p.residual = d->residual; p.res_dtype = d->res_dtype; p.r_ld = d->r_ld; p.r_sb0 = d->r_sb0; p.r_sb1 = d->r_sb1;
  p.act = d->act; p.alpha = d->alpha; p.accumulate = d->accumulate; p.mode = d->mode;
  p.row_vec = d->row_vec; p.row_scale = d->row_scale;
  p.lse_part = reinterpret_cast<float4*>(d->lse_part);

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
    const uint32_t box[4] = {BK, static_cast<uint32_t>(bn), 1, 1};
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

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  });
  if (attr_err != cudaSuccess)
    return set_error(MTASR_ERR_LAUNCH, "gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));

  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
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
  gemm_bf16_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(ma, mb, p);
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
