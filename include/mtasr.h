/* mtasr.h -- C ABI of the B200-native encoder + serialized-CTC hot path (libmtasr.so, sm_100a).
 *
 * The reference (Hao-Shi-SBINT/Multi-talker-ASR-with-LLMs) has no FFI/plugin layer: its hot path calls
 * torch/ATen and HF `transformers` directly.  The boundary a maintainer binds is therefore this
 * header, loaded with ctypes from the Python classes that mirror the reference's own
 * (WavLMModel / Separator / CTC / HybridLoss).  Each entry point names the reference call site it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no sync, no allocation);
 *   - the caller owns every buffer including workspaces;
 *   - return 0 on success, negative MTASR_ERR_* otherwise; mtasr_last_error_string() gives the reason
 *     (thread-local); nothing throws across the boundary;
 *   - bf16 = __nv_bfloat16 bit pattern (uint16_t), f32 = float, i64 = int64_t, i32 = int32_t.
 */
#ifndef MTASR_H_
#define MTASR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTASR_OK 0
#define MTASR_ERR_INVALID_ARG (-1)
#define MTASR_ERR_UNSUPPORTED (-2)
#define MTASR_ERR_LAUNCH (-3)
#define MTASR_ERR_DRIVER (-4)

#define MTASR_DT_BF16 0
#define MTASR_DT_F32 1
#define MTASR_DT_F16 2 /* only as the optional logits output of mtasr_gemm_bf16 mode 1 and the input of mtasr_softmax_from_logits */

int mtasr_version(void);
/* Cap the number of SMs the persistent kernels (GEMM, attention) size their grids for; 0 = all.  Returns the effective count.
 * Data-parallel training leaves a few SMs to the NCCL gradient all-reduce during the backward: the persistent kernels
 * otherwise occupy every SM back to back and the collective only runs once the backward has drained. */
int mtasr_set_sm_budget(int32_t n_sms);
const char* mtasr_last_error_string(void);
/* Number of kernels this library has enqueued since load (process-wide; used by bench.py `gpu_launches`). */
int64_t mtasr_launch_count(void);

/* Per-launch timing of the tcgen05 GEMM kernel (bench.py `roofline`): between begin and end every
 * mtasr_gemm_bf16 launch is bracketed by CUDA events on its own stream; end synchronises and returns the summed
 * kernel milliseconds, the executed flops (2*M*N*K*batches, recompute included) and the launch count. */
int mtasr_profile_begin(void);
int mtasr_profile_end(double* gemm_ms, double* gemm_flops, int64_t* gemm_launches);

/* ------------------------------------------------------------------------------------------------------------
 * Dense contraction on tcgen05 tensor cores (TMA-fed, TMEM accumulators, fp32 accumulate, bf16 operands).
 *   C[b][m][n] = epilogue( alpha * sum_k A[b][m][k] * B[b][n][k] )
 * Replaces every F.linear / F.conv1d / bmm on the path: hf:100-105 (projection), hf:82-90 (pos-conv as implicit
 * GEMM), hf:206-228 (q/k/v/out projections and QK^T / PV), hf:288-295 (FFN), hf:709-727 (conv layers 1-6 as
 * implicit GEMM over a strided channels-last view), hf:803-807 (adapter convs), ref:models/separator.py:158-165,
 * ref:models/ctc.py:139,180,190 (ctc_lo) and all of their backward contractions.
 *
 * Operand addressing (element units, bf16):
 *  A, K-major (a_major=0):   A(b,m,k) = a[b0*a_sb0 + b1*a_sb1 + (m + tap/a_phase)*a_ld + (tap%a_phase)*a_inner + c]
 *                            with k = tap*a_inner + c, c < a_inner.  Plain GEMM: a_inner=K, a_phase=1.
 *                            Implicit conv1d on channels-last x[b][l][c]: a_inner=C_in, a_phase=stride,
 *                            a_ld=stride*C_in, K=kernel*C_in (a_inner % 64 == 0 required when K > a_inner).
 *  A, MN-major (a_major=1):  A(b,m,k) = a[b0*a_sb0 + b1*a_sb1 + k*a_ld + m]
 *  B, K-major (b_major=0):   B(b,n,k) = bm[b0*b_sb0 + b1*b_sb1 + n*b_ld + k]
 *  B, MN-major (b_major=1):  B(b,n,k) = bm[b0*b_sb0 + b1*b_sb1 + k*b_ld + n]
 *  b = b1*batch0 + b0.  A batch stride of 0 broadcasts the operand over that batch dimension.
 *  a_rows / b_rows: number of addressable rows per batch (TMA zero-fills beyond them; defaults: see gemm.cu).
 *  Alignment: base pointers and all strides (bytes) multiples of 16.
 *
 * Epilogue (per element, in this order): v = alpha*acc; v += bias[b0*bias_sb0 + n]; aux[..] = v (optional, bf16,
 * the pre-activation); v = act(v); v += residual[..]; if accumulate: v += C_old; C = v.
 *  act 3 / act 4 are the backward forms: v *= gelu'(residual) / v *= (residual > 0), with `residual` holding the
 *  saved pre-activation (GELU) or activation output (ReLU) instead of being added.
 *  mode 1 (LSE partials, ref:models/ctc.py:53 log_softmax fused): for every row and N-tile writes
 *    {max, sum exp(v-max), argmax index} to lse_part[(b*M+m)*n_tiles + n_tile] (float4, .w unused).  C is optional:
 *    NULL = no tensor output (inference); otherwise c_dtype must be MTASR_DT_F16 and the logits tile v is also written
 *    as fp16 (the training forward keeps it so that the backward turns it into softmax * upstream with one streaming
 *    pass, mtasr_softmax_from_logits, instead of repeating the vocabulary GEMM).
 *  mode 2 (softmax regeneration for the backward): C = exp(v - row_vec[b*M+m]) * row_scale[b*M+m].
 */
typedef struct mtasr_gemm_desc {
  int32_t M, N, K;
  int32_t batch0, batch1;
  int32_t a_major, b_major;
  int32_t block_n; /* 64, 128 or 256; 0 = choose */
  const void* a;
  int64_t a_ld, a_sb0, a_sb1;
  int32_t a_inner, a_phase;
  int64_t a_rows;
  const void* b;
  int64_t b_ld, b_sb0, b_sb1;
  int64_t b_rows;
  void* c;
  int32_t c_dtype;
  int64_t c_ld, c_sb0, c_sb1;
  void* aux; /* bf16, same layout as c, optional */
  const float* bias;
  int64_t bias_sb0;
  const void* residual;
  int32_t res_dtype;
  int64_t r_ld, r_sb0, r_sb1;
  int32_t act; /* 0 none, 1 gelu(erf), 2 relu, 3 gelu-backward, 4 relu-backward */
  float alpha;
  int32_t accumulate;
  int32_t mode;
  const float* row_vec;
  const float* row_scale;
  float* lse_part;
  /* fused dropout (mode 0 only; hf:291,294,323,364): drop_seed = two u32 words in device memory (NULL = off), element index
   * (b*M + m) * drop_ld + n with drop_ld = N rounded up to even, keep probability drop_keep16 / 65536.  Applied to v after
   * bias / activation and BEFORE the residual add (forward sites); with act 3 / 4 the same mask multiplies the gradient
   * (backward of an activation-dropout site: v *= mask * gelu'(residual)).  See mtasr_dropout for the generator. */
  const void* drop_seed;
  uint32_t drop_site;
  uint32_t drop_keep16;
} mtasr_gemm_desc;

int mtasr_gemm_bf16(const mtasr_gemm_desc* desc, void* stream);
/* Number of N tiles mtasr_gemm_bf16 will use for (N, block_n) -- sizes the mode-1 partials buffer. */
int mtasr_gemm_n_tiles(int32_t N, int32_t block_n);

/* ------------------------------------------------------------------------------------------------------------
 * Serialized CTC (one head at a time; ref:models/ctc.py:44-65, torch.nn.CTCLoss(reduction='none',
 * zero_infinity=True, blank=V-1) semantics).  "Compact lattice columns": glog (B,T,Lp) f32 holds logits of column
 * 0 = blank and 1+l = label l of utterance b; lse (B,T) is the row log-sum-exp over the whole vocabulary.
 * ys (B, ys_ld) i64 padded labels, hlens/ylens (B) i64.  Lp >= max_label_len + 1, max_label_len <= 255.
 */
/* padded state count SP = 32*NS used to size alpha_ws (B*T*SP f32); -1 if max_label_len unsupported */
int mtasr_ctc_state_pad(int32_t max_label_len);
/* alpha recursion: nll_out (B) = per-utterance loss with infeasible rows zeroed, nll_raw (B) f64 keeps +inf. */
int mtasr_ctc_alpha_fwd(const float* glog, const float* lse, const int64_t* ys, const int64_t* hlens,
                        const int64_t* ylens, int32_t B, int32_t T, int32_t Lp, int32_t ys_ld, int32_t max_label_len,
                        float* alpha_ws, double* coff_ws, float* nll_out, double* nll_raw, void* stream);
/* beta recursion + gradient: dG (B,T,Lp) = gout[b] * d nll_b / d lp(t, column) (= -gout*occupancy),
 * rowscale (B,T) = gout[b] on valid frames of feasible utterances else 0 (scales the dense softmax term).
 * dG must arrive ZEROED: only the (t < hlens[b], column <= ylens[b]) entries of feasible utterances are written.
 * Both recursions stream their per-frame inputs through a double-buffered shared-memory ring (cp.async), one warp
 * (= one CTA) per utterance; glog / alpha_ws must be 16-byte aligned and Lp a multiple of 4. */
int mtasr_ctc_beta_bwd(const float* glog, const float* lse, const int64_t* ys, const int64_t* hlens,
                       const int64_t* ylens, int32_t B, int32_t T, int32_t Lp, int32_t ys_ld, int32_t max_label_len,
                       const float* alpha_ws, const double* coff_ws, const double* nll_raw, const float* gout,
                       float* dG, float* rowscale, void* stream);
/* Combine the mode-1 GEMM partials: lse (rows) f32 and/or argmax (rows) i64 (first maximal index, like
 * torch.argmax -- ref:models/ctc.py:190). */
int mtasr_lse_finalize(const float* part, int64_t rows, int32_t n_tiles, float* lse, int64_t* argmax, void* stream);
/* Greedy collapse of ref:models/modeling_speech_encoder_decoder_llama.py:902-972 on (B,T) i64 argmax ids:
 * out (B,T) i64 right-padded with pad_id, lengths (B) i32. */
int mtasr_ctc_collapse(const int64_t* ids, int32_t B, int32_t T, int64_t blank_id, int64_t pad_id, int64_t* out,
                       int32_t* lengths, void* stream);
/* Token segments of ref:models/mt_ctctoken_builder.py:56-157 on a greedy path (B,T) i64 with frame mask (B,T) u8 (1 = valid):
 * seg_start / seg_end (B,T) i32 hold the first / last frame of every emitted segment, nseg (B) i32 their count. */
int mtasr_ctc_segments(const int64_t* path, const uint8_t* mask, int32_t B, int32_t T, int64_t blank, int32_t* seg_start,
                       int32_t* seg_end, int32_t* nseg, void* stream);
/* out (B,Lmax,D) f32 = per-segment mean of x (B,T,D) f32 (0 beyond nseg); conf (B,Lmax) = clamp(1 - mean pblank, 0, 1) or NULL.
 * Backward: dx (B,T,D) f32 (zero it first) receives dout / segment length on the frames of each segment. */
int mtasr_segment_mean_fwd(const float* x, const float* pblank, const int32_t* seg_start, const int32_t* seg_end,
                           const int32_t* nseg, int32_t B, int32_t T, int32_t D, int32_t Lmax, float* out, float* conf, void* stream);
int mtasr_segment_mean_bwd(const float* dout, const int32_t* seg_start, const int32_t* seg_end, const int32_t* nseg, int32_t B,
                           int32_t T, int32_t D, int32_t Lmax, float* dx, void* stream);
/* dense (B,T,V) f32 <-> compact lattice columns (B,T,Lp) f32 (scatter ADDS into dense). */
int mtasr_ctc_gather_cols(const float* dense, const int64_t* ys, const int64_t* ylens, int32_t B, int32_t T, int32_t V,
                          int32_t Lp, int32_t ys_ld, int64_t blank, float* out, void* stream);
int mtasr_ctc_scatter_cols(const float* src, const int64_t* ys, const int64_t* ylens, int32_t B, int32_t T, int32_t V,
                           int32_t Lp, int32_t ys_ld, int64_t blank, float* dense, void* stream);
/* rows of the head weight (V,D) bf16 / bias (V) f32 touched by each lattice -> wg (B,Lp,D) bf16, bg (B,Lp) f32;
 * and the reverse scatter-add of their gradients (f32 atomics).  Label ids outside [0, V) (a -100 pad counted into ylens)
 * gather a zero row / scatter nothing, and ylens is clamped to the label row: no access leaves the (V, D) matrix. */
int mtasr_ctc_gather_rows(const void* w_bf16, const float* bias, const int64_t* ys, const int64_t* ylens, int32_t B,
                          int32_t Lp, int32_t D, int32_t ys_ld, int64_t blank, int64_t V, void* wg_bf16, float* bg,
                          void* stream);
int mtasr_ctc_scatter_rows(const float* dwg, const float* dbg, const int64_t* ys, const int64_t* ylens, int32_t B,
                           int32_t Lp, int32_t D, int32_t ys_ld, int64_t blank, int64_t V, float* dw, float* db,
                           void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * HBM-bound row kernels (128-bit vectorised, one warp per row).
 */
/* LayerNorm over the last dim D (<=1024, %8): hf:100-105, hf:355-366, hf:314-329, hf:720-727 (post_gelu=1),
 * ref:models/separator.py:158-165.  x f32|bf16 -> y_bf16 and/or y_f32; mean/rstd (rows) saved for backward. */
int mtasr_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, float eps, int64_t rows,
                        int32_t D, int32_t post_gelu, void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream);
/* dx (+ dres) as f32 and/or bf16; dgamma/dbeta are ACCUMULATED (atomics) -- zero them first. */
int mtasr_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* mean,
                        const float* rstd, const float* gamma, const float* dres, int64_t rows, int32_t D, float* dx_f32,
                        void* dx_bf16, float* dgamma, float* dbeta, void* stream);
/* The same pass additionally ACCUMULATES dxsum[c] += sum_rows dx[row][c] (zero it first): the bias gradient of the Linear
 * whose output gradient this dx is (out-proj / FFN2 of the neighbouring half encoder layer, hf:355-366), so that no
 * separate column-sum pass re-reads dx.  dxsum needs dx_f32 or dx_bf16; any of dgamma / dbeta / dxsum may be NULL.
 * gate_ab (rows, D/64, 2) f32 + gate_w8 (8, 64) f32 (both or neither): the upstream gradient additionally receives the
 * gru_rel_pos gate's path into the LayerNorm output, dy[r][c] += da[r][c/64] * wa[c%64] + db[r][c/64] * wb[c%64] with
 * wa / wb the 4-row sums of gate_w8 (mtasr_relpos_gate_bwd with dab instead of dx) -- the rank-2-per-head term is never
 * materialised as a (rows, D) tensor. */
int mtasr_layernorm_bwd_sums(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* mean,
                             const float* rstd, const float* gamma, const float* dres, int64_t rows, int32_t D,
                             float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, const float* gate_ab,
                             const float* gate_w8, void* stream);
int mtasr_cast_f32_bf16(const float* x, void* y_bf16, int64_t n, void* stream);
/* Label splitter of ref:utils/split_labels_by_sc.py:21-75 on the device: labels (B, L) i64 with row stride ld -> out (K, B, L) i64
 * (must arrive filled with the pad value), lens (K, B) i64, status (3) i32 = {smallest failing row, INT32_MAX if none; kind:
 * 1 = separator count != K-1, 2 = empty segment while !allow_empty; the count / slot seen} (must arrive as {INT32_MAX,0,0}). */
int mtasr_split_labels(const int64_t* labels, int32_t B, int32_t L, int64_t ld, int32_t K, int64_t sep_id, int64_t pad_id,
                       int32_t has_pad, int64_t ignore_id, int32_t has_ignore, int64_t end_id, int32_t has_end, int32_t allow_empty,
                       int64_t* out, int64_t* lens, int32_t* status, void* stream);
/* PCGrad projection of ref:src/trainer_seq2seq.py:1116-1124 on flat fp32 gradient vectors, no host synchronisation:
 * dots: out2 = {<gi,gj>, <gj,gj>};  project: gi -= (dots2[0] < 0 ? dots2[0] / (dots2[1] + 1e-12) : 0) * gj. */
int mtasr_pcgrad_dots(const float* gi, const float* gj, int64_t n, float* out2, void* stream);
int mtasr_pcgrad_project(float* gi, const float* gj, int64_t n, const float* dots2, void* stream);
/* Dropout as a stand-alone pass: y[r][c] = x[r][c] * mask(seed, site, r * ld_even + c) * 65536 / keep16 for r < rows,
 * c < cols (contiguous rows; ld_even = cols rounded up to even), x f32 or bf16 -> y f32 or bf16.  seed: two u32 words in
 * device memory drawn from torch's CUDA generator; the mask is a pure function of (seed, site, index): forward, backward and
 * checkpoint replays regenerate it.  Same generator as the fused GEMM / attention sites. */
int mtasr_dropout(const void* x, int32_t x_dtype, int64_t rows, int64_t cols, const void* seed, uint32_t site, uint32_t keep16,
                  void* y, int32_t y_dtype, void* stream);
/* fp32-accurate GEMM mode (north_star tolerance "<= 1e-4 in fp32"; the reference's fp32 run is torch fp32 matmul,
 * ref:models/modeling_wavlm.py:412-465 outside autocast).  x (n fp32 values, rows of chunks of c values, c % 8 == 0)
 * -> y (terms*n bf16).  terms 3: every chunk becomes [lo | hi | hi] (order 0, A-side operand) or [hi | lo | hi]
 * (order 1, B-side), hi = bf16(x), lo = bf16(x - hi): one bf16 GEMM over the 3x longer contraction evaluates
 * a_lo b_hi + a_hi b_lo + a_hi b_hi with fp32 accumulation (~2^-17 per product; small products first, because the
 * tensor core's accumulation truncates).  terms 6: three-way split x = x1 + x2 + x3, A side [a3|a2|a1|a2|a1|a1],
 * B side [b1|b2|b3|b1|b2|b1] (all products down to 2^-24). */
int mtasr_split_bf16(const float* x, int64_t n, int64_t c, int32_t order, int32_t terms, void* y_bf16, void* stream);
/* P[r][v] (bf16, row stride ld) = exp(logits[r][v] - lse[r]) * rowscale[r] for v < V; logits fp16 (row stride ld, written by
 * mtasr_gemm_bf16 mode 1), ld % 8 == 0.  The dense term of d nll / d logits of the CTC head (ref:models/ctc.py:53-54).
 * colsum (V) f32, optional: column sums of P (the dense part of the bias gradient) are ACCUMULATED into it (zero it first). */
int mtasr_softmax_from_logits(const void* logits_f16, const float* lse, const float* rowscale, int64_t rows, int32_t V,
                              int64_t ld, void* P_bf16, float* colsum, void* stream);
/* weight_norm over the last dim (torch.nn.utils.parametrizations.weight_norm(conv, dim=2), hf:48-66): v (R, Kt) f32 with
 * R = all leading dims flattened, g (Kt); w = g v / ||v[:, c]||.  sumsq and dot are (1 + MTASR_WN_PARTS) * Kt floats: [0, Kt)
 * receives the column reduction, the rest is scratch for per-CTA partials summed in a fixed order (deterministic: no
 * atomics); sumsq[0, Kt) from the forward is an input of the backward.  dv (R, Kt), dg (Kt) written.  Kt must divide 256. */
#define MTASR_WN_PARTS 128
int mtasr_weightnorm_fwd(const float* v, const float* g, int64_t R, int32_t Kt, float* w, float* sumsq, void* stream);
int mtasr_weightnorm_bwd(const float* dw, const float* v, const float* g, const float* sumsq, int64_t R, int32_t Kt, float* dv,
                         float* dg, float* dot, void* stream);
/* out[n] = sum_m x[m][n] (bias gradients) */
int mtasr_colsum(const void* x, int32_t dtype, int64_t M, int32_t N, int64_t ld, float* out, void* stream);
/* gru_rel_pos gate of hf:167-176 for every (b, t, head), head_dim 64: w8 (8,64) / b8 (8) = gru_rel_pos_linear as stored
 * (rows 0..3 and 4..7 are summed inside the kernel: view(..., 2, 4).sum(-1)), cst (H) = gru_rel_pos_const; x (B,T,H*64)
 * f32|bf16 is the attention input.  gate (B,H,T) f32 = sigmoid(a) * (sigmoid(b) * cst[h] - 1) + 2.  Backward: dx
 * (B,T,H*64) f32 and / or dab (B*T, H, 2) f32 = the gradients wrt the two pre-sigmoid sums (dx = da * wa + db * wb per head
 * slice: consumers that can add this rank-2 term themselves take dab and pass dx = NULL) written, dw8 (8,64) / db8 (8) /
 * dcst (H) ACCUMULATED (zero them first). */
int mtasr_relpos_gate_fwd(const void* x, int32_t x_dtype, const float* w8, const float* b8, const float* cst, int32_t B,
                          int32_t T, int32_t H, float* gate, void* stream);
int mtasr_relpos_gate_bwd(const void* x, int32_t x_dtype, const float* w8, const float* b8, const float* cst,
                          const float* dgate, int32_t B, int32_t T, int32_t H, float* dx, float* dab, float* dw8, float* db8,
                          float* dcst, void* stream);
/* softmax(S*scale + gate[b,h,q]*table[h,k-q+T-1]) over keys k < klen[b] (hf:167-180 + hf:206-228 fused);
 * S (B,H,T,Tp) f32, gate (B,H,T) f32, table (H,2T-1) f32, klen (B) i32 or NULL, P (B,H,T,Tp) bf16. */
int mtasr_attn_softmax_fwd(const float* S, const float* gate, const float* table, const int32_t* klen, int32_t B,
                           int32_t H, int32_t T, int32_t Tp, float scale, void* P_bf16, void* stream);
/* same softmax, written as the split A-side operand Ps (B,H,T,terms*Tp) bf16 (terms 3 or 6, see mtasr_split_bf16). */
int mtasr_attn_softmax_fwd_split(const float* S, const float* gate, const float* table, const int32_t* klen, int32_t B,
                                 int32_t H, int32_t T, int32_t Tp, float scale, int32_t terms, void* Ps_bf16, void* stream);
/* dS = scale * P*(dP - rowdot) bf16; dgate (B,H,T); dtable (H,2T-1) ACCUMULATED (zero it first). */
int mtasr_attn_softmax_bwd(const void* P_bf16, const float* dP, const float* gate, const float* table, int32_t B,
                           int32_t H, int32_t T, int32_t Tp, float scale, void* dS_bf16, float* dgate, float* dtable,
                           void* stream);
/* Fused attention forward (hf:147-271 + torch MHA hf:206-228), head_dim 64: qkv (B*T, 3*H*64) bf16 = [q | k | v] head-major,
 * gate (B,H,T) f32, table (H,2T-1) f32 (Toeplitz rel-pos bias), klen (B) i32 valid key counts or NULL.
 * out (B*T, H*64) bf16 = softmax_k(q.k*scale + gate*table[k-q+T-1]) v ; lse (B,H,T) f32 row log-sum-exp (for the backward).
 * S / P tiles never leave TMEM / shared memory. */
int mtasr_attn_fwd(const void* qkv_bf16, const float* gate, const float* table, const int32_t* klen, int32_t B, int32_t H,
                   int32_t T, float scale, void* out_bf16, float* lse, const void* drop_seed, uint32_t drop_site,
                   uint32_t drop_keep16, void* stream);
/* Fused attention backward.  out / dout (B*T, H*64) bf16 are the forward output and its gradient; lse from the forward.
 * Writes dqkv (B*T, 3*H*64) bf16 = [dq | dk | dv], dgate (B,H,T) f32 and dtable (H,2T-1) f32; dq32 (B*T, H*64) f32
 * (16-byte aligned) and delta (B,H,T) f32 are scratch.  dq32 / dgate / dtable are zeroed by the call itself (inside the
 * delta pre-pass) and then accumulated into by the backward kernel: the caller does not pre-zero them. */
int mtasr_attn_bwd(const void* qkv_bf16, const void* out_bf16, const void* dout_bf16, const float* lse, const float* gate,
                   const float* table, const int32_t* klen, int32_t B, int32_t H, int32_t T, float scale, void* dqkv_bf16,
                   float* dq32, float* delta, float* dgate, float* dtable, const void* drop_seed, uint32_t drop_site,
                   uint32_t drop_keep16, void* stream);
/* drop_seed (two u32 words in device memory, NULL = off) / drop_site / drop_keep16: dropout on the attention PROBABILITIES
 * (hf:217, attention_dropout): O = (softmax o mask / keep) V, mask index ((b*H + h)*T + q) * ld_even(T) + k, same generator
 * as mtasr_dropout; the backward must receive the forward's triple. */
/* y (B,Tpad,D) bf16 = zero-pad(x (B,T,D), pad_l rows left), rows t >= vlen[b] zeroed when vlen != NULL. */
int mtasr_pad_cast(const void* x, int32_t x_dtype, int32_t B, int32_t T, int32_t D, int32_t pad_l, int32_t Tpad,
                   const int32_t* vlen, void* y_bf16, void* stream);
/* GLU over channel halves of channels-last rows (hf:803-807): x (rows, 2C) -> y (rows, C). */
int mtasr_glu_fwd(const void* x, int32_t x_dtype, int64_t rows, int32_t C, void* y_bf16, float* y_f32, void* stream);
int mtasr_glu_bwd(const void* x, int32_t x_dtype, const void* dy, int32_t dy_dtype, int64_t rows, int32_t C,
                  void* dx_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Persistent LSTM recurrence (ref:models/separator.py:6-59; gate order i,f,g,o; h0 = c0 = 0).
 * xg (B,T,4Hs) f32 = x_t W_ih^T + b for every step (one batched GEMM beforehand); whh (4Hs, ldw) bf16 is the
 * recurrent half W[:, in:] with row stride ldw.  B <= 64, Hs % 16 == 0, Hs/8 <= #SMs (cooperative launch).
 * Outputs: h (B,T,Hs) bf16 (+ optional f32), c (B,T,Hs) f32, gates (B,T,4Hs) f32 activations (saved for BPTT).
 * `barrier` is a 16-byte aligned device scratch area of mtasr_lstm_scratch_bytes(B, Hs, backward) bytes (group counters
 * and the step-flagged exchange words of the recurrence; initialised by the call).  Backward: dgates (B,T,4Hs) bf16 =
 * gradient wrt the gate pre-activations; dx / dW / db follow as GEMMs / column sums over it.
 */
int64_t mtasr_lstm_scratch_bytes(int32_t B, int32_t Hs, int32_t backward);
int mtasr_lstm_fwd(const float* xg, const void* whh_bf16, int32_t ldw, int32_t B, int32_t T, int32_t Hs, void* h_bf16,
                   float* h_f32, float* c_all, float* gates, uint32_t* barrier, void* stream);
int mtasr_lstm_bwd(const float* dh_out, const float* gates, const float* c_all, const void* whh_bf16, int32_t ldw,
                   int32_t B, int32_t T, int32_t Hs, void* dgates_bf16, uint32_t* barrier, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fp32 parity mode (csrc/precise.cu; mtasr_b200.precise): the reference keeps separator, CTC head and loss in fp32
 * (ref:models/losses.py:265-268, ref:models/ctc.py:53, ref:inference_asr.py:120).  Contractions run on the tcgen05 GEMM
 * with split operands (mtasr_split_bf16); what is not a contraction is below.
 *
 * LSTM recurrence of ref:models/separator.py:6-59 in plain fp32: xg (B,T,4Hs) = x W_ih^T + b (precomputed), whh (4Hs, Hs)
 * f32 with row stride ldw.  Forward: h_all, c_all (B,T,Hs), gates_act (B,T,4Hs) post-activation i,f,g,o.  Backward:
 * dgates (B,T,4Hs) f32 = gradient wrt the gate pre-activations; dc_carry (B,Hs) f32 scratch.  One launch per step.
 */
int mtasr_lstm_fwd_f32(const float* xg, const float* whh, int32_t ldw, int32_t B, int32_t T, int32_t Hs, float* h_all,
                       float* c_all, float* gates_act, void* stream);
int mtasr_lstm_bwd_f32(const float* dh_out, const float* whh, int32_t ldw, int32_t B, int32_t T, int32_t Hs,
                       const float* c_all, const float* gates_act, float* dgates, float* dc_carry, void* stream);
/* du = dy where y > 0 else 0 (all f32, n elements). */
int mtasr_relu_bwd_f32(const float* dy, const float* y, int64_t n, float* du, void* stream);
/* out[r][c] = exp(logits[r][c] - lse[r]) * rowscale[r] for c < V, 0 for V <= c < ld (f32, in place allowed): the dense
 * term of d nll / d logits of the CTC head (ref:models/ctc.py:53-54) in fp32. */
int mtasr_softmax_scale_f32(const float* logits, const float* lse, const float* rowscale, int64_t rows, int32_t V,
                            int64_t ld, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Feature-extractor layer 0 (hf:709-751): conv1d(1 -> C0, k, stride) on the waveform x (B,S) f32, weights (C0,1,k)
 * f32.  mode 1 ("layer" norm, WavLM-Large): + LayerNorm over channels + GELU -> y_bf16 (B,L0,C0) channels-last.
 * mode 0 ("group" norm, WavLM-Base+): raw conv (+bias) -> y_f32 (B,L0,C0); follow with mtasr_groupnorm_gelu, which
 * normalises every (b, channel) over time (GroupNorm with C groups), applies the affine and GELU -> bf16.
 */
int mtasr_conv0_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta, float eps,
                    int32_t B, int32_t S, int32_t C0, int32_t k, int32_t stride, int32_t mode, void* y_bf16, float* y_f32,
                    void* stream);
int mtasr_groupnorm_gelu(const float* x, const float* gamma, const float* beta, float eps, int32_t B, int32_t L, int32_t C,
                         float* mean_ws, float* rstd_ws, void* y_bf16, void* stream);
/* same, fp32 output (fp32-accurate parity mode) */
int mtasr_groupnorm_gelu_f32(const float* x, const float* gamma, const float* beta, float eps, int32_t B, int32_t L,
                             int32_t C, float* mean_ws, float* rstd_ws, float* y_f32, void* stream);
/* du = dy * act'(.) as bf16: act 3 = GELU from the saved pre-activation, act 4 = ReLU from the saved output. */
int mtasr_act_bwd(const void* dy, int32_t dy_dtype, const void* src_bf16, int32_t act, int64_t n, void* du_bf16,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTASR_H_ */
