"""Tensor-level wrappers over the C ABI (include/mtasr.h).  torch is used for device memory and streams only.

Every function enqueues hand-written sm_100a kernels on torch's current CUDA stream and returns torch tensors that
own the output buffers.  Nothing here falls back to torch arithmetic: without a CUDA device or without libmtasr.so
the calls raise.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import ctypes as C
import torch

from . import _lib
from ._lib import GemmDesc, check

BF16, F32, F16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_GELU_BWD, ACT_RELU_BWD = 0, 1, 2, 3, 4


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor], offset: int = 0) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.MtasrError("mtasr kernels need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr() + offset * t.element_size()


def empty_act(shape, dtype, device) -> torch.Tensor:
    """Activation buffer with a little slack after the last element: implicit-conv TMA views may address (never
    use) up to one stride-row past the end of the last utterance."""
    n = 1
    for d in shape:
        n *= int(d)
    flat = torch.empty(n + 8192, device=device, dtype=dtype)
    return flat[:n].view(*shape)


def set_sm_budget(n_sms: int) -> int:
    """Cap the SM count the persistent kernels size their grids for (0 = all SMs); returns the effective count."""
    return int(_lib.load().mtasr_set_sm_budget(int(n_sms)))


def launch_count() -> int:
    return int(_lib.load().mtasr_launch_count())


def profile_begin() -> None:
    check(_lib.load().mtasr_profile_begin(), "mtasr_profile_begin")


def profile_end():
    """-> (summed GEMM-kernel ms, executed flops, launches) since profile_begin()."""
    ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
    check(_lib.load().mtasr_profile_end(C.byref(ms), C.byref(fl), C.byref(n)), "mtasr_profile_end")
    return ms.value, fl.value, n.value


# ----------------------------------------------------------------------------------------------------------- GEMM
@dataclass
class Operand:
    """One GEMM operand view: element strides into `t` (bf16).  major 0 = K-major, 1 = MN-major."""
    t: torch.Tensor
    ld: int
    major: int = 0
    sb0: int = 0
    sb1: int = 0
    offset: int = 0
    inner: int = 0      # A only: channels per tap for implicit conv
    phase: int = 1      # A only: conv stride
    rows: int = 0


@dataclass
class Out:
    t: torch.Tensor
    ld: int
    sb0: int = 0
    sb1: int = 0
    offset: int = 0


def gemm(a: Operand, b: Operand, M: int, N: int, K: int, out: Optional[Out], *, batch: Tuple[int, int] = (1, 1),
         bias: Optional[torch.Tensor] = None, bias_sb0: int = 0, act: int = ACT_NONE, residual: Optional[Out] = None,
         aux: Optional[torch.Tensor] = None, alpha: float = 1.0, accumulate: bool = False, mode: int = 0,
         row_vec: Optional[torch.Tensor] = None, row_scale: Optional[torch.Tensor] = None,
         lse_part: Optional[torch.Tensor] = None, block_n: int = 0, drop: Optional[tuple] = None) -> None:
    """C = epilogue(alpha * A @ B^T) on the tcgen05 kernel (csrc/gemm.cu); see include/mtasr.h for the contract."""
    lib = _lib.load()
    if a.t.dtype != torch.bfloat16 or b.t.dtype != torch.bfloat16:
        raise TypeError("gemm operands must be bf16")
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.batch0, d.batch1 = batch
    d.a_major, d.b_major = a.major, b.major
    d.block_n = block_n
    d.a = _p(a.t, a.offset)
    d.a_ld, d.a_sb0, d.a_sb1 = a.ld, a.sb0, a.sb1
    d.a_inner, d.a_phase, d.a_rows = a.inner, a.phase, a.rows
    d.b = _p(b.t, b.offset)
    d.b_ld, d.b_sb0, d.b_sb1, d.b_rows = b.ld, b.sb0, b.sb1, b.rows
    if out is not None:
        d.c = _p(out.t, out.offset)
        d.c_dtype = _dt(out.t)
        d.c_ld, d.c_sb0, d.c_sb1 = out.ld, out.sb0, out.sb1
    if aux is not None:
        if aux.dtype != torch.bfloat16:
            raise TypeError("aux must be bf16")
        d.aux = _p(aux)
    if bias is not None:
        if bias.dtype != torch.float32:
            raise TypeError("bias must be fp32")
        d.bias = _p(bias)
        d.bias_sb0 = bias_sb0
    if residual is not None:
        d.residual = _p(residual.t, residual.offset)
        d.res_dtype = _dt(residual.t)
        d.r_ld, d.r_sb0, d.r_sb1 = residual.ld, residual.sb0, residual.sb1
    d.act = act
    d.alpha = alpha
    d.accumulate = 1 if accumulate else 0
    d.mode = mode
    d.row_vec = _p(row_vec)
    d.row_scale = _p(row_scale)
    d.lse_part = _p(lse_part)
    if drop is not None:                      # (seed (2,) int32 device tensor, site, keep16)
        d.drop_seed, d.drop_site, d.drop_keep16 = _p(drop[0]), int(drop[1]), int(drop[2])
    check(lib.mtasr_gemm_bf16(C.byref(d), _stream()), "mtasr_gemm_bf16")


def gemm_n_tiles(N: int, block_n: int = 0) -> int:
    return int(_lib.load().mtasr_gemm_n_tiles(N, block_n))


def keep16(p: float) -> int:
    """Quantised keep probability of a dropout rate p: round((1 - p) * 65536), at least 1."""
    return max(1, min(65536, int(round((1.0 - float(p)) * 65536.0))))


def dropout(x: torch.Tensor, seed: torch.Tensor, site: int, k16: int, out_dtype=None) -> torch.Tensor:
    """y = x * mask * 65536 / k16 over the trailing dim as columns (mtasr_dropout); mask index = row * ld_even + col."""
    x = x.contiguous()
    cols = x.shape[-1]
    rows = x.numel() // cols
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    check(_lib.load().mtasr_dropout(_p(x), _dt(x), rows, cols, _p(seed), int(site), int(k16), _p(y), _dt(y), _stream()), "mtasr_dropout")
    return y


def linear_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
               residual: Optional[torch.Tensor] = None, out_dtype=torch.bfloat16, want_aux: bool = False,
               out: Optional[torch.Tensor] = None, drop: Optional[tuple] = None):
    """y[M,N] = act(x[M,K] @ w[N,K]^T + bias) (+ residual).  Returns y (and the bf16 pre-activation if want_aux)."""
    M, K = x.shape
    N = w.shape[0]
    y = out if out is not None else torch.empty(M, N, device=x.device, dtype=out_dtype)
    aux = torch.empty(M, N, device=x.device, dtype=torch.bfloat16) if want_aux else None
    gemm(Operand(x, x.stride(0)), Operand(w, w.stride(0)), M, N, K, Out(y, y.stride(0)), bias=bias, act=act,
         residual=None if residual is None else Out(residual, residual.stride(0)), aux=aux, drop=drop)
    return (y, aux) if want_aux else y


def linear_dgrad(dy: torch.Tensor, w: torch.Tensor, *, out_dtype=torch.bfloat16, act: int = ACT_NONE,
                 act_src: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                 accumulate_into: Optional[torch.Tensor] = None, drop: Optional[tuple] = None) -> torch.Tensor:
    """dx[M,K] = dy[M,N] @ w[N,K]; optional fused activation backward (act 3/4 with act_src) or residual add."""
    M, N = dy.shape
    K = w.shape[1]
    dx = accumulate_into if accumulate_into is not None else torch.empty(M, K, device=dy.device, dtype=out_dtype)
    res = None
    if act in (ACT_GELU_BWD, ACT_RELU_BWD):
        res = Out(act_src, act_src.stride(0))
    elif residual is not None:
        res = Out(residual, residual.stride(0))
    gemm(Operand(dy, dy.stride(0)), Operand(w, w.stride(0), major=1), M, K, N, Out(dx, dx.stride(0)), act=act,
         residual=res, accumulate=accumulate_into is not None, drop=drop)
    return dx


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, *, accumulate_into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dw[N,K] (fp32) = dy[M,N]^T @ x[M,K]  (both operands MN-major: no transposes are materialised)."""
    M, N = dy.shape
    K = x.shape[1]
    dw = accumulate_into if accumulate_into is not None else torch.empty(N, K, device=dy.device, dtype=torch.float32)
    gemm(Operand(dy, dy.stride(0), major=1), Operand(x, x.stride(0), major=1), N, K, M, Out(dw, dw.stride(0)),
         accumulate=accumulate_into is not None)
    return dw


# ----------------------------------------------------------------------------------------------------------- rows
def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, out_bf16: bool = True,
                  out_f32: bool = False, post_gelu: bool = False, save_stats: bool = True):
    D = x.shape[-1]
    rows = x.numel() // D
    x = x.contiguous()
    yb = empty_act(x.shape, torch.bfloat16, x.device) if out_bf16 else None
    yf = torch.empty(x.shape, device=x.device, dtype=torch.float32) if out_f32 else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    check(_lib.load().mtasr_layernorm_fwd(_p(x), _dt(x), _p(gamma), _p(beta), eps, rows, D, int(post_gelu), _p(yb), _p(yf),
                                          _p(mean), _p(rstd), _stream()), "mtasr_layernorm_fwd")
    return yb, yf, mean, rstd


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor, gamma: torch.Tensor, *,
                  dres: Optional[torch.Tensor] = None, want_f32: bool = True, want_bf16: bool = False,
                  want_param_grads: bool = True, want_dxsum: bool = False, gate_ab: Optional[torch.Tensor] = None,
                  gate_w8: Optional[torch.Tensor] = None, zeroed: Optional[torch.Tensor] = None):
    """-> dx_f32, dx_bf16, dgamma, dbeta[, dxsum].  One kernel: dx (+ dres) and every requested column reduction
    (dxsum = sum over rows of dx, the bias gradient of the Linear that produced the tensor dx is the gradient of).
    gate_ab (rows, D/64, 2) + gate_w8 (8, 64): dy additionally receives the gru_rel_pos gate's rank-2-per-head path
    (relpos_gate_bwd(want_dx=False))."""
    D = x.shape[-1]
    rows = x.numel() // D
    dy = dy.contiguous()
    dxf = torch.empty(x.shape, device=x.device, dtype=torch.float32) if want_f32 else None
    dxb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    want_dxsum = want_dxsum and (want_f32 or want_bf16)
    nsum = (2 if want_param_grads else 0) + (1 if want_dxsum else 0)
    # one fill for all reductions -- or none: `zeroed` is a caller-provided ZERO fp32 buffer of >= nsum * D elements (a
    # block's backward zeroes the accumulators of all its row kernels with a single fill)
    if nsum and zeroed is not None:
        sums = zeroed[:nsum * D].view(nsum, D)
    else:
        sums = torch.zeros(nsum, D, device=x.device, dtype=torch.float32) if nsum else None
    dg, db = (sums[0], sums[1]) if want_param_grads else (None, None)
    dxs = sums[nsum - 1] if want_dxsum else None
    check(_lib.load().mtasr_layernorm_bwd_sums(_p(dy), _dt(dy), _p(x), _dt(x), _p(mean), _p(rstd), _p(gamma), _p(dres), rows, D,
                                               _p(dxf), _p(dxb), _p(dg), _p(db), _p(dxs), _p(gate_ab), _p(gate_w8), _stream()),
          "mtasr_layernorm_bwd_sums")
    if want_dxsum:
        return dxf, dxb, dg, db, dxs
    return dxf, dxb, dg, db


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    if x.dtype == torch.bfloat16:
        return x
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    if x.numel():
        check(_lib.load().mtasr_cast_f32_bf16(_p(x), _p(y), x.numel(), _stream()), "mtasr_cast_f32_bf16")
    return y


def split_bf16(x: torch.Tensor, chunk: int, order: int, terms: int = 3) -> torch.Tensor:
    """fp32 (..., n*chunk) -> bf16 (..., n*terms*chunk) split operand (order 0 = A side, 1 = B side; include/mtasr.h)."""
    if x.dtype != torch.float32:
        raise TypeError("split_bf16 takes fp32")
    x = x.contiguous()
    if x.shape[-1] % chunk:
        raise ValueError(f"split_bf16: last dim {x.shape[-1]} is not a multiple of the chunk {chunk}")
    y = empty_act(tuple(x.shape[:-1]) + (terms * x.shape[-1],), torch.bfloat16, x.device)
    check(_lib.load().mtasr_split_bf16(_p(x), x.numel(), chunk, order, terms, _p(y), _stream()), "mtasr_split_bf16")
    return y


def split_rows_bf16(x: torch.Tensor, order: int, terms: int = 3) -> torch.Tensor:
    """fp32 (R, C) -> bf16 (terms*R, C): the split parts stacked along the ROWS -- the split operand of a contraction that runs
    over the rows (MN-major GEMM operand): order 0 = [lo; hi; hi] (A side), 1 = [hi; lo; hi] (B side)."""
    R, Cc = x.shape
    return split_bf16(x.reshape(1, R * Cc), R * Cc, order, terms).view(terms * R, Cc)


def lstm_fwd_f32(xg: torch.Tensor, whh: torch.Tensor):
    """fp32 recurrence (csrc/precise.cu): xg (B,T,4Hs) f32, whh (4Hs,Hs) f32 (row stride may exceed Hs) -> h, c, gates_act."""
    B, T, H4 = xg.shape
    Hs = H4 // 4
    h = torch.empty(B, T, Hs, device=xg.device, dtype=torch.float32)
    c = torch.empty(B, T, Hs, device=xg.device, dtype=torch.float32)
    ga = torch.empty(B, T, H4, device=xg.device, dtype=torch.float32)
    check(_lib.load().mtasr_lstm_fwd_f32(_p(xg), _p(whh), whh.stride(0), B, T, Hs, _p(h), _p(c), _p(ga), _stream()), "mtasr_lstm_fwd_f32")
    return h, c, ga


def lstm_bwd_f32(dh: torch.Tensor, whh: torch.Tensor, c: torch.Tensor, gates_act: torch.Tensor) -> torch.Tensor:
    B, T, Hs = dh.shape
    dh = dh.contiguous()
    dg = torch.empty(B, T, 4 * Hs, device=dh.device, dtype=torch.float32)
    carry = torch.empty(B, Hs, device=dh.device, dtype=torch.float32)
    check(_lib.load().mtasr_lstm_bwd_f32(_p(dh), _p(whh), whh.stride(0), B, T, Hs, _p(c), _p(gates_act), _p(dg), _p(carry), _stream()),
          "mtasr_lstm_bwd_f32")
    return dg


def relu_bwd_f32(dy: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    dy = dy.contiguous()
    du = torch.empty_like(dy)
    check(_lib.load().mtasr_relu_bwd_f32(_p(dy), _p(y), dy.numel(), _p(du), _stream()), "mtasr_relu_bwd_f32")
    return du


def softmax_scale_f32_(logits: torch.Tensor, lse: torch.Tensor, rowscale: torch.Tensor, V: int) -> torch.Tensor:
    """In place: logits (rows, ld) f32 -> exp(logit - lse[row]) * rowscale[row] (columns >= V zeroed)."""
    rows, ld = logits.shape
    check(_lib.load().mtasr_softmax_scale_f32(_p(logits), _p(lse), _p(rowscale), rows, V, ld, _p(logits), _stream()),
          "mtasr_softmax_scale_f32")
    return logits


def attn_softmax_fwd_split(S, gate, table, klen, B, H, T, Tp, scale, terms):
    Ps = torch.empty(B, H, T, terms * Tp, device=S.device, dtype=torch.bfloat16)
    check(_lib.load().mtasr_attn_softmax_fwd_split(_p(S), _p(gate), _p(table), _p(klen), B, H, T, Tp, scale, terms, _p(Ps), _stream()),
          "mtasr_attn_softmax_fwd_split")
    return Ps


def softmax_from_logits(logits16: torch.Tensor, lse: torch.Tensor, rowscale: torch.Tensor, V: int, want_colsum: bool = False):
    """(rows, ld) fp16 logits -> (rows, ld) bf16 exp(logit - lse[row]) * rowscale[row] (columns >= V zero), and optionally
    the (V,) fp32 column sums of it."""
    rows, ld = logits16.shape
    P = torch.empty(rows, ld, device=logits16.device, dtype=torch.bfloat16)
    cs = torch.zeros(V, device=logits16.device, dtype=torch.float32) if want_colsum else None
    check(_lib.load().mtasr_softmax_from_logits(_p(logits16), _p(lse), _p(rowscale), rows, V, ld, _p(P), _p(cs), _stream()),
          "mtasr_softmax_from_logits")
    return (P, cs) if want_colsum else P


WN_PARTS = 128      # MTASR_WN_PARTS of include/mtasr.h


def weightnorm_fwd(v: torch.Tensor, g: torch.Tensor):
    """v (..., Kt) f32, g (Kt) -> (w like v, sumsq (Kt))."""
    v = v.contiguous()
    Kt = v.shape[-1]
    w = torch.empty_like(v)
    sumsq = torch.empty((1 + WN_PARTS) * Kt, device=v.device, dtype=torch.float32)   # result + per-CTA partials
    check(_lib.load().mtasr_weightnorm_fwd(_p(v), _p(g), v.numel() // Kt, Kt, _p(w), _p(sumsq), _stream()), "mtasr_weightnorm_fwd")
    return w, sumsq


def weightnorm_bwd(dw: torch.Tensor, v: torch.Tensor, g: torch.Tensor, sumsq: torch.Tensor):
    dw, v = dw.contiguous(), v.contiguous()
    Kt = v.shape[-1]
    dv = torch.empty_like(v)
    dg = torch.empty(Kt, device=v.device, dtype=torch.float32)
    dot = torch.empty((1 + WN_PARTS) * Kt, device=v.device, dtype=torch.float32)
    check(_lib.load().mtasr_weightnorm_bwd(_p(dw), _p(v), _p(g), _p(sumsq), v.numel() // Kt, Kt, _p(dv), _p(dg), _p(dot), _stream()),
          "mtasr_weightnorm_bwd")
    return dv, dg


def colsum(x: torch.Tensor) -> torch.Tensor:
    M, N = x.shape
    out = torch.empty(N, device=x.device, dtype=torch.float32)
    check(_lib.load().mtasr_colsum(_p(x), _dt(x), M, N, x.stride(0), _p(out), _stream()), "mtasr_colsum")
    return out


def relpos_gate_fwd(x, w8, b8, cst, B, T, H):
    """w8 (8,64) / b8 (8) fp32 = gru_rel_pos_linear as stored (the 4-row sums of hf:170-173 happen inside the kernel)."""
    gate = torch.empty(B, H, T, device=x.device, dtype=torch.float32)
    check(_lib.load().mtasr_relpos_gate_fwd(_p(x), _dt(x), _p(w8), _p(b8), _p(cst), B, T, H, _p(gate), _stream()),
          "mtasr_relpos_gate_fwd")
    return gate


GATE_ACC = 520      # floats of the gate kernels' accumulator in front of the H per-head constants


def relpos_gate_bwd(x, w8, b8, cst, dgate, B, T, H, want_dx: bool = True, zeroed: Optional[torch.Tensor] = None):
    """-> dx (B,T,H*64) f32, dw8 (8,64), db8 (8), dcst (H); with want_dx=False the first result is dab (B*T, H, 2) f32 instead
    (the gradients wrt the two pre-sigmoid sums; layernorm_bwd(gate_ab=dab, gate_w8=w8) adds da * wa + db * wb itself)."""
    dx = torch.empty(B, T, H * 64, device=x.device, dtype=torch.float32) if want_dx else None
    dab = None if want_dx else torch.empty(B * T, H, 2, device=x.device, dtype=torch.float32)
    # one fill for the three -- or the caller's ZERO fp32 buffer of >= GATE_ACC + H elements (see layernorm_bwd)
    acc = zeroed[:GATE_ACC + H] if zeroed is not None else torch.zeros(GATE_ACC + H, device=x.device, dtype=torch.float32)
    dw8, db8, dcst = acc[:512].view(8, 64), acc[512:520], acc[520:520 + H]
    check(_lib.load().mtasr_relpos_gate_bwd(_p(x), _dt(x), _p(w8), _p(b8), _p(cst), _p(dgate), B, T, H, _p(dx), _p(dab), _p(dw8),
                                            _p(db8), _p(dcst), _stream()), "mtasr_relpos_gate_bwd")
    return (dx if want_dx else dab), dw8, db8, dcst


def attn_fwd(qkv, gate, table, klen, B, H, T, scale, drop=None):
    """Fused attention forward -> (out (B*T, H*64) bf16, lse (B,H,T) f32).  drop = (seed, site, keep16): dropout on the
    attention probabilities (hf:217), mask index ((b*H + h)*T + q) * ld_even(T) + k."""
    out = torch.empty(B * T, H * 64, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=qkv.device, dtype=torch.float32)
    ds, dsite, dk = (_p(drop[0]), int(drop[1]), int(drop[2])) if drop is not None else (None, 0, 65536)
    check(_lib.load().mtasr_attn_fwd(_p(qkv), _p(gate), _p(table), _p(klen), B, H, T, scale, _p(out), _p(lse), ds, dsite, dk, _stream()),
          "mtasr_attn_fwd")
    return out, lse


def attn_bwd(qkv, out, dout, lse, gate, table, klen, B, H, T, scale, drop=None):
    """Fused attention backward -> (dqkv (B*T, 3*H*64) bf16, dgate (B,H,T) f32, dtable (H,2T-1) f32)."""
    dev = qkv.device
    D = H * 64
    dqkv = torch.empty(B * T, 3 * D, device=dev, dtype=torch.bfloat16)
    dq32 = torch.empty(B * T, D, device=dev, dtype=torch.float32)            # zeroed by the call (delta pre-pass), like acc
    delta = torch.empty(B, H, T, device=dev, dtype=torch.float32)
    n_g = (B * H * T + 3) // 4 * 4
    acc = torch.empty(n_g + H * (2 * T - 1), device=dev, dtype=torch.float32)
    dgate, dtable = acc[:B * H * T].view(B, H, T), acc[n_g:].view(H, 2 * T - 1)
    ds, dsite, dk = (_p(drop[0]), int(drop[1]), int(drop[2])) if drop is not None else (None, 0, 65536)
    check(_lib.load().mtasr_attn_bwd(_p(qkv), _p(out), _p(dout), _p(lse), _p(gate), _p(table), _p(klen), B, H, T, scale, _p(dqkv),
                                     _p(dq32), _p(delta), _p(dgate), _p(dtable), ds, dsite, dk, _stream()), "mtasr_attn_bwd")
    return dqkv, dgate, dtable


def attn_softmax_fwd(S, gate, table, klen, B, H, T, Tp, scale):
    P = torch.empty(B, H, T, Tp, device=S.device, dtype=torch.bfloat16)
    check(_lib.load().mtasr_attn_softmax_fwd(_p(S), _p(gate), _p(table), _p(klen), B, H, T, Tp, scale, _p(P), _stream()),
          "mtasr_attn_softmax_fwd")
    return P


def attn_softmax_bwd(P, dP, gate, table, B, H, T, Tp, scale):
    dS = torch.empty(B, H, T, Tp, device=P.device, dtype=torch.bfloat16)
    dgate = torch.empty(B, H, T, device=P.device, dtype=torch.float32)
    dtable = torch.zeros(H, 2 * T - 1, device=P.device, dtype=torch.float32)
    check(_lib.load().mtasr_attn_softmax_bwd(_p(P), _p(dP), _p(gate), _p(table), B, H, T, Tp, scale, _p(dS), _p(dgate),
                                             _p(dtable), _stream()), "mtasr_attn_softmax_bwd")
    return dS, dgate, dtable


def pad_cast(x: torch.Tensor, pad_l: int, Tpad: int, vlen: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, T, D = x.shape
    x = x.contiguous()
    y = empty_act((B, Tpad, D), torch.bfloat16, x.device)
    check(_lib.load().mtasr_pad_cast(_p(x), _dt(x), B, T, D, pad_l, Tpad, _p(vlen), _p(y), _stream()), "mtasr_pad_cast")
    return y


def glu_fwd(x: torch.Tensor, *, out_f32: bool = False):
    C2 = x.shape[-1]
    Cc = C2 // 2
    rows = x.numel() // C2
    x = x.contiguous()
    yb = empty_act(tuple(x.shape[:-1]) + (Cc,), torch.bfloat16, x.device)
    yf = torch.empty(*x.shape[:-1], Cc, device=x.device, dtype=torch.float32) if out_f32 else None
    check(_lib.load().mtasr_glu_fwd(_p(x), _dt(x), rows, Cc, _p(yb), _p(yf), _stream()), "mtasr_glu_fwd")
    return yb, yf


def glu_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    C2 = x.shape[-1]
    rows = x.numel() // C2
    dy = dy.contiguous()
    dx = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    check(_lib.load().mtasr_glu_bwd(_p(x), _dt(x), _p(dy), _dt(dy), rows, C2 // 2, _p(dx), _stream()), "mtasr_glu_bwd")
    return dx


# ----------------------------------------------------------------------------------------------------------- CTC
def ctc_state_pad(max_label_len: int) -> int:
    sp = int(_lib.load().mtasr_ctc_state_pad(max_label_len))
    if sp < 0:
        raise _lib.MtasrError(f"CTC label length {max_label_len} > 255 is not supported by the warp-per-utterance kernel")
    return sp


def ctc_alpha_fwd(glog, lse, ys, hlens, ylens, max_label_len):
    B, T, Lp = glog.shape
    sp = ctc_state_pad(max_label_len)
    alpha = torch.empty(B, T, sp, device=glog.device, dtype=torch.float32)
    coff = torch.empty(B, T, device=glog.device, dtype=torch.float64)
    nll = torch.empty(B, device=glog.device, dtype=torch.float32)
    nll_raw = torch.empty(B, device=glog.device, dtype=torch.float64)
    ys_ld = ys.stride(0) if ys.numel() else 0
    check(_lib.load().mtasr_ctc_alpha_fwd(_p(glog), _p(lse), _p(ys) if ys.numel() else None, _p(hlens), _p(ylens), B, T, Lp,
                                          ys_ld, max_label_len, _p(alpha), _p(coff), _p(nll), _p(nll_raw), _stream()),
          "mtasr_ctc_alpha_fwd")
    return nll, nll_raw, alpha, coff


def ctc_beta_bwd(glog, lse, ys, hlens, ylens, max_label_len, alpha, coff, nll_raw, gout):
    B, T, Lp = glog.shape
    dG = torch.zeros(B, T, Lp, device=glog.device, dtype=torch.float32)     # the kernel writes reachable (t, column) only
    rowscale = torch.empty(B, T, device=glog.device, dtype=torch.float32)
    ys_ld = ys.stride(0) if ys.numel() else 0
    check(_lib.load().mtasr_ctc_beta_bwd(_p(glog), _p(lse), _p(ys) if ys.numel() else None, _p(hlens), _p(ylens), B, T, Lp,
                                         ys_ld, max_label_len, _p(alpha), _p(coff), _p(nll_raw), _p(gout), _p(dG),
                                         _p(rowscale), _stream()), "mtasr_ctc_beta_bwd")
    return dG, rowscale


def lse_finalize(part: torch.Tensor, rows: int, n_tiles: int, want_lse=True, want_argmax=False):
    lse = torch.empty(rows, device=part.device, dtype=torch.float32) if want_lse else None
    am = torch.empty(rows, device=part.device, dtype=torch.int64) if want_argmax else None
    check(_lib.load().mtasr_lse_finalize(_p(part), rows, n_tiles, _p(lse), _p(am), _stream()), "mtasr_lse_finalize")
    return lse, am


def ctc_collapse(ids: torch.Tensor, blank_id: int, pad_id: int):
    B, T = ids.shape
    ids = ids.contiguous()
    out = torch.empty(B, T, device=ids.device, dtype=torch.int64)
    lens = torch.empty(B, device=ids.device, dtype=torch.int32)
    check(_lib.load().mtasr_ctc_collapse(_p(ids), B, T, blank_id, pad_id, _p(out), _p(lens), _stream()), "mtasr_ctc_collapse")
    return out, lens


def ctc_segments(path: torch.Tensor, mask: torch.Tensor, blank: int):
    B, T = path.shape
    path = path.contiguous().to(torch.int64)
    m8 = mask.contiguous().to(torch.uint8)
    ss = torch.empty(B, T, device=path.device, dtype=torch.int32)
    se = torch.empty(B, T, device=path.device, dtype=torch.int32)
    n = torch.empty(B, device=path.device, dtype=torch.int32)
    check(_lib.load().mtasr_ctc_segments(_p(path), _p(m8), B, T, blank, _p(ss), _p(se), _p(n), _stream()), "mtasr_ctc_segments")
    return ss, se, n


def segment_mean_fwd(x: torch.Tensor, pblank, ss, se, n, Lmax: int):
    B, T, D = x.shape
    out = torch.empty(B, Lmax, D, device=x.device, dtype=torch.float32)
    conf = torch.empty(B, Lmax, device=x.device, dtype=torch.float32) if pblank is not None else None
    check(_lib.load().mtasr_segment_mean_fwd(_p(x), _p(pblank), _p(ss), _p(se), _p(n), B, T, D, Lmax, _p(out), _p(conf), _stream()),
          "mtasr_segment_mean_fwd")
    return out, conf


def segment_mean_bwd(dout: torch.Tensor, ss, se, n, T: int):
    B, Lmax, D = dout.shape
    dx = torch.zeros(B, T, D, device=dout.device, dtype=torch.float32)
    check(_lib.load().mtasr_segment_mean_bwd(_p(dout), _p(ss), _p(se), _p(n), B, T, D, Lmax, _p(dx), _stream()), "mtasr_segment_mean_bwd")
    return dx


def split_labels(labels: torch.Tensor, K: int, sep_id: int, pad_id, ignore_id, end_id, allow_empty: bool):
    """Device-side label splitter (mtasr_split_labels): -> out (K,B,L) i64, lens (K,B) i64, status (3) i32; no host sync."""
    B, L = labels.shape
    labels = labels.to(torch.int64)
    if labels.stride(1) != 1:
        labels = labels.contiguous()
    fill = int(pad_id) if pad_id is not None else 0
    out = torch.full((K, B, max(L, 1)), fill, device=labels.device, dtype=torch.int64)
    lens = torch.empty(K, B, device=labels.device, dtype=torch.int64)
    status = torch.tensor([2 ** 31 - 1, 0, 0], device=labels.device, dtype=torch.int32)
    opt = lambda v: (int(v), 1) if v is not None else (0, 0)
    (pv, hp), (iv, hi), (ev, he) = opt(pad_id), opt(ignore_id), opt(end_id)
    check(_lib.load().mtasr_split_labels(_p(labels), B, L, labels.stride(0), K, int(sep_id), pv, hp, iv, hi, ev, he, int(bool(allow_empty)),
                                         _p(out), _p(lens), _p(status), _stream()), "mtasr_split_labels")
    return out, lens, status


def pcgrad_project_(gi: torch.Tensor, gj: torch.Tensor, scratch: torch.Tensor) -> None:
    """gi -= (<gi,gj> < 0 ? <gi,gj> / (<gj,gj> + 1e-12) : 0) * gj on flat fp32 vectors, decided on the device."""
    check(_lib.load().mtasr_pcgrad_dots(_p(gi), _p(gj), gi.numel(), _p(scratch), _stream()), "mtasr_pcgrad_dots")
    check(_lib.load().mtasr_pcgrad_project(_p(gi), _p(gj), gi.numel(), _p(scratch), _stream()), "mtasr_pcgrad_project")


def ctc_gather_cols(dense, ys, ylens, Lp, blank):
    B, T, V = dense.shape
    out = torch.empty(B, T, Lp, device=dense.device, dtype=torch.float32)
    ys_ld = ys.stride(0) if ys.numel() else 0
    check(_lib.load().mtasr_ctc_gather_cols(_p(dense), _p(ys) if ys.numel() else None, _p(ylens), B, T, V, Lp, ys_ld, blank,
                                            _p(out), _stream()), "mtasr_ctc_gather_cols")
    return out


def ctc_scatter_cols(src, ys, ylens, dense, blank):
    B, T, Lp = src.shape
    V = dense.shape[-1]
    ys_ld = ys.stride(0) if ys.numel() else 0
    check(_lib.load().mtasr_ctc_scatter_cols(_p(src), _p(ys) if ys.numel() else None, _p(ylens), B, T, V, Lp, ys_ld, blank,
                                             _p(dense), _stream()), "mtasr_ctc_scatter_cols")


def ctc_gather_rows(w_bf16, bias, ys, ylens, Lp, blank):
    B = ylens.numel()
    D = w_bf16.shape[1]
    wg = torch.empty(B, Lp, D, device=w_bf16.device, dtype=torch.bfloat16)
    bg = torch.empty(B, Lp, device=w_bf16.device, dtype=torch.float32)
    ys_ld = ys.stride(0) if ys.numel() else 0
    check(_lib.load().mtasr_ctc_gather_rows(_p(w_bf16), _p(bias), _p(ys) if ys.numel() else None, _p(ylens), B, Lp, D, ys_ld,
                                            blank, w_bf16.shape[0], _p(wg), _p(bg), _stream()), "mtasr_ctc_gather_rows")
    return wg, bg


def ctc_scatter_rows(dwg, dbg, ys, ylens, blank, dw, db):
    B, Lp, D = dwg.shape
    ys_ld = ys.stride(0) if ys.numel() else 0
    check(_lib.load().mtasr_ctc_scatter_rows(_p(dwg), _p(dbg), _p(ys) if ys.numel() else None, _p(ylens), B, Lp, D, ys_ld,
                                             blank, dw.shape[0], _p(dw), _p(db), _stream()), "mtasr_ctc_scatter_rows")


# ----------------------------------------------------------------------------------------------------------- misc
def act_bwd(dy: torch.Tensor, src_bf16: torch.Tensor, act: int) -> torch.Tensor:
    dy = dy.contiguous()
    du = torch.empty(dy.shape, device=dy.device, dtype=torch.bfloat16)
    check(_lib.load().mtasr_act_bwd(_p(dy), _dt(dy), _p(src_bf16), act, dy.numel(), _p(du), _stream()), "mtasr_act_bwd")
    return du


def conv0_fwd(x: torch.Tensor, w: torch.Tensor, bias, gamma, beta, eps: float, k: int, stride: int, layer_norm: bool):
    """First feature-extractor conv (+LayerNorm+GELU when layer_norm) -> (B, L0, C0); bf16 if fused else raw fp32."""
    B, S = x.shape
    C0 = w.shape[0]
    L0 = (S - k) // stride + 1
    x = x.contiguous()
    if layer_norm:
        y = empty_act((B, L0, C0), torch.bfloat16, x.device)
        check(_lib.load().mtasr_conv0_fwd(_p(x), _p(w), _p(bias), _p(gamma), _p(beta), eps, B, S, C0, k, stride, 1, _p(y), None,
                                          _stream()), "mtasr_conv0_fwd")
    else:
        y = torch.empty(B, L0, C0, device=x.device, dtype=torch.float32)
        check(_lib.load().mtasr_conv0_fwd(_p(x), _p(w), _p(bias), None, None, eps, B, S, C0, k, stride, 0, None, _p(y),
                                          _stream()), "mtasr_conv0_fwd")
    return y


def groupnorm_gelu(x: torch.Tensor, gamma, beta, eps: float, out_f32: bool = False) -> torch.Tensor:
    B, L, Cc = x.shape
    mean = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
    rstd = torch.empty(B, Cc, device=x.device, dtype=torch.float32)
    if out_f32:
        y = torch.empty(B, L, Cc, device=x.device, dtype=torch.float32)
        check(_lib.load().mtasr_groupnorm_gelu_f32(_p(x), _p(gamma), _p(beta), eps, B, L, Cc, _p(mean), _p(rstd), _p(y), _stream()),
              "mtasr_groupnorm_gelu_f32")
        return y
    y = empty_act((B, L, Cc), torch.bfloat16, x.device)
    check(_lib.load().mtasr_groupnorm_gelu(_p(x), _p(gamma), _p(beta), eps, B, L, Cc, _p(mean), _p(rstd), _p(y), _stream()),
          "mtasr_groupnorm_gelu")
    return y


def lstm_fwd(xg: torch.Tensor, whh: torch.Tensor, ldw: int, want_h_f32: bool = False):
    """xg (B,T,4Hs) f32; whh bf16 view of W[:, in:] (row stride ldw).  Returns h_bf16, h_f32|None, c, gates."""
    B, T, H4 = xg.shape
    Hs = H4 // 4
    dev = xg.device
    h = torch.empty(B, T, Hs, device=dev, dtype=torch.bfloat16)
    hf = torch.empty(B, T, Hs, device=dev, dtype=torch.float32) if want_h_f32 else None
    c = torch.empty(B, T, Hs, device=dev, dtype=torch.float32)
    gates = torch.empty(B, T, H4, device=dev, dtype=torch.float32)
    bar = torch.empty(int(_lib.load().mtasr_lstm_scratch_bytes(B, Hs, 0)), device=dev, dtype=torch.uint8)   # initialised by the call
    check(_lib.load().mtasr_lstm_fwd(_p(xg), _p(whh), ldw, B, T, Hs, _p(h), _p(hf), _p(c), _p(gates), _p(bar), _stream()),
          "mtasr_lstm_fwd")
    return h, hf, c, gates


def lstm_bwd(dh: torch.Tensor, gates: torch.Tensor, c: torch.Tensor, whh: torch.Tensor, ldw: int) -> torch.Tensor:
    B, T, Hs = dh.shape
    dh = dh.contiguous()
    dgates = torch.empty(B, T, 4 * Hs, device=dh.device, dtype=torch.bfloat16)
    bar = torch.empty(int(_lib.load().mtasr_lstm_scratch_bytes(B, Hs, 1)), device=dh.device, dtype=torch.uint8)
    check(_lib.load().mtasr_lstm_bwd(_p(dh), _p(gates), _p(c), _p(whh), ldw, B, T, Hs, _p(dgates), _p(bar), _stream()),
          "mtasr_lstm_bwd")
    return dgates
