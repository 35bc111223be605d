"""fp32-accurate forward of the encoder path (north_star: "encoder hidden states <= 1e-4 relative in fp32").

The reference's fp32 run is torch fp32 matmul outside autocast (ref:models/modeling_wavlm.py:412-465 calling hf:48-373).
Blackwell's tensor cores have no fp32-operand mode, so every contraction here is the same tcgen05 bf16 GEMM kernel run on
*split* operands (`mtasr_split_bf16`): x = hi + lo (two bf16 values), A rows become [lo|hi|hi], B rows [hi|lo|hi], and
one GEMM over the 3x longer contraction dimension accumulates a_lo b_hi + a_hi b_lo + a_hi b_hi in fp32 (operand error
~2^-17 per product instead of 2^-8).  Activations stay fp32 between kernels; attention materialises S (fp32) and the
split P, because the fused attention kernel keeps P in bf16.

Measured on B200 (tools/diag_fp32.py, against the oracle in fp64): the residual error (1.5e-5 after the conv stack,
1.6e-5 encoder output, 3.7e-5 after the adapter; the oracle's own fp32 run sits at 2-4e-6) is NOT the operand split --
the three-way split (MTASR_FP32_TERMS=6: x1+x2+x3, all products down to 2^-24) is no better -- but the tensor core's
fp32 accumulation, which truncates at every K=16 step, so the error grows with the length of the accumulation chain.
That is why the small products come first (they are accumulated while the accumulator is still ~2^-8 of its final
size), and why the two-way split is the default.

This is a parity / validation mode: 3x the tensor work, fp32 activations, forward only (it raises under autograd).
Enable with `mtasr_b200.precise.set_precision("fp32")`, the `precise.precision("fp32")` context manager, or
MTASR_PRECISION=fp32.
"""
import os
from contextlib import contextmanager

import torch
import torch.nn.functional as F

from . import kernels as K

BF, F32 = torch.bfloat16, torch.float32

_precision = os.environ.get("MTASR_PRECISION", "bf16").lower()
if _precision not in ("bf16", "fp32"):
    raise ValueError(f"MTASR_PRECISION must be bf16 or fp32, got {_precision!r}")


R = int(os.environ.get("MTASR_FP32_TERMS", "3"))      # operand split: 3 = two-way (hi, lo), 6 = three-way
if R not in (3, 6):
    raise ValueError("MTASR_FP32_TERMS must be 3 or 6")


def get_precision() -> str:
    return _precision


def set_precision(mode: str) -> None:
    global _precision
    if mode not in ("bf16", "fp32"):
        raise ValueError(f"precision must be 'bf16' or 'fp32', got {mode!r}")
    _precision = mode


@contextmanager
def precision(mode: str):
    prev = get_precision()
    set_precision(mode)
    try:
        yield
    finally:
        set_precision(prev)


def _a(x: torch.Tensor, chunk: int = 0) -> torch.Tensor:
    return K.split_bf16(x.float(), chunk or x.shape[-1], 0, R)


def _b(w: torch.Tensor, chunk: int = 0) -> torch.Tensor:
    return K.split_bf16(w.detach().float(), chunk or w.shape[-1], 1, R)


def _f(t):
    return None if t is None else t.detach().float().contiguous()


def linear(x: torch.Tensor, w: torch.Tensor, b=None, *, act: int = K.ACT_NONE, residual=None) -> torch.Tensor:
    """fp32 y = act(x W^T + b) (+ residual); x (..., K) fp32, w (N, K)."""
    shp = x.shape
    Kd = shp[-1]
    x2 = x.contiguous().view(-1, Kd)
    M, N = x2.shape[0], w.shape[0]
    y = torch.empty(M, N, device=x.device, dtype=F32)
    res = None if residual is None else K.Out(residual.contiguous().view(M, N), N)
    K.gemm(K.Operand(_a(x2), R * Kd), K.Operand(_b(w), R * Kd), M, N, R * Kd, K.Out(y, N), bias=_f(b), act=act, residual=res)
    return y.view(*shp[:-1], N)


def layer_norm(x, ln, post_gelu=False):
    return K.layernorm_fwd(x, _f(ln.weight), _f(ln.bias), ln.eps, out_bf16=False, out_f32=True, post_gelu=post_gelu,
                           save_stats=False)[1]


def feature_extractor(model, x: torch.Tensor) -> torch.Tensor:
    """hf:754-789 -> (B, T, C) channels-last fp32."""
    cfg = model.config
    layer_norm_mode = cfg.feat_extract_norm == "layer"
    l0 = model.feature_extractor.conv_layers[0]
    raw = K.conv0_fwd(x.float(), _f(l0.conv.weight), _f(l0.conv.bias), None, None, 0.0, cfg.conv_kernel[0], cfg.conv_stride[0], False)
    if layer_norm_mode:
        y = layer_norm(raw, l0.layer_norm, post_gelu=True)
    else:
        y = K.groupnorm_gelu(raw, _f(l0.layer_norm.weight), _f(l0.layer_norm.bias), l0.layer_norm.eps, out_f32=True)
    for i in range(1, len(model.feature_extractor.conv_layers)):
        lyr = model.feature_extractor.conv_layers[i]
        B, L, C = y.shape
        k, s = cfg.conv_kernel[i], cfg.conv_stride[i]
        Cout = lyr.conv.weight.shape[0]
        wk = _b(lyr.conv.weight.detach().float().permute(0, 2, 1).contiguous(), C).view(Cout, k * R * C)
        Lout = (L - k) // s + 1
        C3 = R * C
        a = K.Operand(_a(y), s * C3, sb1=L * C3, inner=C3, phase=s, rows=(L + s - 1) // s)
        pre = torch.empty(B, Lout, Cout, device=y.device, dtype=F32)
        K.gemm(a, K.Operand(wk, k * C3), Lout, Cout, k * C3, K.Out(pre, Cout, sb1=Lout * Cout), batch=(1, B), bias=_f(lyr.conv.bias),
               act=K.ACT_NONE if layer_norm_mode else K.ACT_GELU)
        y = layer_norm(pre, lyr.layer_norm, post_gelu=True) if layer_norm_mode else pre
    return y


def pos_conv(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, G: int) -> torch.Tensor:
    """x + GELU(grouped_conv1d(x, w, bias, k, pad=k//2)[:T]) (hf:48-90 + hf:403/481), fp32."""
    B, T, D = x.shape
    cg = D // G
    cgp = (cg + 63) // 64 * 64
    k = w.shape[2]
    pad = k // 2
    Tpad = T + k
    xp = F.pad(x.float(), (0, 0, pad, Tpad - T - pad))                      # zero padding: data movement only
    xg = xp.view(B, Tpad, G, cg)
    if cgp != cg:
        xg = F.pad(xg, (0, cgp - cg))
    xs = _a(xg.reshape(B, Tpad, G * cgp), cgp)                              # (B, Tpad, G*R*cgp)
    wk = torch.zeros(G, cg, k, cgp, device=x.device, dtype=F32)             # (g, out, tap, c) K-major
    wk[..., :cg] = w.detach().float().view(G, cg, cg, k).permute(0, 1, 3, 2)
    ws = _b(wk, cgp)                                                        # (G, cg, k, R*cgp)
    c3 = R * cgp
    y = torch.empty(B, T, D, device=x.device, dtype=F32)
    xres = x.float().contiguous()
    K.gemm(K.Operand(xs, G * c3, sb0=c3, sb1=Tpad * G * c3, inner=c3, phase=1, rows=Tpad),
           K.Operand(ws, k * c3, sb0=cg * k * c3), T, cg, k * c3, K.Out(y, D, sb0=cg, sb1=T * D), batch=(G, B),
           bias=_f(bias), bias_sb0=cg, act=K.ACT_GELU, residual=K.Out(xres, D, sb0=cg, sb1=T * D))
    return y


def attention(h: torch.Tensor, res: torch.Tensor, at, gate: torch.Tensor, table: torch.Tensor, klen, H: int) -> torch.Tensor:
    """res + out_proj(softmax(QK^T/sqrt(d) + gate*relpos + key mask) V), fp32 activations, split contractions."""
    B, T, D = h.shape
    d = D // H
    if d % 8:
        raise NotImplementedError("mtasr_b200 fp32 mode: head_dim must be a multiple of 8")
    wqkv = torch.cat([at.q_proj.weight, at.k_proj.weight, at.v_proj.weight], 0)
    bqkv = torch.cat([at.q_proj.bias, at.k_proj.bias, at.v_proj.bias], 0)
    qkv = linear(h.view(B * T, D), wqkv, bqkv)                              # (B*T, 3D) fp32
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    d3 = R * d
    qs = _a(q.contiguous(), d)                                              # (B*T, H*R*d): split per head
    ks = _b(k.contiguous(), d)                                              # (B*T, H*R*d)
    Tp = (T + 7) // 8 * 8
    S = torch.empty(B, H, T, Tp, device=h.device, dtype=F32)
    K.gemm(K.Operand(qs, H * d3, sb0=d3, sb1=T * H * d3, rows=T), K.Operand(ks, H * d3, sb0=d3, sb1=T * H * d3, rows=T),
           T, T, d3, K.Out(S, Tp, sb0=T * Tp, sb1=H * T * Tp), batch=(H, B))
    scale = float(d) ** -0.5
    Ps = K.attn_softmax_fwd_split(S, gate.detach().float().contiguous(), table.detach().float().contiguous(), klen, B, H, T, Tp, scale, R)
    del S
    # V stacked along the contraction (key) dimension: R row blocks of Tp keys per (utterance, head)
    vp = F.pad(v.contiguous().view(B, T, H, d).permute(0, 2, 1, 3), (0, 0, 0, Tp - T)).contiguous()      # (B,H,Tp,d)
    vs = K.split_bf16(vp, d, 1, R).view(B, H, Tp, R, d).permute(0, 1, 3, 2, 4).contiguous().view(B, H, R * Tp, d)
    O = torch.empty(B * T, D, device=h.device, dtype=F32)
    K.gemm(K.Operand(Ps, R * Tp, sb0=T * R * Tp, sb1=H * T * R * Tp), K.Operand(vs, d, major=1, sb0=R * Tp * d, sb1=H * R * Tp * d, rows=R * Tp),
           T, d, R * Tp, K.Out(O, D, sb0=d, sb1=T * D), batch=(H, B))
    y = linear(O, at.out_proj.weight, at.out_proj.bias, residual=res.contiguous().view(B * T, D))
    return y.view(B, T, D)


def ffn(h: torch.Tensor, res: torch.Tensor, ff) -> torch.Tensor:
    a = linear(h, ff.intermediate_dense.weight, ff.intermediate_dense.bias, act=K.ACT_GELU)
    return linear(a, ff.output_dense.weight, ff.output_dense.bias, residual=res)


def adapter(mod, x: torch.Tensor):
    """3 x [conv1d(D -> 2D, k3 s2 p1) -> GLU], tap after layer index 1 (ref:models/modeling_wavlm.py:223-254), fp32."""
    h = x.float()
    if mod.proj is not None:
        h = layer_norm(linear(h, mod.proj.weight, mod.proj.bias), mod.proj_layer_norm)
    tap = None
    for i, layer in enumerate(mod.layers):
        conv = layer.conv
        B, T, D = h.shape
        k, s, pad = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        Tpad = T + 2 * pad
        Tpad += (-Tpad) % s
        xs = _a(F.pad(h, (0, 0, pad, Tpad - T - pad)))                      # (B, Tpad, R*D)
        Lout = (T + 2 * pad - k) // s + 1
        C2 = conv.weight.shape[0]
        D3 = R * D
        wk = _b(conv.weight.detach().float().permute(0, 2, 1).contiguous(), D).view(C2, k * D3)
        y = torch.empty(B, Lout, C2, device=h.device, dtype=F32)
        K.gemm(K.Operand(xs, s * D3, sb1=Tpad * D3, inner=D3, phase=s, rows=Tpad // s), K.Operand(wk, k * D3), Lout, C2, k * D3,
               K.Out(y, C2, sb1=Lout * C2), batch=(1, B), bias=_f(conv.bias))
        h = K.glu_fwd(y, out_f32=True)[1]
        if i == 1:
            tap = h
    return h, tap


# ======================================================================================================================
# fp32-class Separator and CTC head, forward AND backward  (ref:models/separator.py:151-166, ref:models/ctc.py:129-160,
# ref:models/losses.py:265-268: the reference keeps this part of the path in fp32 even under AMP, and decodes with fp32
# weights, ref:inference_asr.py:120).  Every contraction -- forward, input gradient, weight gradient -- is the tcgen05 GEMM
# on split operands: contractions over the feature dimension split along it (`_a` / `_b`), contractions over the rows (the
# weight gradients, and x W for the input gradients) use the row-stacked split (`K.split_rows_bf16`) as MN-major operands.
# The LSTM recurrence, the ReLU backward and the dense softmax term of the CTC gradient are plain fp32 kernels
# (csrc/precise.cu); LayerNorm, the CTC lattice (alpha / beta) and the column sums are fp32 in the throughput path already.
# ======================================================================================================================
def _lin32(x2: torch.Tensor, w: torch.Tensor, b=None, act: int = K.ACT_NONE) -> torch.Tensor:
    """(M, Kd) f32 @ (N, Kd)^T (+ b) -> (M, N) f32."""
    M, Kd = x2.shape
    N = w.shape[0]
    y = torch.empty(M, N, device=x2.device, dtype=F32)
    K.gemm(K.Operand(_a(x2), R * Kd), K.Operand(_b(w), R * Kd), M, N, R * Kd, K.Out(y, N), bias=_f(b), act=act)
    return y


def _dgrad32(du: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """(M, N) f32 @ (N, Kd) -> (M, Kd) f32: contraction over N = the ROWS of w (row-stacked split, MN-major B operand)."""
    M, N = du.shape
    Kd = w.shape[1]
    out = torch.empty(M, Kd, device=du.device, dtype=F32)
    K.gemm(K.Operand(_a(du), R * N), K.Operand(K.split_rows_bf16(w.detach().float().contiguous(), 1, R), Kd, major=1), M, Kd, R * N,
           K.Out(out, Kd))
    return out


def _wgrad32(du: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """(M, N)^T @ (M, Kd) -> (N, Kd) f32: contraction over the rows of both operands."""
    M, N = du.shape
    Kd = x2.shape[1]
    out = torch.empty(N, Kd, device=du.device, dtype=F32)
    K.gemm(K.Operand(K.split_rows_bf16(du.contiguous(), 0, R), N, major=1), K.Operand(K.split_rows_bf16(x2.contiguous(), 1, R), Kd, major=1),
           N, Kd, R * M, K.Out(out, Kd))
    return out


class LinearF32Fn(torch.autograd.Function):
    """y = act(x W^T + b) with fp32 activations and split-operand contractions; act in {none, ReLU}."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        if act not in (K.ACT_NONE, K.ACT_RELU):
            raise NotImplementedError("mtasr_b200 fp32 mode: LinearF32Fn supports no activation or ReLU")
        shp = x.shape
        x2 = x.contiguous().view(-1, shp[-1]).float()
        y = _lin32(x2, w, b, act)
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(x2, w, y if act == K.ACT_RELU else None)
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        N = w.shape[0]
        du = dy.contiguous().view(-1, N).float()
        if ctx.act == K.ACT_RELU:
            du = K.relu_bwd_f32(du, y)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad32(du, w).view(*dy.shape[:-1], w.shape[1]).to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            dw = _wgrad32(du, x2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(du)
        return dx, dw, db, None


class LSTMLayerF32Fn(torch.autograd.Function):
    """One layer of ref:models/separator.py:27-59 in fp32: input half as one split-operand GEMM, recurrence in the fp32
    step kernels (csrc/precise.cu), BPTT gradients wrt x, W = [W_ih | W_hh] and b."""

    @staticmethod
    def forward(ctx, x, W, b):
        B, T, In = x.shape
        Hs = W.shape[0] // 4
        x2 = x.contiguous().view(B * T, In).float()
        Wf = W.detach().float()
        w_in, whh = Wf[:, :In].contiguous(), Wf[:, In:].contiguous()
        xg = _lin32(x2, w_in, b)
        h, c, ga = K.lstm_fwd_f32(xg.view(B, T, 4 * Hs), whh)
        ctx.dims = (B, T, In, Hs)
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(x2, w_in, whh, h, c, ga)
        return h

    @staticmethod
    def backward(ctx, dh):
        x2, w_in, whh, h, c, ga = ctx.saved_tensors
        B, T, In, Hs = ctx.dims
        dg2 = K.lstm_bwd_f32(dh.float(), whh, c, ga).view(B * T, 4 * Hs)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = _dgrad32(dg2, w_in).view(B, T, In).to(ctx.x_dtype)
        if ctx.needs_input_grad[1]:
            hprev = torch.zeros_like(h)                                        # h_{t-1} next to dgates_t (data movement)
            hprev[:, 1:] = h[:, :-1]
            dW = torch.cat([_wgrad32(dg2, x2), _wgrad32(dg2, hprev.view(B * T, Hs))], 1)
        if ctx.needs_input_grad[2]:
            db = K.colsum(dg2)
        return dx, dW, db


def separator(mod, x: torch.Tensor, relu_outputs=None):
    """`Separator.forward` (ref:models/separator.py:151-166) in fp32-class arithmetic, differentiable.  `relu_outputs`: optional
    list that receives every ReLU output in evaluation order (test hook: lets a float64 oracle use the same sub-gradient
    choice at pre-activations that are zero to within rounding)."""
    from . import ops
    import torch.nn as nn
    if mod.proj_activation not in ("relu", None):
        raise NotImplementedError("mtasr_b200 fp32 mode: Separator proj_activation must be 'relu' or None")
    act = K.ACT_RELU if mod.proj_activation == "relu" else K.ACT_NONE
    y = LinearF32Fn.apply(x.float(), mod.pre_proj.weight, mod.pre_proj.bias, act)
    if relu_outputs is not None and act == K.ACT_RELU:
        relu_outputs.append(y)
    y = ops.layer_norm(y, mod.pre_ln.weight, mod.pre_ln.bias, mod.pre_ln.eps, F32)
    for l, cell in enumerate(mod.lstm.cells):
        y = LSTMLayerF32Fn.apply(y, cell.W.weight, cell.W.bias)
        if mod.lstm.norms:
            n = mod.lstm.norms[l]
            y = ops.layer_norm(y, n.weight, n.bias, n.eps, F32)
        y = mod.lstm.dropout(y)
    y = ops.layer_norm(y, mod.post_ln.weight, mod.post_ln.bias, mod.post_ln.eps, F32)
    outs = []
    for br in mod.sep_branches:
        h = y
        mods = list(br)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                h = LinearF32Fn.apply(h, m.weight, m.bias, K.ACT_RELU if relu else K.ACT_NONE)
                if relu_outputs is not None and relu:
                    relu_outputs.append(h)
                i += 2 if relu else 1
            elif isinstance(m, nn.Dropout):
                h = m(h)
                i += 1
            elif isinstance(m, nn.LayerNorm):
                h = ops.layer_norm(h, m.weight, m.bias, m.eps, F32)
                i += 1
            else:
                raise NotImplementedError(type(m))
        outs.append(h)
    return outs


def _head_operands(hs: torch.Tensor, w: torch.Tensor):
    B, T, D = hs.shape
    hs2 = hs.contiguous().view(B * T, D).float()
    return hs2, _a(hs2), _b(w)                                                 # (B*T, R*D), (V, R*D)


def _head_lse_argmax(A, Bw, bias, rows, V, D, want_lse=True, want_argmax=False):
    nt = K.gemm_n_tiles(V)
    part = torch.empty(rows, nt, 4, device=A.device, dtype=F32)
    K.gemm(K.Operand(A, R * D), K.Operand(Bw, R * D), rows, V, R * D, None, bias=bias, mode=1, lse_part=part)
    return K.lse_finalize(part, rows, nt, want_lse=want_lse, want_argmax=want_argmax)


class CTCHeadF32Fn(torch.autograd.Function):
    """ops.CTCHeadFn in fp32-class arithmetic: ctc_lo -> log_softmax -> CTCLoss(reduction='none', zero_infinity=True)
    (ref:models/ctc.py:129-160, 51-65).  Forward: the vocabulary GEMM on split operands with the row log-sum-exp in its
    epilogue (no (B,T,V) tensor), lattice columns from gathered rows of the split weight, alpha recursion.  Backward: the
    fp32 logits ARE regenerated (rows x V fp32: this is the validation mode, not the memory-lean one), turned in place into
    softmax * upstream, the sparse occupancy term is scattered in, and dH / dW are two split-operand GEMMs over it."""

    @staticmethod
    def forward(ctx, hs, w, bias, hlens, ys, ylens, blank):
        B, T, D = hs.shape
        V = w.shape[0]
        hs2, A, Bw = _head_operands(hs, w)
        bf = bias.detach().float()
        lse, _ = _head_lse_argmax(A, Bw, bf, B * T, V, D)
        Lmax = int(ys.shape[1]) if ys.numel() else 0
        Lp = (Lmax + 1 + 63) // 64 * 64
        ys = ys.contiguous()
        hlens = hlens.to(torch.int64).contiguous()
        ylens = ylens.to(torch.int64).contiguous()
        wg, bg = K.ctc_gather_rows(Bw, bf, ys, ylens, Lp, blank)              # rows of the SPLIT weight: (B, Lp, R*D)
        glog = torch.empty(B, T, Lp, device=hs.device, dtype=F32)
        K.gemm(K.Operand(A, R * D, sb0=T * R * D), K.Operand(wg, R * D, sb0=Lp * R * D), T, Lp, R * D, K.Out(glog, Lp, sb0=T * Lp),
               batch=(B, 1), bias=bg, bias_sb0=Lp)
        lse = lse.view(B, T)
        nll, nll_raw, alpha, coff = K.ctc_alpha_fwd(glog, lse, ys, hlens, ylens, Lmax)
        ctx.dims = (B, T, D, V, Lp, Lmax, blank)
        ctx.hs_dtype = hs.dtype
        ctx.save_for_backward(hs2, w, bf, glog, lse, alpha, coff, nll_raw, hlens, ys, ylens)
        return nll

    @staticmethod
    def backward(ctx, gout):
        hs2, w, bf, glog, lse, alpha, coff, nll_raw, hlens, ys, ylens = ctx.saved_tensors
        B, T, D, V, Lp, Lmax, blank = ctx.dims
        dG, rowscale = K.ctc_beta_bwd(glog, lse, ys, hlens, ylens, Lmax, alpha, coff, nll_raw, gout.contiguous().float())
        Vp = (V + 7) // 8 * 8
        rows = B * T
        dl = torch.empty(rows, Vp, device=hs2.device, dtype=F32)
        K.gemm(K.Operand(_a(hs2), R * D), K.Operand(_b(w), R * D), rows, V, R * D, K.Out(dl, Vp), bias=bf)
        K.softmax_scale_f32_(dl, lse.view(-1), rowscale.view(-1), V)          # dense term, columns >= V zeroed
        K.ctc_scatter_cols(dG, ys, ylens, dl.view(B, T, Vp), blank)           # + sparse occupancy term (fp32 atomics)
        dh = dw = db = None
        if ctx.needs_input_grad[0]:
            wpad = w.detach().float()
            if Vp != V:
                wpad = torch.cat([wpad, wpad.new_zeros(Vp - V, D)], 0)
            dh = _dgrad32(dl, wpad).view(B, T, D).to(ctx.hs_dtype)
        if ctx.needs_input_grad[1]:
            dw = _wgrad32(dl, hs2)[:V]
        if ctx.needs_input_grad[2]:
            db = K.colsum(dl)[:V].contiguous()
        return dh, dw, db, None, None, None, None


def ctc_head_argmax(hs: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """argmax_v (hs W^T + b) at fp32-class accuracy, first maximal index on ties (ref:models/ctc.py:182-190)."""
    B, T, D = hs.shape
    _, A, Bw = _head_operands(hs, w)
    _, am = _head_lse_argmax(A, Bw, bias.detach().float(), B * T, w.shape[0], D, want_lse=False, want_argmax=True)
    return am.view(B, T)


def ctc_head_logits(hs: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    return LinearF32Fn.apply(hs, w, bias, K.ACT_NONE)
