"""torch.autograd.Functions over the C-ABI kernels.

Every Function keeps its saved tensors untouched in backward (safe under retain_graph=True: the reference's PCGrad
trainer back-propagates K+1 times per step, ref:src/trainer_seq2seq.py:1071-1141), holds no hidden global state
(safe under torch.utils.checkpoint) and honours ctx.needs_input_grad (frozen sub-modules,
ref:utils/unfreeze_utils.py:39-96).  Parameters stay ordinary fp32 nn.Parameters; bf16 operand copies are made
on the fly and cached per parameter version.
"""
import os
import weakref
from typing import Optional

import torch
from torch.autograd import Function

from . import kernels as K

BF = torch.bfloat16
F32 = torch.float32

_bf16_cache = {}
_UNFUSED_ATTN = os.environ.get("MTASR_UNFUSED_ATTN", "") not in ("", "0")   # debugging / A-B switch: materialise S and P
# one autograd node per half encoder layer (PreLNAttentionFn / PreLNFFNFn); "0" = the separate LayerNorm / gate / attention /
# FFN nodes (A-B switch; identical kernels)
_FUSED_LAYERS = os.environ.get("MTASR_FUSED_LAYERS", "1") not in ("", "0")
# CTC head: keep the fp16 logits of the forward vocabulary GEMM for the backward (2 B x B*T x V per head, 4.1 GB at cfg2)
# instead of regenerating the softmax with a second vocabulary GEMM.  "0" = memory-lean recompute path.
_CTC_KEEP_LOGITS = os.environ.get("MTASR_CTC_KEEP_LOGITS", "1") not in ("", "0")
# Row pitch of the (B*T, V) logits / softmax matrices of the CTC head, in elements.  A multiple of 64 two-byte elements makes
# every row start on a 128-byte line, so each 128-byte row of a TMA box (fp16 logits store in the forward, bf16 softmax
# operand loads in the two gradient GEMMs) is exactly one L2 line and four full sectors; with the minimal pitch (multiple of
# 8 = the 16-byte stride rule of tensor maps) V = 128259 gives rows 16 bytes off the line grid: five sectors per box row and
# partially written lines at every tile border.  Pad columns are never read (the tensor maps clip at V).
_VP_ALIGN = max(8, int(os.environ.get("MTASR_VP_ALIGN", "64")) // 8 * 8)


def _vocab_pitch(V: int) -> int:
    return (V + _VP_ALIGN - 1) // _VP_ALIGN * _VP_ALIGN


_CAPTURING = False      # inside graphs.GraphedTrainStep capture: parameter-derived operands must be produced BY graph nodes


class capturing:
    """Context manager: bypass the per-parameter-version operand caches (`bf16_of`, `cat_cached`).  A CUDA graph replay
    cannot observe that an optimizer step changed a weight, so while a step is being captured every cast / concatenation
    of a parameter is issued as a kernel of the graph and re-executed by each replay."""

    def __enter__(self):
        global _CAPTURING
        self._prev = _CAPTURING
        _CAPTURING = True
        return self

    def __exit__(self, *exc):
        global _CAPTURING
        _CAPTURING = self._prev
        return False


def bf16_of(p: torch.Tensor) -> torch.Tensor:
    """bf16 copy of a (parameter) tensor, cached on (identity, version) so frozen weights are converted once."""
    if p.dtype == BF:
        return p
    if _CAPTURING:
        return K.cast_bf16(p.detach())
    key = id(p)
    hit = _bf16_cache.get(key)
    if hit is not None:
        ref, ver, ptr, val = hit
        if ref() is p and ver == p._version and ptr == p.data_ptr():
            return val
    val = K.cast_bf16(p.detach())
    try:
        _bf16_cache[key] = (weakref.ref(p, lambda _r, k=key: _bf16_cache.pop(k, None)), p._version, p.data_ptr(), val)
    except TypeError:
        pass
    return val


_cat_cache = {}


def cat_cached(params, dtype):
    """torch.cat(params, 0) converted to `dtype`, cached on the identities / versions of the parameters (the fused
    QKV weight and bias are rebuilt only after an optimizer step changed them)."""
    if _CAPTURING:
        return torch.cat([p.detach().to(dtype) for p in params], 0).contiguous()
    key = tuple(id(p) for p in params) + (dtype,)
    sig = tuple((p._version, p.data_ptr()) for p in params)
    hit = _cat_cache.get(key)
    if hit is not None and hit[0] == sig and all(r() is p for r, p in zip(hit[1], params)):
        return hit[2]
    val = torch.cat([p.detach().to(dtype) for p in params], 0).contiguous()
    try:
        refs = tuple(weakref.ref(p) for p in params)
    except TypeError:
        return val
    _cat_cache[key] = (sig, refs, val)
    return val


# ------------------------------------------------------------------------------------------------------ dropout
# Training-mode dropout of the encoder (hf:291 activation_dropout after the FFN's GELU; hf:294 / 323 / 364 hidden_dropout on the
# FFN and attention outputs before the residual add; hf:407 / 483 on the encoder input after the positional conv; hf:217
# attention_dropout on the attention probabilities).  The reference trains with all three at 0.1
# (ref:utils/create_from_pretrained.py:209-212 zeroes only feat_proj_dropout / final_dropout / layerdrop).
# Masks come from a counter-based generator (csrc/common.cuh drop_hash) keyed by a per-forward seed tensor drawn from torch's
# CUDA generator and a static site number, so the backward -- and a gradient-checkpoint replay, which receives the same
# seed tensor -- regenerate them; nothing is stored.  Fused sites: GEMM epilogues and the attention kernels.
SITE_ENTRY, SITE_FEAT_PROJ = 1, 2
SITE_ATTN_OUT, SITE_FFN_ACT, SITE_FFN_OUT, SITE_ATTN_PROB = 0, 1, 2, 3


def site_id(layer: int, which: int) -> int:
    return 16 + 4 * int(layer) + which


class DropState:
    """Dropout configuration of one forward pass: seed (2,) int32 on the device + quantised keep probabilities."""
    __slots__ = ("seed", "k_hidden", "k_act", "k_attn")

    def __init__(self, device, p_hidden: float, p_act: float, p_attn: float):
        self.seed = torch.randint(0, 2 ** 31 - 1, (2,), device=device, dtype=torch.int32)     # consumes torch's CUDA generator
        self.k_hidden, self.k_act, self.k_attn = K.keep16(p_hidden), K.keep16(p_act), K.keep16(p_attn)

    def hidden(self, layer, which):
        return (self.seed, site_id(layer, which), self.k_hidden) if self.k_hidden < 65536 else None

    def act(self, layer):
        return (self.seed, site_id(layer, SITE_FFN_ACT), self.k_act) if self.k_act < 65536 else None

    def attn(self, layer):
        return (self.seed, site_id(layer, SITE_ATTN_PROB), self.k_attn) if self.k_attn < 65536 else None


class DropoutFn(Function):
    """Stand-alone dropout pass y = x * mask / keep (sites that are not a GEMM epilogue: the encoder input)."""

    @staticmethod
    def forward(ctx, x, seed, site, k16):
        ctx.site, ctx.k16 = site, k16
        ctx.save_for_backward(seed)
        return K.dropout(x, seed, site, k16)

    @staticmethod
    def backward(ctx, dy):
        (seed,) = ctx.saved_tensors
        return K.dropout(dy, seed, ctx.site, ctx.k16), None, None, None


def _masked_bf16(dy: torch.Tensor, drop) -> torch.Tensor:
    """bf16 (rows, D) GEMM operand of the gradient that flows back through a fused output-dropout site: dy * mask / keep."""
    return K.dropout(dy.view(-1, dy.shape[-1]), drop[0], drop[1], drop[2], out_dtype=BF)


def _flat2d(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous().view(-1, x.shape[-1])


# ------------------------------------------------------------------------------------------------------ LayerNorm
class LayerNormFn(Function):
    """hf:100-105, hf:314-366, hf:513, ref:models/separator.py:158-165 -- LayerNorm over the last dim."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, out_dtype):
        want_bf = out_dtype == BF
        yb, yf, mean, rstd = K.layernorm_fwd(x, gamma.detach().float(), beta.detach().float(), eps, out_bf16=want_bf, out_f32=not want_bf)
        ctx.save_for_backward(x, gamma, mean, rstd)
        return yb if want_bf else yf

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        need_x, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        x_f32 = x.dtype == F32
        dxf, dxb, dg, db = K.layernorm_bwd(dy, x, mean, rstd, gamma.detach().float(), want_f32=need_x and x_f32,
                                           want_bf16=need_x and not x_f32, want_param_grads=need_p)
        dx = (dxf if x_f32 else dxb) if need_x else None
        return dx, dg, db, None, None


def layer_norm(x, gamma, beta, eps, out_dtype=BF):
    return LayerNormFn.apply(x, gamma, beta, eps, out_dtype)


# ------------------------------------------------------------------------------------------------------ Linear
class LinearFn(Function):
    """y = act(x W^T + b) (+ residual) on the tcgen05 GEMM; backward = dgrad + wgrad GEMMs + bias column sum."""

    @staticmethod
    def forward(ctx, x, w, b, act, residual, out_dtype):
        shp = x.shape
        xb = K.cast_bf16(_flat2d(x))
        wb = bf16_of(w)
        res2 = None if residual is None else _flat2d(residual)
        want_aux = act != K.ACT_NONE
        out = K.linear_fwd(xb, wb, None if b is None else b.detach().float(), act=act, residual=res2, out_dtype=out_dtype,
                           want_aux=want_aux)
        y, aux = out if want_aux else (out, None)
        ctx.act = act
        ctx.x_dtype = x.dtype
        ctx.has_res = residual is not None
        ctx.has_bias = b is not None
        ctx.save_for_backward(xb, wb, aux)
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xb, wb, aux = ctx.saved_tensors
        dy2 = _flat2d(dy)
        du = K.act_bwd(dy2, aux, K.ACT_GELU_BWD if ctx.act == K.ACT_GELU else K.ACT_RELU_BWD) if ctx.act != K.ACT_NONE \
            else K.cast_bf16(dy2)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = K.linear_dgrad(du, wb, out_dtype=ctx.x_dtype).view(*dy.shape[:-1], wb.shape[1])
        if ctx.needs_input_grad[1]:
            dw = K.linear_wgrad(du, xb)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = K.colsum(du)
        dres = dy if (ctx.has_res and ctx.needs_input_grad[4]) else None
        return dx, dw, db, None, dres, None


def linear(x, w, b=None, act=K.ACT_NONE, residual=None, out_dtype=BF):
    return LinearFn.apply(x, w, b, act, residual, out_dtype)


# ------------------------------------------------------------------------------------------------------ FFN
class FFNFn(Function):
    """res + W2 GELU(W1 h + b1) + b2  (hf:274-295 with the residual add of hf:366/329 fused into the 2nd epilogue).
    The GELU backward is fused into the W2-dgrad epilogue (act 3)."""

    @staticmethod
    def forward(ctx, h, res, w1, b1, w2, b2, drop_act=None, drop_out=None):
        shp = h.shape
        hb = K.cast_bf16(_flat2d(h))
        w1b, w2b = bf16_of(w1), bf16_of(w2)
        a, u = K.linear_fwd(hb, w1b, b1.detach().float(), act=K.ACT_GELU, want_aux=True, drop=drop_act)
        y = K.linear_fwd(a, w2b, b2.detach().float(), residual=_flat2d(res), out_dtype=F32, drop=drop_out)
        ctx.h_dtype = h.dtype
        ctx.drops = (drop_act, drop_out)
        ctx.save_for_backward(hb, u, a, w1b, w2b)
        return y.view(shp[:-1] + (w2.shape[0],))

    @staticmethod
    def backward(ctx, dy):
        hb, u, a, w1b, w2b = ctx.saved_tensors
        drop_act, drop_out = ctx.drops
        dy = dy.contiguous()
        dyb = _masked_bf16(dy, drop_out) if drop_out is not None else K.cast_bf16(_flat2d(dy))
        du = K.linear_dgrad(dyb, w2b, act=K.ACT_GELU_BWD, act_src=u, drop=drop_act)
        dh = dw1 = db1 = dw2 = db2 = None
        if ctx.needs_input_grad[0]:
            dh = K.linear_dgrad(du, w1b, out_dtype=ctx.h_dtype).view(dy.shape[:-1] + (w1b.shape[1],))
        if ctx.needs_input_grad[2]:
            dw1 = K.linear_wgrad(du, hb)
        if ctx.needs_input_grad[3]:
            db1 = K.colsum(du)
        if ctx.needs_input_grad[4]:
            dw2 = K.linear_wgrad(dyb, a)
        if ctx.needs_input_grad[5]:
            db2 = K.colsum(dyb)
        return dh, (dy if ctx.needs_input_grad[1] else None), dw1, db1, dw2, db2, None, None


# ------------------------------------------------------------------------------------------------------ rel-pos gate
class RelPosGateFn(Function):
    """gru_rel_pos gate of hf:167-176 -> (B,H,T) fp32:  view(Linear_{64->8}(x_h), 2, 4).sum(-1) -> two sigmoids ->
    ga * (gb * const_h - 1) + 2.  The 4-row sums of the tiny (8,64) weight are formed on the parameter side; the per-frame
    work (and its backward into the layer input, the weight, the bias and gru_rel_pos_const) is one warp-per-frame kernel."""

    @staticmethod
    def forward(ctx, h, weight, bias, const):
        B, T, D = h.shape
        if weight.shape != (8, 64) or D % 64 != 0:
            raise NotImplementedError("mtasr_b200: gru_rel_pos gate expects head_dim 64 and an (8, 64) projection")
        H = D // 64
        w8, b8, cst = _gate_params(weight, bias, const, H)
        x = h.detach().contiguous()
        gate = K.relpos_gate_fwd(x, w8, b8, cst, B, T, H)
        ctx.dims = (B, T, H)
        ctx.save_for_backward(x, w8, b8, cst)
        ctx.const_shape = const.shape
        return gate

    @staticmethod
    def backward(ctx, dgate):
        x, w8, b8, cst = ctx.saved_tensors
        B, T, H = ctx.dims
        dx, dw8, db8, dcst = K.relpos_gate_bwd(x, w8, b8, cst, dgate.contiguous().float(), B, T, H)
        dh = dx.to(x.dtype) if ctx.needs_input_grad[0] else None
        dw = dw8 if ctx.needs_input_grad[1] else None
        db = db8 if ctx.needs_input_grad[2] else None
        dc = dcst.view(ctx.const_shape) if ctx.needs_input_grad[3] else None
        return dh, dw, db, dc


# ------------------------------------------------------------------------------------------------------ attention
class AttentionFn(Function):
    """res + out_proj(softmax(Q K^T / sqrt(d) + gate * relpos + key mask) V)   (hf:147-241, torch:6244-6695).

    h (B,T,D) is the attention input (bf16), gate (B,H,T) fp32 the gru_rel_pos gate (hf:167-176), table (H, 2T-1) the
    Toeplitz relative-position bias (hf:243-271) so bias[h,q,k] = table[h, k-q+T-1]; klen (B,) int32 valid key
    lengths or None.  QK^T and PV are batched tcgen05 GEMMs over (head, utterance); the gated bias, mask and softmax
    run in one row kernel that never materialises the (B*H,T,T) fp32 bias of the reference."""

    @staticmethod
    def forward(ctx, h, res, wq, bq, wk, bk, wv, bv, wo, bo, gate, table, klen, H, drop_out=None, drop_attn=None):
        B, T, D = h.shape
        d = D // H
        hb = K.cast_bf16(_flat2d(h))
        wqkv = cat_cached((wq, wk, wv), BF)
        bqkv = cat_cached((bq, bk, bv), F32)
        wob = bf16_of(wo)
        qkv = K.linear_fwd(hb, wqkv, bqkv)                                   # (B*T, 3D) bf16: [q | k | v], head-major
        scale = float(d) ** -0.5
        gate = gate.detach().contiguous().float()
        table = table.detach().contiguous().float()
        fused = d == 64 and not _UNFUSED_ATTN
        if fused:
            O, lse = K.attn_fwd(qkv, gate, table, klen, B, H, T, scale, drop=drop_attn)   # S / P tiles never leave TMEM / smem
            P = None
            Tp = 0
        else:
            if drop_attn is not None:
                raise NotImplementedError("mtasr_b200: attention_dropout > 0 in training needs the fused attention kernels "
                                          "(head_dim 64, MTASR_UNFUSED_ATTN unset)")
            Tp = (T + 7) // 8 * 8
            S = torch.empty(B, H, T, Tp, device=h.device, dtype=F32)
            K.gemm(K.Operand(qkv, 3 * D, sb0=d, sb1=T * 3 * D, rows=T), K.Operand(qkv, 3 * D, sb0=d, sb1=T * 3 * D, offset=D, rows=T),
                   T, T, d, K.Out(S, Tp, sb0=T * Tp, sb1=H * T * Tp), batch=(H, B))
            P = K.attn_softmax_fwd(S, gate, table, klen, B, H, T, Tp, scale)
            del S
            lse = None
            O = torch.empty(B * T, D, device=h.device, dtype=BF)
            K.gemm(K.Operand(P, Tp, sb0=T * Tp, sb1=H * T * Tp), K.Operand(qkv, 3 * D, major=1, sb0=d, sb1=T * 3 * D, offset=2 * D, rows=T),
                   T, d, T, K.Out(O, D, sb0=d, sb1=T * D), batch=(H, B))
        y = K.linear_fwd(O, wob, bo.detach().float(), residual=_flat2d(res), out_dtype=F32, drop=drop_out)
        ctx.dims = (B, T, D, H, Tp, scale)
        ctx.h_dtype = h.dtype
        ctx.fused = fused
        ctx.drops = (drop_out, drop_attn)
        ctx.save_for_backward(hb, qkv, P if P is not None else lse, O, wqkv, wob, gate, table, klen)
        return y.view(B, T, D)

    @staticmethod
    def backward(ctx, dy):
        hb, qkv, P, O, wqkv, wob, gate, table, klen = ctx.saved_tensors
        B, T, D, H, Tp, scale = ctx.dims
        d = D // H
        need = ctx.needs_input_grad
        drop_out, drop_attn = ctx.drops
        dy = dy.contiguous()
        dyb = _masked_bf16(dy, drop_out) if drop_out is not None else K.cast_bf16(_flat2d(dy))
        dO = K.linear_dgrad(dyb, wob)                                         # (B*T, D) bf16
        dwo = K.linear_wgrad(dyb, O) if need[8] else None
        dbo = K.colsum(dyb) if need[9] else None
        if ctx.fused:
            dqkv, dgate, dtable = K.attn_bwd(qkv, O, dO, P, gate, table, klen, B, H, T, scale, drop=drop_attn)   # P slot holds the row LSE
        else:
            dP = torch.empty(B, H, T, Tp, device=dy.device, dtype=F32)
            K.gemm(K.Operand(dO, D, sb0=d, sb1=T * D, rows=T), K.Operand(qkv, 3 * D, sb0=d, sb1=T * 3 * D, offset=2 * D, rows=T),
                   T, T, d, K.Out(dP, Tp, sb0=T * Tp, sb1=H * T * Tp), batch=(H, B))
            dqkv = torch.empty(B * T, 3 * D, device=dy.device, dtype=BF)
            # dV = P^T dO
            K.gemm(K.Operand(P, Tp, major=1, sb0=T * Tp, sb1=H * T * Tp, rows=T), K.Operand(dO, D, major=1, sb0=d, sb1=T * D, rows=T),
                   T, d, T, K.Out(dqkv, 3 * D, sb0=d, sb1=T * 3 * D, offset=2 * D), batch=(H, B))
            dS, dgate, dtable = K.attn_softmax_bwd(P, dP, gate, table, B, H, T, Tp, scale)
            del dP
            # dQ = dS K ; dK = dS^T Q   (dS already carries the 1/sqrt(d) factor)
            K.gemm(K.Operand(dS, Tp, sb0=T * Tp, sb1=H * T * Tp), K.Operand(qkv, 3 * D, major=1, sb0=d, sb1=T * 3 * D, offset=D, rows=T),
                   T, d, T, K.Out(dqkv, 3 * D, sb0=d, sb1=T * 3 * D, offset=0), batch=(H, B))
            K.gemm(K.Operand(dS, Tp, major=1, sb0=T * Tp, sb1=H * T * Tp, rows=T), K.Operand(qkv, 3 * D, major=1, sb0=d, sb1=T * 3 * D, offset=0, rows=T),
                   T, d, T, K.Out(dqkv, 3 * D, sb0=d, sb1=T * 3 * D, offset=D), batch=(H, B))
        dh = K.linear_dgrad(dqkv, wqkv, out_dtype=ctx.h_dtype).view(B, T, D) if need[0] else None
        dwq = dwk = dwv = dbq = dbk = dbv = None
        if need[2] or need[4] or need[6]:
            dwqkv = K.linear_wgrad(dqkv, hb)
            dwq, dwk, dwv = dwqkv[:D], dwqkv[D:2 * D], dwqkv[2 * D:]
        if need[3] or need[5] or need[7]:
            dbqkv = K.colsum(dqkv)
            dbq, dbk, dbv = dbqkv[:D], dbqkv[D:2 * D], dbqkv[2 * D:]
        return (dh, dy if need[1] else None, dwq, dbq, dwk, dbk, dwv, dbv, dwo, dbo,
                dgate if need[10] else None, dtable if need[11] else None, None, None, None, None)


# ------------------------------------------------------------------------------------------------------ fused pre-LN blocks
# The residual-stream gradient leaves one block's LayerNorm backward as fp32 and is immediately cast to bf16 by the next
# block's backward (operand of its dgrad / wgrad GEMMs).  The LayerNorm backward kernel can write both precisions in one
# pass, so the producer publishes the bf16 twin under the identity of the fp32 tensor it returns to autograd; the consumer
# uses it only if it receives that very tensor object, unmodified (autograd hands a single-consumer gradient through
# unchanged; anything else -- accumulation, hooks, checkpoint replays -- misses the registry and falls back to the cast).
_bf16_twins = {}


def _publish_twin(t32: torch.Tensor, tb: torch.Tensor, colsum: Optional[torch.Tensor] = None) -> None:
    if len(_bf16_twins) > 16:
        _bf16_twins.clear()
    try:
        _bf16_twins[id(t32)] = (weakref.ref(t32), t32._version, tb, colsum)
    except TypeError:
        pass


def _twin_and_sum(dy: torch.Tensor):
    """(bf16 (rows, D) operand copy of the fp32 gradient `dy`, its fp32 column sums or None).  The column sums are the bias
    gradient of the Linear that produced the tensor `dy` is the gradient of; the LayerNorm backward that emitted `dy`
    reduced them in the same pass (K.layernorm_bwd(want_dxsum=True))."""
    hit = _bf16_twins.pop(id(dy), None)
    if hit is not None and hit[0]() is dy and hit[1] == dy._version and hit[2].numel() == dy.numel():
        return hit[2].view(-1, dy.shape[-1]), hit[3]
    return K.cast_bf16(dy.view(-1, dy.shape[-1])), None


def _cast_or_twin(dy: torch.Tensor) -> torch.Tensor:
    """bf16 (rows, D) operand copy of the fp32 gradient `dy`."""
    return _twin_and_sum(dy)[0]


def _gate_params(weight, bias, const, H):
    """gru_rel_pos_linear (8,64) / its bias (8) / gru_rel_pos_const (H,) as contiguous fp32 views -- no arithmetic here: the
    4-row sums of hf:170-173 and their transposed expansion in the backward live inside the relpos_gate kernels (they were
    eight tiny torch launches per layer and step)."""
    return (weight.detach().float().contiguous(), bias.detach().float().contiguous(),
            const.detach().float().reshape(H).contiguous())


class PreLNAttentionFn(Function):
    """x + out_proj(attention(LN(x)))  -- the first half of a stable-layer-norm (WavLM-Large) encoder layer, hf:355-362,
    as ONE autograd node: LayerNorm, gru_rel_pos gate, fused QKV GEMM, fused attention, out-proj (+ residual epilogue).

    Same kernels as LayerNormFn -> RelPosGateFn -> AttentionFn; what the fusion removes is autograd's own arithmetic
    between them: the three gradient contributions to LN(x) (QKV dgrad, gate) are combined in the dgrad GEMM's residual
    epilogue and the residual-stream gradient is added inside the LayerNorm backward kernel (`dres`), instead of three
    65 MB elementwise adds / casts per layer launched by the autograd engine."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps, wq, bq, wk, bk, wv, bv, wo, bo, gw, gb, gconst, table, klen, H, drop_out=None,
                drop_attn=None):
        B, T, D = x.shape
        d = D // H
        if d != 64 or gw.shape != (8, 64):
            raise NotImplementedError("mtasr_b200: the fused pre-LN attention block needs head_dim 64")
        xf = x.contiguous()
        gamma = ln_w.detach().float()
        h1, _, mean, rstd = K.layernorm_fwd(xf, gamma, ln_b.detach().float(), eps, out_bf16=True, out_f32=False)
        wab, bab, cst = _gate_params(gw, gb, gconst, H)
        gate = K.relpos_gate_fwd(h1, wab, bab, cst, B, T, H)
        wqkv = cat_cached((wq, wk, wv), BF)
        bqkv = cat_cached((bq, bk, bv), F32)
        wob = bf16_of(wo)
        h2d = h1.view(B * T, D)
        qkv = K.linear_fwd(h2d, wqkv, bqkv)
        scale = float(d) ** -0.5
        tab = table.detach().contiguous().float()
        O, lse = K.attn_fwd(qkv, gate, tab, klen, B, H, T, scale, drop=drop_attn)
        y = K.linear_fwd(O, wob, bo.detach().float(), residual=xf.view(B * T, D), out_dtype=F32, drop=drop_out)
        ctx.drops = (drop_out, drop_attn)
        ctx.dims = (B, T, D, H, scale)
        ctx.const_shape = gconst.shape
        ctx.save_for_backward(xf, gamma, mean, rstd, h1, qkv, lse, O, wqkv, wob, gate, tab, klen, wab, bab, cst)
        return y.view(B, T, D)

    @staticmethod
    def backward(ctx, dy):
        xf, gamma, mean, rstd, h1, qkv, lse, O, wqkv, wob, gate, tab, klen, wab, bab, cst = ctx.saved_tensors
        B, T, D, H, scale = ctx.dims
        need = ctx.needs_input_grad
        drop_out, drop_attn = ctx.drops
        dy = dy.contiguous()
        dyb, dysum = (_masked_bf16(dy, drop_out), None) if drop_out is not None else _twin_and_sum(dy)
        dO = K.linear_dgrad(dyb, wob)
        dwo = K.linear_wgrad(dyb, O) if need[10] else None
        dbo = (dysum if dysum is not None else K.colsum(dyb)) if need[11] else None
        dqkv, dgate, dtable = K.attn_bwd(qkv, O, dO, lse, gate, tab, klen, B, H, T, scale, drop=drop_attn)
        # gate path into LN(x): only the two scalars per (frame, head); the LayerNorm backward expands da * wa + db * wb itself
        # the accumulators of this block's row kernels (gate kernel: 520 + H floats, LayerNorm backward: 3 x D) share ONE fill
        n_gate = (K.GATE_ACC + H + 3) // 4 * 4
        zbuf = torch.zeros(n_gate + 3 * D, device=dy.device, dtype=F32)
        dab, dwab, dbab, dcst = K.relpos_gate_bwd(h1, wab, bab, cst, dgate, B, T, H, want_dx=False, zeroed=zbuf)
        dx = dlnw = dlnb = None
        if need[0] or need[1] or need[2]:
            dh1 = K.linear_dgrad(dqkv, wqkv)
            dxf, dxb, dlnw, dlnb, dxs = K.layernorm_bwd(dh1.view(B, T, D), xf, mean, rstd, gamma, dres=dy, want_f32=True,
                                                        want_bf16=need[0], want_param_grads=need[1] or need[2],
                                                        want_dxsum=need[0], gate_ab=dab, gate_w8=wab, zeroed=zbuf[n_gate:])
            dx = dxf if need[0] else None
            if dx is not None:
                _publish_twin(dx, dxb, dxs)
        dwq = dwk = dwv = dbq = dbk = dbv = None
        if need[4] or need[6] or need[8]:
            dwqkv = K.linear_wgrad(dqkv, h1.view(B * T, D))
            dwq, dwk, dwv = dwqkv[:D], dwqkv[D:2 * D], dwqkv[2 * D:]
        if need[5] or need[7] or need[9]:
            dbqkv = K.colsum(dqkv)
            dbq, dbk, dbv = dbqkv[:D], dbqkv[D:2 * D], dbqkv[2 * D:]
        dgw = dwab if need[12] else None
        dgb = dbab if need[13] else None
        dgc = dcst.view(ctx.const_shape) if need[14] else None
        return (dx, dlnw, dlnb, None, dwq, dbq, dwk, dbk, dwv, dbv, dwo, dbo, dgw, dgb, dgc,
                dtable if need[15] else None, None, None, None, None)


class PreLNFFNFn(Function):
    """x + W2 GELU(W1 LN(x) + b1) + b2  -- the second half of a stable-layer-norm encoder layer (hf:363-366) as one
    autograd node (LayerNormFn -> FFNFn with the residual-stream gradient added inside the LayerNorm backward)."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps, w1, b1, w2, b2, drop_act=None, drop_out=None):
        shp = x.shape
        D = shp[-1]
        xf = x.contiguous()
        gamma = ln_w.detach().float()
        hb, _, mean, rstd = K.layernorm_fwd(xf, gamma, ln_b.detach().float(), eps, out_bf16=True, out_f32=False)
        w1b, w2b = bf16_of(w1), bf16_of(w2)
        h2d = hb.view(-1, D)
        a, u = K.linear_fwd(h2d, w1b, b1.detach().float(), act=K.ACT_GELU, want_aux=True, drop=drop_act)
        y = K.linear_fwd(a, w2b, b2.detach().float(), residual=xf.view(-1, D), out_dtype=F32, drop=drop_out)
        ctx.drops = (drop_act, drop_out)
        ctx.save_for_backward(xf, gamma, mean, rstd, hb, u, a, w1b, w2b)
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        xf, gamma, mean, rstd, hb, u, a, w1b, w2b = ctx.saved_tensors
        need = ctx.needs_input_grad
        D = xf.shape[-1]
        drop_act, drop_out = ctx.drops
        dy = dy.contiguous()
        dyb, dysum = (_masked_bf16(dy, drop_out), None) if drop_out is not None else _twin_and_sum(dy)
        du = K.linear_dgrad(dyb, w2b, act=K.ACT_GELU_BWD, act_src=u, drop=drop_act)
        dx = dlnw = dlnb = None
        if need[0] or need[1] or need[2]:
            dh = K.linear_dgrad(du, w1b)
            dxf, dxb, dlnw, dlnb, dxs = K.layernorm_bwd(dh.view(xf.shape), xf, mean, rstd, gamma, dres=dy, want_f32=True,
                                                        want_bf16=need[0], want_param_grads=need[1] or need[2],
                                                        want_dxsum=need[0])
            dx = dxf if need[0] else None
            if dx is not None:
                _publish_twin(dx, dxb, dxs)
        dw1 = K.linear_wgrad(du, hb.view(-1, D)) if need[4] else None
        db1 = K.colsum(du) if need[5] else None
        dw2 = K.linear_wgrad(dyb, a) if need[6] else None
        db2 = (dysum if dysum is not None else K.colsum(dyb)) if need[7] else None
        return dx, dlnw, dlnb, None, dw1, db1, dw2, db2, None, None


# ------------------------------------------------------------------------------------------------------ pos conv
def _group_pad(x_bf16: torch.Tensor, G: int, cgp: int) -> torch.Tensor:
    """(B,T,G*cg) -> (B,T,G*cgp) with every group's channels zero-padded to cgp (pure data movement; only needed when
    the group width is not a multiple of the 64-element TMA/UMMA K atom, e.g. WavLM-Base+: 768/16 = 48)."""
    B, T, D = x_bf16.shape
    cg = D // G
    if cgp == cg:
        return x_bf16
    out = K.empty_act((B, T, G * cgp), BF, x_bf16.device)
    o4 = out.view(B, T, G, cgp)
    o4[..., :cg] = x_bf16.view(B, T, G, cg)
    o4[..., cg:] = 0
    return out


class WeightNormFn(Function):
    """w = g * v / ||v|| with the norm over every dim but the last (torch weight_norm(dim=2) of the positional conv,
    hf:48-66): v = parametrizations.weight.original1 (out, in/groups, k), g = original0 (1, 1, k)."""

    @staticmethod
    def forward(ctx, v, g):
        vf, gf = v.detach().float().contiguous(), g.detach().float().reshape(-1).contiguous()
        w, sumsq = K.weightnorm_fwd(vf, gf)
        ctx.save_for_backward(vf, gf, sumsq)
        ctx.g_shape = g.shape
        return w

    @staticmethod
    def backward(ctx, dw):
        vf, gf, sumsq = ctx.saved_tensors
        dv, dg = K.weightnorm_bwd(dw.float(), vf, gf, sumsq)
        return (dv if ctx.needs_input_grad[0] else None), (dg.view(ctx.g_shape) if ctx.needs_input_grad[1] else None)


def pos_conv_weight(conv) -> torch.Tensor:
    """The (weight-normalised) kernel of the positional conv.  With the stock torch parametrization (weight_norm over the
    last dim, 256 % k == 0) the normalisation runs in our kernels; any other set-up reads `conv.weight` (torch computes it)."""
    pz = getattr(conv, "parametrizations", None)
    plist = getattr(pz, "weight", None) if pz is not None else None
    if plist is not None and len(plist) == 1 and type(plist[0]).__name__ == "_WeightNorm" and hasattr(plist, "original0"):
        v, g = plist.original1, plist.original0
        dim = plist[0].dim
        if v.is_cuda and v.dim() == 3 and dim in (2, -1) and g.numel() == v.shape[2] and 256 % v.shape[2] == 0:
            return WeightNormFn.apply(v, g)
    return conv.weight


class PosConvFn(Function):
    """x + GELU(grouped_conv1d(x, w, bias, k, pad=k//2)[:T])   (hf:48-90 + hf:403/481; `w` is the already
    weight-normalised kernel g*v/||v||, computed by torch so autograd carries the gradient to original0/original1).
    One implicit-GEMM launch over (group, utterance): A = taps x per-group channel slices of the zero-padded input."""

    @staticmethod
    def forward(ctx, x, w, bias, G, vlen):
        B, T, D = x.shape
        cg = D // G
        cgp = (cg + 63) // 64 * 64
        k = w.shape[2]
        pad = k // 2
        Tpad = T + k
        xp = K.pad_cast(x, pad, Tpad, vlen)                                  # (B, Tpad, D) bf16; rows >= vlen zeroed
        xg = _group_pad(xp, G, cgp)
        wk = torch.zeros(G, cg, k, cgp, device=x.device, dtype=BF)           # (g, out, tap, c) K-major
        wk[..., :cg] = w.detach().view(G, cg, cg, k).permute(0, 1, 3, 2)
        u = torch.empty(B, T, D, device=x.device, dtype=BF)
        y = torch.empty(B, T, D, device=x.device, dtype=F32)
        xres = x if vlen is None else (x * (torch.arange(T, device=x.device)[None, :] < vlen[:, None]).unsqueeze(-1))
        xres = xres.contiguous()
        K.gemm(K.Operand(xg, G * cgp, sb0=cgp, sb1=Tpad * G * cgp, inner=cgp, phase=1, rows=Tpad),
               K.Operand(wk, k * cgp, sb0=cg * k * cgp), T, cg, k * cgp, K.Out(y, D, sb0=cg, sb1=T * D), batch=(G, B),
               bias=bias.detach().float(), bias_sb0=cg, act=K.ACT_GELU, aux=u, residual=K.Out(xres, D, sb0=cg, sb1=T * D))
        ctx.dims = (B, T, D, G, k)
        ctx.save_for_backward(xp, u, w, vlen)
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, u, w, vlen = ctx.saved_tensors
        B, T, D, G, k = ctx.dims
        cg = D // G
        cgp = (cg + 63) // 64 * 64
        pad = k // 2
        Tpad = T + k
        du = K.act_bwd(dy.contiguous().view(B * T, D), u.view(B * T, D), K.ACT_GELU_BWD).view(B, T, D)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dx[t'] = sum_{o,tap} du[t' - tap + pad, o] w[o,c,tap]: conv of du (left pad k-1-pad) with flipped taps
            dup = _group_pad(K.pad_cast(du, k - 1 - pad, Tpad), G, cgp)
            wf = torch.zeros(G, cg, k, cgp, device=dy.device, dtype=BF)        # (g, c_in, tap', o)
            wf[..., :cg] = w.detach().view(G, cg, cg, k).flip(3).permute(0, 2, 3, 1)
            dxc = torch.empty(B, T, D, device=dy.device, dtype=F32)
            K.gemm(K.Operand(dup, G * cgp, sb0=cgp, sb1=Tpad * G * cgp, inner=cgp, phase=1, rows=Tpad),
                   K.Operand(wf, k * cgp, sb0=cg * k * cgp), T, cg, k * cgp, K.Out(dxc, D, sb0=cg, sb1=T * D), batch=(G, B),
                   residual=K.Out(dy.contiguous(), D, sb0=cg, sb1=T * D))
            if vlen is not None:
                dxc = dxc * (torch.arange(T, device=dy.device)[None, :] < vlen[:, None]).unsqueeze(-1)
            dx = dxc
        if ctx.needs_input_grad[1]:
            # dW[g,o,tap,c] = sum_{b,t} du[b,t,g*cg+o] xp[b,t+tap,g*cg+c]; rows flattened over (b, padded t)
            du2 = K.pad_cast(du, 0, Tpad)                                      # zero rows kill cross-utterance terms
            xpe = torch.zeros(B * Tpad + k, D, device=dy.device, dtype=BF)
            xpe[: B * Tpad] = xp.view(B * Tpad, D)
            dwk = torch.empty(G, cg, k, cg, device=dy.device, dtype=F32)       # (g, o, tap, c)
            if (cg * 2) % 16 == 0:                                             # one launch over (tap, group)
                K.gemm(K.Operand(du2, D, major=1, sb1=cg, rows=B * Tpad), K.Operand(xpe, D, major=1, sb0=D, sb1=cg, rows=B * Tpad),
                       cg, cg, B * Tpad, K.Out(dwk, k * cg, sb0=cg, sb1=cg * k * cg), batch=(k, G))
            else:                                                              # group offsets not 16-byte aligned: per group
                for g in range(G):
                    K.gemm(K.Operand(du2, D, major=1, offset=g * cg, rows=B * Tpad),
                           K.Operand(xpe, D, major=1, sb0=D, offset=g * cg, rows=B * Tpad), cg, cg, B * Tpad,
                           K.Out(dwk, k * cg, sb0=cg, offset=g * cg * k * cg), batch=(k, 1))
            dw = dwk.permute(0, 1, 3, 2).reshape(D, cg, k)
        if ctx.needs_input_grad[2]:
            db = K.colsum(du.view(B * T, D))
        return dx, dw, db, None, None


# ------------------------------------------------------------------------------------------------------ LSTM
class LSTMLayerFn(Function):
    """One layer of ref:models/separator.py:27-59 (CustomLSTMCell: gates = W [x_t, h_t] + b, order i,f,g,o).
    x (B,T,In) -> h (B,T,Hs) fp32.  Input half as one batched GEMM, recurrent half in the persistent kernel."""

    @staticmethod
    def forward(ctx, x, W, b):
        B, T, In = x.shape
        Hs = W.shape[0] // 4
        xb = K.cast_bf16(_flat2d(x))
        Wb = bf16_of(W)                                                        # (4Hs, In+Hs)
        ld = Wb.stride(0)
        xg = torch.empty(B * T, 4 * Hs, device=x.device, dtype=F32)
        K.gemm(K.Operand(xb, In), K.Operand(Wb, ld), B * T, 4 * Hs, In, K.Out(xg, 4 * Hs), bias=b.detach().float())
        whh = Wb[:, In:]
        outs = []
        for b0 in range(0, B, 64):
            outs.append(K.lstm_fwd(xg.view(B, T, 4 * Hs)[b0:b0 + 64].contiguous() if B > 64 else xg.view(B, T, 4 * Hs), whh, ld,
                                   want_h_f32=True))
        if len(outs) == 1:
            hb, hf, c, gates = outs[0]
        else:
            hb, hf, c, gates = (torch.cat([o[i] for o in outs], 0) for i in range(4))
        ctx.dims = (B, T, In, Hs)
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(xb, Wb, hb, c, gates)
        return hf

    @staticmethod
    def backward(ctx, dh):
        xb, Wb, hb, c, gates = ctx.saved_tensors
        B, T, In, Hs = ctx.dims
        ld = Wb.stride(0)
        whh = Wb[:, In:]
        dh = dh.contiguous().float()
        parts = []
        for b0 in range(0, B, 64):
            sl = slice(b0, b0 + 64)
            parts.append(K.lstm_bwd(dh[sl].contiguous(), gates[sl].contiguous(), c[sl].contiguous(), whh, ld))
        dg = parts[0] if len(parts) == 1 else torch.cat(parts, 0)              # (B,T,4Hs) bf16
        dg2 = dg.view(B * T, 4 * Hs)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(B * T, In, device=dh.device, dtype=ctx.x_dtype)
            K.gemm(K.Operand(dg2, 4 * Hs), K.Operand(Wb, ld, major=1), B * T, In, 4 * Hs, K.Out(dx, In))
            dx = dx.view(B, T, In)
        if ctx.needs_input_grad[1]:
            dW = torch.empty(4 * Hs, In + Hs, device=dh.device, dtype=F32)
            # input half: dgates^T x
            K.gemm(K.Operand(dg2, 4 * Hs, major=1), K.Operand(xb, In, major=1), 4 * Hs, In, B * T, K.Out(dW, In + Hs))
            # recurrent half: sum_{b,t} dgates[b,t]^T h[b,t-1] as ONE contraction over all B*T rows against the hidden
            # states shifted by one step (h[b,-1] = 0): a 28 MB copy instead of B launches of a K = T-1 GEMM each
            # (64 launches, 2.5 ms per cfg2 step in profiles/gemm_traffic_r2_by_shape.txt)
            hprev = torch.empty_like(hb)
            hprev[:, 0].zero_()
            if T > 1:
                hprev[:, 1:].copy_(hb[:, :-1])
            K.gemm(K.Operand(dg2, 4 * Hs, major=1), K.Operand(hprev.view(B * T, Hs), Hs, major=1), 4 * Hs, Hs, B * T,
                   K.Out(dW, In + Hs, offset=In))
        if ctx.needs_input_grad[2]:
            db = K.colsum(dg2)
        return dx, dW, db


# ------------------------------------------------------------------------------------------------------ CTC head
class CTCHeadFn(Function):
    """Per-utterance CTC negative log-likelihood of one head, fused with its vocabulary projection.

    Replaces ctc_lo -> log_softmax -> torch.nn.CTCLoss(reduction='none', zero_infinity=True, blank=V-1)
    (ref:models/ctc.py:129-160, 51-65).  The (B,T,V) logits / log-probs are never written: the vocab GEMM's epilogue
    produces per-tile (max, sum-exp) partials -> row LSE, the <= L+1 lattice columns of every utterance come from a
    small gathered GEMM, the alpha/beta recursions run one warp per utterance.  Backward regenerates softmax tiles
    (GEMM mode 2, scaled per row by the upstream gradient) for the dense term and adds the sparse occupancy term."""

    @staticmethod
    def forward(ctx, hs, w, bias, hlens, ys, ylens, blank):
        B, T, D = hs.shape
        V = w.shape[0]
        hb = K.cast_bf16(_flat2d(hs))
        wb = bf16_of(w)
        bf = bias.detach().float()
        nt = K.gemm_n_tiles(V)
        part = torch.empty(B * T, nt, 4, device=hs.device, dtype=F32)
        Vp = _vocab_pitch(V)
        keep = _CTC_KEEP_LOGITS and any(ctx.needs_input_grad[:3])
        logits16 = torch.empty(B * T, Vp, device=hs.device, dtype=torch.float16) if keep else None
        K.gemm(K.Operand(hb, D), K.Operand(wb, D), B * T, V, D, K.Out(logits16, Vp) if keep else None, bias=bf, mode=1, lse_part=part)
        lse, _ = K.lse_finalize(part, B * T, nt)
        del part
        Lmax = int(ys.shape[1]) if ys.numel() else 0
        Lp = (Lmax + 1 + 63) // 64 * 64
        ys = ys.contiguous()
        hlens = hlens.to(torch.int64).contiguous()
        ylens = ylens.to(torch.int64).contiguous()
        wg, bg = K.ctc_gather_rows(wb, bf, ys, ylens, Lp, blank)
        glog = torch.empty(B, T, Lp, device=hs.device, dtype=F32)
        K.gemm(K.Operand(hb, D, sb0=T * D), K.Operand(wg, D, sb0=Lp * D), T, Lp, D, K.Out(glog, Lp, sb0=T * Lp), batch=(B, 1),
               bias=bg, bias_sb0=Lp)
        lse = lse.view(B, T)
        nll, nll_raw, alpha, coff = K.ctc_alpha_fwd(glog, lse, ys, hlens, ylens, Lmax)
        ctx.dims = (B, T, D, V, Lp, Lmax, blank)
        ctx.hs_dtype = hs.dtype
        ctx.save_for_backward(hb, wb, bf, wg, glog, lse, alpha, coff, nll_raw, hlens, ys, ylens, logits16)
        return nll

    @staticmethod
    def backward(ctx, gout):
        hb, wb, bf, wg, glog, lse, alpha, coff, nll_raw, hlens, ys, ylens, logits16 = ctx.saved_tensors
        B, T, D, V, Lp, Lmax, blank = ctx.dims
        dev = gout.device
        dG, rowscale = K.ctc_beta_bwd(glog, lse, ys, hlens, ylens, Lmax, alpha, coff, nll_raw, gout.contiguous().float())
        dGb = K.cast_bf16(dG)
        Vp = _vocab_pitch(V)
        need_w, need_b = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        db_dense = None
        if logits16 is not None:                                               # softmax * upstream from the kept logits
            P, db_dense = K.softmax_from_logits(logits16, lse.view(-1), rowscale.view(-1), V, want_colsum=True)
        else:                                                                  # ... or regenerated by a second vocab GEMM
            P = torch.empty(B * T, Vp, device=dev, dtype=BF)
            K.gemm(K.Operand(hb, D), K.Operand(wb, D), B * T, V, D, K.Out(P, Vp), bias=bf, mode=2, row_vec=lse.view(-1),
                   row_scale=rowscale.view(-1))
        dh = dw = db = None
        if ctx.needs_input_grad[0]:
            dhf = torch.empty(B * T, D, device=dev, dtype=F32)
            K.gemm(K.Operand(P, Vp), K.Operand(wb, D, major=1), B * T, D, V, K.Out(dhf, D))
            K.gemm(K.Operand(dGb, Lp, sb0=T * Lp), K.Operand(wg, D, major=1, sb0=Lp * D), T, D, Lp, K.Out(dhf, D, sb0=T * D),
                   batch=(B, 1), accumulate=True)
            dh = dhf.view(B, T, D).to(ctx.hs_dtype)
        if need_w or need_b:
            dw = torch.empty(V, D, device=dev, dtype=F32)
            K.gemm(K.Operand(P, Vp, major=1), K.Operand(hb, D, major=1), V, D, B * T, K.Out(dw, D))
            dwg = torch.empty(B, Lp, D, device=dev, dtype=F32)
            K.gemm(K.Operand(dGb, Lp, major=1, sb0=T * Lp, rows=T), K.Operand(hb, D, major=1, sb0=T * D, rows=T), Lp, D, T,
                   K.Out(dwg, D, sb0=Lp * D), batch=(B, 1))
            db = db_dense if db_dense is not None else K.colsum(P)[:V].contiguous()
            K.ctc_scatter_rows(dwg, dG.sum(1).contiguous(), ys, ylens, blank, dw, db)
            if not need_w:
                dw = None
            if not need_b:
                db = None
        return dh, dw, db, None, None, None, None


def ctc_head_argmax(hs: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """argmax_v (hs W^T + b) without writing the logits (ref:models/ctc.py:182-190): GEMM mode 1 + finalize."""
    B, T, D = hs.shape
    V = w.shape[0]
    hb = K.cast_bf16(_flat2d(hs))
    wb = bf16_of(w)
    nt = K.gemm_n_tiles(V)
    part = torch.empty(B * T, nt, 4, device=hs.device, dtype=F32)
    K.gemm(K.Operand(hb, D), K.Operand(wb, D), B * T, V, D, None, bias=bias.detach().float(), mode=1, lse_part=part)
    _, am = K.lse_finalize(part, B * T, nt, want_lse=False, want_argmax=True)
    return am.view(B, T)


def ctc_head_path_and_pblank(hs: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, blank: int):
    """Greedy path (B,T) int64 and blank posterior (B,T) fp32 = exp(logit_blank - LSE) without the (B,T,V) tensor:
    one fused vocab GEMM (row LSE + argmax partials) plus the blank column from the gathered-row GEMM."""
    B, T, D = hs.shape
    V = w.shape[0]
    hb = K.cast_bf16(_flat2d(hs))
    wb = bf16_of(w)
    bf = bias.detach().float()
    nt = K.gemm_n_tiles(V)
    part = torch.empty(B * T, nt, 4, device=hs.device, dtype=F32)
    K.gemm(K.Operand(hb, D), K.Operand(wb, D), B * T, V, D, None, bias=bf, mode=1, lse_part=part)
    lse, am = K.lse_finalize(part, B * T, nt, want_lse=True, want_argmax=True)
    ys = torch.zeros(B, 1, device=hs.device, dtype=torch.int64)
    ylens = torch.zeros(B, device=hs.device, dtype=torch.int64)
    wg, bg = K.ctc_gather_rows(wb, bf, ys, ylens, 64, blank)               # column 0 of every lattice = blank
    glog = torch.empty(B, T, 64, device=hs.device, dtype=F32)
    K.gemm(K.Operand(hb, D, sb0=T * D), K.Operand(wg, D, sb0=64 * D), T, 64, D, K.Out(glog, 64, sb0=T * 64), batch=(B, 1),
           bias=bg, bias_sb0=64)
    pblank = torch.exp(glog[..., 0] - lse.view(B, T))
    return am.view(B, T), pblank.contiguous()


class SegmentMeanFn(Function):
    """Per-segment mean of frame features (ref:models/mt_ctctoken_builder.py:100-126) with gradient to the features."""

    @staticmethod
    def forward(ctx, x, pblank, ss, se, n, Lmax):
        xf = x.detach().float().contiguous()
        out, conf = K.segment_mean_fwd(xf, pblank, ss, se, n, Lmax)
        ctx.save_for_backward(ss, se, n)
        ctx.T = x.shape[1]
        ctx.x_dtype = x.dtype
        ctx.mark_non_differentiable(conf)
        return out.to(x.dtype), conf

    @staticmethod
    def backward(ctx, dout, _dconf):
        ss, se, n = ctx.saved_tensors
        dx = K.segment_mean_bwd(dout.contiguous().float(), ss, se, n, ctx.T)
        return dx.to(ctx.x_dtype), None, None, None, None, None


def ctc_head_logits(hs: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """Dense (B,T,V) fp32 logits for the API-compat methods CTC.logits/.softmax/.log_softmax (not on the hot path)."""
    return LinearFn.apply(hs, w, bias, K.ACT_NONE, None, F32)
