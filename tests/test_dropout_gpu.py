"""Training-mode dropout of the encoder (hf:217 attention_dropout, hf:291 activation_dropout, hf:294/323/364/407/483
hidden_dropout; the reference trains with 0.1 each, ref:utils/create_from_pretrained.py:209-212).

The reference's Philox stream cannot be reproduced, so parity is asserted in two layers:
  * SEMANTICS, exactly: the mask of a site is a pure function of (seed, site, element index) that the stand-alone kernel
    `mtasr_dropout` exposes; every fused site (GEMM epilogues, attention forward / backward) is compared with a plain PyTorch
    fp32 restatement of the reference arithmetic that applies THAT mask where the reference applies nn.Dropout / F.dropout;
  * STATISTICS: keep rate, scaling (unbiasedness), independence between sites / seeds, identical masks on a gradient-
    checkpoint replay and in the backward, eval mode untouched.
"""
import math

import pytest
import torch

from test_attention_gpu import _inputs, _rel

pytestmark = pytest.mark.gpu


def _seed(cuda, s=0):
    g = torch.Generator(device=cuda).manual_seed(1000 + s)
    return torch.randint(0, 2 ** 31 - 1, (2,), device=cuda, dtype=torch.int32, generator=g)


def _mask(rows, cols, seed, site, k16, cuda):
    """mask * 65536 / k16 of the (rows, cols) index space of a site."""
    from mtasr_b200 import kernels as Kn
    return Kn.dropout(torch.ones(rows, cols, device=cuda), seed, site, k16)


@pytest.mark.parametrize("p", [0.1, 0.5, 0.05])
def test_dropout_kernel_statistics(cuda, p):
    from mtasr_b200 import kernels as Kn
    k16 = Kn.keep16(p)
    keep = k16 / 65536.0
    rows, cols = 4096, 1023                                       # odd column count: row pitch rounds up to even
    m = _mask(rows, cols, _seed(cuda), 17, k16, cuda)
    vals = torch.unique(m)
    assert vals.numel() == 2 and vals[0].item() == 0.0 and abs(vals[1].item() - 1.0 / keep) < 1e-6
    n = rows * cols
    rate = (m > 0).float().mean().item()
    assert abs(rate - keep) < 5 * math.sqrt(keep * (1 - keep) / n), (rate, keep)
    assert abs(m.mean().item() - 1.0) < 5 * math.sqrt((1 - keep) / keep / n)          # unbiased
    kb = (m > 0).float()
    # neighbours share one 32-bit hash (two 16-bit halves): still uncorrelated; rows and columns too
    for a, b in ((kb[:, :-1], kb[:, 1:]), (kb[:-1], kb[1:]), (kb[:, ::2][:, :511], kb[:, 1::2][:, :511])):
        c = ((a - keep) * (b - keep)).mean().item() / (keep * (1 - keep))
        assert abs(c) < 5 / math.sqrt(a.numel()), c
    per_row = kb.mean(1)
    assert (per_row - keep).abs().max().item() < 6 * math.sqrt(keep * (1 - keep) / cols)
    # pure function of (seed, site, index); other site / seed: independent
    assert torch.equal(m, _mask(rows, cols, _seed(cuda), 17, k16, cuda))
    for other in (_mask(rows, cols, _seed(cuda), 18, k16, cuda), _mask(rows, cols, _seed(cuda, 1), 17, k16, cuda)):
        both = ((other > 0) & (m > 0)).float().mean().item()
        assert abs(both - keep * keep) < 5 * math.sqrt(keep * keep * (1 - keep * keep) / n)
    # bf16 in / out and the backward use the same mask
    x = torch.randn(rows, cols, device=cuda)
    y = Kn.dropout(x.to(torch.bfloat16), _seed(cuda), 17, k16, out_dtype=torch.float32)
    assert torch.equal(y, x.to(torch.bfloat16).float() * m)


@pytest.mark.parametrize("M,N,K", [(777, 1024, 512), (300, 4096, 1024), (130, 96, 128)])
def test_gemm_epilogue_dropout_sites(cuda, M, N, K):
    """The three fused GEMM sites against torch with the exported mask: (a) hidden dropout before the residual add (out-proj,
    FFN2), (b) activation dropout after GELU with the pre-activation tap untouched (FFN1), (c) its backward inside the
    GELU-backward dgrad epilogue (dropout o gelu')."""
    from mtasr_b200 import kernels as Kn
    g = torch.Generator(device=cuda).manual_seed(3)
    x = (torch.randn(M, K, device=cuda, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=cuda, generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device=cuda, generator=g)
    res = torch.randn(M, N, device=cuda, generator=g)
    seed, site, k16 = _seed(cuda, 2), 21, Kn.keep16(0.1)
    m = _mask(M, N, seed, site, k16, cuda)
    pre = x.float() @ w.float().t() + b
    # (a)
    y = Kn.linear_fwd(x, w, b, residual=res, out_dtype=torch.float32, drop=(seed, site, k16))
    assert _rel(y, pre * m + res) < 2e-3
    assert _rel(y - res, pre * m) < 5e-3
    # (b)
    a, u = Kn.linear_fwd(x, w, b, act=Kn.ACT_GELU, want_aux=True, drop=(seed, site, k16))
    assert _rel(u, pre) < 5e-3
    assert _rel(a, torch.nn.functional.gelu(pre) * m) < 8e-3
    assert ((a.float() == 0) | (m > 0)).all()                      # dropped elements are exactly zero
    # (c) du = (dy W2) o mask o gelu'(u): contraction over N2 of a second layer whose output gradient is dy
    N2 = 256
    dy = (torch.randn(M, N2, device=cuda, generator=g) * 0.5).to(torch.bfloat16)
    w2 = (torch.randn(N2, N, device=cuda, generator=g) * 0.05).to(torch.bfloat16)
    du = Kn.linear_dgrad(dy, w2, act=Kn.ACT_GELU_BWD, act_src=u, drop=(seed, site, k16))
    uf = u.float().requires_grad_(True)
    (ref,) = torch.autograd.grad((torch.nn.functional.gelu(uf) * m * (dy.float() @ w2.float())).sum(), [uf])
    assert _rel(du, ref) < 1e-2
    # no dropout: unchanged
    y0 = Kn.linear_fwd(x, w, b, residual=res, out_dtype=torch.float32)
    assert _rel(y0, pre + res) < 2e-3


def _attn_ref_drop(qkv, gate, table, klen, B, H, T, scale, mask):
    """hf:206-228 with F.dropout on the probabilities replaced by the given (B*H*T, T) multiplier."""
    D = H * 64
    x = qkv.view(B, T, 3, H, 64)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    idx = (torch.arange(T, device=qkv.device)[None, :] - torch.arange(T, device=qkv.device)[:, None]) + T - 1
    z = q @ k.transpose(-1, -2) * scale + gate[..., None] * table[:, idx][None]
    kmask = torch.arange(T, device=qkv.device)[None, :] >= klen[:, None]
    z = z.masked_fill(kmask[:, None, None, :], float("-inf"))
    p = torch.softmax(z, -1) * mask.view(B, H, T, T)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * T, D)


@pytest.mark.parametrize("B,H,T", [(2, 2, 64), (2, 3, 197), (2, 16, 499), (1, 2, 749), (1, 1, 129), (2, 4, 385)])
def test_attention_probability_dropout_fwd_bwd(cuda, B, H, T):
    """Fused attention with dropout on the probabilities (single-pass kernel for T <= 512, two-pass beyond; odd T exercises
    the even row pitch of the index space), forward and every gradient, against torch autograd with the exported mask."""
    from mtasr_b200 import kernels as Kn
    qkv, gate, table, klen = _inputs(B, H, T, cuda, seed=4)
    seed, site, k16 = _seed(cuda, 5), 35, Kn.keep16(0.1)
    drop = (seed, site, k16)
    mask = _mask(B * H * T, T, seed, site, k16, cuda)
    D = H * 64
    out, lse = Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125, drop=drop)
    out0, lse0 = Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125)
    assert torch.equal(lse, lse0)                                  # the softmax normaliser ignores the dropout
    qr = qkv.float().requires_grad_(True)
    gr = gate.clone().requires_grad_(True)
    tr = table.clone().requires_grad_(True)
    ref = _attn_ref_drop(qr, gr, tr, klen.long(), B, H, T, 0.125, mask)
    assert _rel(out, ref) < 8e-3, _rel(out, ref)
    assert _rel(out, out0) > 0.1                                   # the mask does something
    g = torch.Generator(device=cuda).manual_seed(5)
    dout = (torch.randn(B * T, D, device=cuda, generator=g) * 0.5).to(torch.bfloat16)
    dqkv, dgate, dtable = Kn.attn_bwd(qkv, out, dout, lse, gate, table, klen, B, H, T, 0.125, drop=drop)
    gq, gg, gt = torch.autograd.grad((ref * dout.float()).sum(), [qr, gr, tr])
    gq = gq.view(B, T, 3, D)
    mine = dqkv.float().view(B, T, 3, D)
    errs = [_rel(mine[:, :, i], gq[:, :, i]) for i in range(3)]
    assert max(errs) < 2e-2, errs
    assert _rel(dgate, gg) < 2e-2 and _rel(dtable, gt) < 2e-2, (_rel(dgate, gg), _rel(dtable, gt))


def _tiny_model(cuda, p=0.1, kind="tiny_large"):
    from _util import build_ours, load_model_golden
    from oracle.model_ref import make_config
    g, params, _ = load_model_golden(kind)
    cfg = make_config(kind, hidden_dropout=p, activation_dropout=p, attention_dropout=p)
    enc, sep, heads, loss_mod = build_ours(cfg, int(g["n_spk"]), int(g["hidden_sep"]), int(g["vocab"]), params)
    wav, mask = torch.from_numpy(g["wav"]).to(cuda), torch.from_numpy(g["mask"]).to(cuda)
    return enc, wav, mask


@pytest.mark.parametrize("kind", ["tiny_large", "tiny_base"])
def test_encoder_training_mode_dropout(cuda, kind):
    """Model level: eval mode is untouched; training mode draws its seed from torch's CUDA generator (same torch seed ->
    same output, other seed -> other output); the expectation over masks stays near the eval output; gradient
    checkpointing replays identical masks (ref:run.sh:239 trains with --gradient_checkpointing)."""
    enc, wav, mask = _tiny_model(cuda, 0.1, kind)
    enc0, _, _ = _tiny_model(cuda, 0.0, kind)
    with torch.no_grad():
        ev = enc(wav, attention_mask=mask)[1]
        assert torch.equal(ev, enc0(wav, attention_mask=mask)[1])
        enc.train()
        torch.manual_seed(11)
        a = enc(wav, attention_mask=mask)[1]
        torch.manual_seed(11)
        b = enc(wav, attention_mask=mask)[1]
        torch.manual_seed(12)
        c = enc(wav, attention_mask=mask)[1]
        assert torch.equal(a, b) and not torch.equal(a, c)
        fm = enc._get_feature_vector_attention_mask_x0(ev.shape[1], mask)
        d1 = _rel(a[fm], ev[fm])
        assert 0.02 < d1 < 1.0, d1
        acc = torch.zeros_like(ev)
        n = 24
        for s in range(n):
            torch.manual_seed(100 + s)
            acc += enc(wav, attention_mask=mask)[1]
        dn = _rel((acc / n)[fm], ev[fm])
        assert dn < 0.6 * d1, (dn, d1)                             # averaging over masks moves back towards the eval output
    # backward: checkpointed (recomputed) layers see the same masks as the plain run
    params = [p for p in enc.parameters() if p.requires_grad]
    torch.manual_seed(21)
    out = enc(wav, attention_mask=mask)[1]
    g_plain = torch.autograd.grad(out[fm].square().sum(), params, allow_unused=True)
    enc.gradient_checkpointing_enable()
    torch.manual_seed(21)
    out2 = enc(wav, attention_mask=mask)[1]
    g_ckpt = torch.autograd.grad(out2[fm].square().sum(), params, allow_unused=True)
    assert torch.equal(out, out2)
    for x, y in zip(g_plain, g_ckpt):
        if x is not None and x.numel() >= 256:
            assert _rel(y, x) < 1e-4
    enc.eval()


def test_layerdrop_in_training_raises(cuda):
    from _util import build_ours
    from oracle.model_ref import make_config
    cfg = make_config("tiny_large", layerdrop=0.1)
    enc, _, _, _ = build_ours(cfg, 2, 96, 33)
    enc.train()
    with pytest.raises(NotImplementedError):
        enc(torch.randn(1, 8000, device=cuda))
