"""Fused tcgen05 attention (gated relative-position bias + key mask folded into the softmax tile) vs a plain PyTorch fp32
restatement of hf:147-271 (gate * Toeplitz bias added to QK^T / sqrt(d), -inf on padded keys, softmax, PV)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _ref(qkv, gate, table, klen, B, H, T, scale):
    D = H * 64
    x = qkv.float().view(B, T, 3, H, 64)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))            # (B,H,T,64)
    idx = (torch.arange(T, device=qkv.device)[None, :] - torch.arange(T, device=qkv.device)[:, None]) + T - 1
    z = q @ k.transpose(-1, -2) * scale + gate[..., None] * table[:, idx][None]
    kmask = torch.arange(T, device=qkv.device)[None, :] >= klen[:, None]
    z = z.masked_fill(kmask[:, None, None, :], float("-inf"))
    p = torch.softmax(z, -1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * T, D)
    return o, torch.logsumexp(z, -1), (q, k, v, z, p)


def _poison(cuda):
    """Fill recycled allocator blocks with NaN so that any read of an unwritten output / scratch element shows up."""
    junk = torch.full((32 * 1024 * 1024,), float("nan"), device=cuda)
    del junk


def _inputs(B, H, T, cuda, seed=0, ragged=True):
    g = torch.Generator(device=cuda).manual_seed(seed)
    qkv = (torch.randn(B * T, 3 * H * 64, device=cuda, generator=g) * 0.8).to(torch.bfloat16)
    gate = torch.rand(B, H, T, device=cuda, generator=g) * 2
    table = torch.randn(H, 2 * T - 1, device=cuda, generator=g)
    klen = torch.full((B,), T, device=cuda, dtype=torch.int32)
    if ragged and B > 1:
        klen[1:] = max(1, T - 37)
    return qkv, gate, table, klen


@pytest.mark.parametrize("B,H,T", [(2, 2, 64), (2, 3, 197), (1, 2, 499), (3, 16, 499), (1, 2, 1100), (2, 1, 128), (1, 1, 129),
                                   (4, 16, 300),     # three key tiles (odd): S/P buffer parity restarts per item, several items per CTA
                                   (2, 16, 512), (2, 4, 385), (2, 16, 513)])   # full last tile / one key in the last tile / first two-pass length
def test_attn_fwd(cuda, B, H, T):
    from mtasr_b200 import kernels as Kn
    qkv, gate, table, klen = _inputs(B, H, T, cuda)
    _poison(cuda)
    out, lse = Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125)
    ref, ref_lse, _ = _ref(qkv, gate, table, klen.long(), B, H, T, 0.125)
    assert _rel(out, ref) < 6e-3, _rel(out, ref)
    assert (lse - ref_lse).abs().max().item() < 2e-3
    out2, _ = Kn.attn_fwd(qkv, gate, table, None, B, H, T, 0.125)
    ref2, _, _ = _ref(qkv, gate, table, torch.full((B,), T, device=cuda), B, H, T, 0.125)
    assert _rel(out2, ref2) < 6e-3


@pytest.mark.parametrize("B,H,T", [(2, 2, 64), (2, 3, 197), (2, 16, 499), (1, 2, 749), (1, 1, 129), (4, 16, 300), (2, 4, 385)])
def test_attn_bwd(cuda, B, H, T):
    """dq/dk/dv, dgate and the Toeplitz table gradient vs torch autograd through the fp32 restatement."""
    from mtasr_b200 import kernels as Kn
    qkv, gate, table, klen = _inputs(B, H, T, cuda, seed=1)
    D = H * 64
    out, lse = Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125)
    g = torch.Generator(device=cuda).manual_seed(5)
    dout = (torch.randn(B * T, D, device=cuda, generator=g) * 0.5).to(torch.bfloat16)
    _poison(cuda)
    dqkv, dgate, dtable = Kn.attn_bwd(qkv, out, dout, lse, gate, table, klen, B, H, T, 0.125)
    qr = qkv.float().requires_grad_(True)
    gr = gate.clone().requires_grad_(True)
    tr = table.clone().requires_grad_(True)
    ref, _, _ = _ref(qr, gr, tr, klen.long(), B, H, T, 0.125)
    gq, gg, gt = torch.autograd.grad((ref * dout.float()).sum(), [qr, gr, tr])
    gq = gq.view(B, T, 3, D)
    mine = dqkv.float().view(B, T, 3, D)
    errs = [_rel(mine[:, :, i], gq[:, :, i]) for i in range(3)]
    assert max(errs) < 1.5e-2, errs
    assert _rel(dgate, gg) < 1.5e-2, _rel(dgate, gg)
    assert _rel(dtable, gt) < 1.5e-2, _rel(dtable, gt)
    # padded keys receive no gradient
    for b in range(B):
        kl = int(klen[b])
        if kl < T:
            assert mine[b, kl:, 1:].abs().max().item() == 0.0
