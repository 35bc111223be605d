"""tcgen05 GEMM (csrc/gemm.cu) vs a plain PyTorch fp32 reference on the same bf16-rounded operands."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def _rand(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K,bn", [(128, 256, 64, 0), (256, 256, 128, 256), (300, 200, 136, 0), (128, 64, 512, 64),
                                      (1000, 96, 72, 128), (4096, 2048, 1024, 256), (129, 257, 1032, 0)])
def test_plain(cuda, M, N, K, bn):
    from mtasr_b200 import kernels as Kn
    a, b = _rand((M, K), cuda, seed=1), _rand((N, K), cuda, seed=2)
    c = torch.full((M, N), float("nan"), device=cuda, dtype=torch.float32)
    Kn.gemm(Kn.Operand(a, K), Kn.Operand(b, K), M, N, K, Kn.Out(c, N), block_n=bn)
    ref = a.float() @ b.float().t()
    assert torch.isfinite(c).all()
    assert _rel(c, ref) < 1e-5, _rel(c, ref)


@pytest.mark.parametrize("amaj,bmaj", [(0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (200, 136, 1000), (1024, 1024, 4096)])
def test_majors(cuda, amaj, bmaj, M, N, K):
    from mtasr_b200 import kernels as Kn
    a = _rand((K, M) if amaj else (M, K), cuda, seed=3)
    b = _rand((K, N) if bmaj else (N, K), cuda, seed=4)
    c = torch.full((M, N), float("nan"), device=cuda, dtype=torch.float32)
    Kn.gemm(Kn.Operand(a, a.stride(0), major=amaj), Kn.Operand(b, b.stride(0), major=bmaj), M, N, K, Kn.Out(c, N))
    A = a.float().t() if amaj else a.float()
    Bm = b.float().t() if bmaj else b.float()
    ref = A @ Bm.t()
    assert _rel(c, ref) < 1e-5, _rel(c, ref)


def test_epilogue_bias_gelu_residual_aux(cuda):
    from mtasr_b200 import kernels as Kn
    M, N, K = 300, 200, 136
    a, b = _rand((M, K), cuda, 0.5, 5), _rand((N, K), cuda, 0.5, 6)
    bias = torch.randn(N, device=cuda)
    res = torch.randn(M, N, device=cuda)
    y, aux = Kn.linear_fwd(a, b, bias, act=Kn.ACT_GELU, residual=res, out_dtype=torch.float32, want_aux=True)
    pre = a.float() @ b.float().t() + bias
    ref = torch.nn.functional.gelu(pre) + res
    assert _rel(y, ref) < 1e-5
    assert _rel(aux, pre) < 5e-3
    yb = Kn.linear_fwd(a, b, bias, act=Kn.ACT_RELU, out_dtype=torch.bfloat16)
    assert _rel(yb, torch.relu(pre)) < 5e-3
    # accumulate + alpha
    c0 = torch.randn(M, N, device=cuda)
    c = c0.clone()
    Kn.gemm(Kn.Operand(a, K), Kn.Operand(b, K), M, N, K, Kn.Out(c, N), alpha=0.25, accumulate=True)
    assert _rel(c, c0 + 0.25 * (a.float() @ b.float().t())) < 1e-5


def test_linear_grads(cuda):
    from mtasr_b200 import kernels as Kn
    M, N, K = 520, 192, 328
    x, w, dy = _rand((M, K), cuda, 1, 7), _rand((N, K), cuda, 1, 8), _rand((M, N), cuda, 1, 9)
    u = _rand((M, K), cuda, 1, 10)
    dx = Kn.linear_dgrad(dy, w, out_dtype=torch.float32)
    assert _rel(dx, dy.float() @ w.float()) < 1e-5
    dxg = Kn.linear_dgrad(dy, w, out_dtype=torch.float32, act=Kn.ACT_GELU_BWD, act_src=u)
    uf = u.float().requires_grad_(True)
    (gref,) = torch.autograd.grad(torch.nn.functional.gelu(uf), uf, dy.float() @ w.float())
    assert _rel(dxg, gref) < 1e-4
    dxr = Kn.linear_dgrad(dy, w, out_dtype=torch.float32, act=Kn.ACT_RELU_BWD, act_src=u)
    assert _rel(dxr, (dy.float() @ w.float()) * (u.float() > 0)) < 1e-5
    dw = Kn.linear_wgrad(dy, x)
    assert _rel(dw, dy.float().t() @ x.float()) < 1e-5


def test_batched_strided(cuda):
    """attention-shaped: Q,K,V packed (B,T,3,H,64); S = Q K^T per (b,h); O = P V written head-major into (B,T,D)."""
    from mtasr_b200 import kernels as Kn
    B, T, H, d = 2, 197, 3, 64
    D = H * d
    Tp = (T + 7) // 8 * 8
    qkv = _rand((B, T, 3 * D), cuda, 0.5, 11)
    S = torch.zeros(B, H, T, Tp, device=cuda)
    Kn.gemm(Kn.Operand(qkv, 3 * D, sb0=d, sb1=T * 3 * D, rows=T), Kn.Operand(qkv, 3 * D, sb0=d, sb1=T * 3 * D, offset=D, rows=T),
            T, T, d, Kn.Out(S, Tp, sb0=T * Tp, sb1=H * T * Tp), batch=(H, B))
    q = qkv.view(B, T, 3, H, d)[:, :, 0].permute(0, 2, 1, 3).float()
    k = qkv.view(B, T, 3, H, d)[:, :, 1].permute(0, 2, 1, 3).float()
    v = qkv.view(B, T, 3, H, d)[:, :, 2].permute(0, 2, 1, 3).float()
    assert _rel(S[..., :T], q @ k.transpose(-1, -2)) < 1e-5
    P = torch.zeros(B, H, T, Tp, device=cuda, dtype=torch.bfloat16)
    P[..., :T] = torch.softmax(S[..., :T] / 8, -1).to(torch.bfloat16)
    O = torch.zeros(B, T, D, device=cuda, dtype=torch.bfloat16)
    Kn.gemm(Kn.Operand(P, Tp, sb0=T * Tp, sb1=H * T * Tp), Kn.Operand(qkv, 3 * D, major=1, sb0=d, sb1=T * 3 * D, offset=2 * D, rows=T),
            T, d, T, Kn.Out(O, D, sb0=d, sb1=T * D), batch=(H, B))
    ref = (P[..., :T].float() @ v).permute(0, 2, 1, 3).reshape(B, T, D)
    assert _rel(O, ref) < 5e-3
    # dV = P^T dO  (both MN-major), written into a packed dqkv buffer
    dO = _rand((B, T, D), cuda, 0.5, 12)
    dqkv = torch.zeros(B, T, 3 * D, device=cuda, dtype=torch.bfloat16)
    Kn.gemm(Kn.Operand(P, Tp, major=1, sb0=T * Tp, sb1=H * T * Tp, rows=T), Kn.Operand(dO, D, major=1, sb0=d, sb1=T * D, rows=T),
            T, d, T, Kn.Out(dqkv, 3 * D, sb0=d, sb1=T * 3 * D, offset=2 * D), batch=(H, B))
    dOh = dO.view(B, T, H, d).permute(0, 2, 1, 3).float()
    refdv = (P[..., :T].float().transpose(-1, -2) @ dOh).permute(0, 2, 1, 3).reshape(B, T, D)
    assert _rel(dqkv.view(B, T, 3, D)[:, :, 2], refdv) < 5e-3


@pytest.mark.parametrize("k,s,C,Cout,L", [(3, 2, 64, 96, 401), (2, 2, 128, 64, 300), (3, 2, 512, 512, 1999)])
def test_implicit_conv(cuda, k, s, C, Cout, L):
    """conv1d(k, stride s, no padding) on channels-last x (B, L, C) as an implicit GEMM (hf:709-727 layers 1-6)."""
    from mtasr_b200 import kernels as Kn
    B = 2
    x = _rand((B, L, C), cuda, 0.5, 13)
    w = _rand((Cout, C, k), cuda, 0.2, 14)                      # torch conv layout (out, in, tap)
    wk = w.permute(0, 2, 1).contiguous().view(Cout, k * C)      # (out, tap*C + c)
    Lout = (L - k) // s + 1
    y = torch.zeros(B, Lout, Cout, device=cuda)
    Kn.gemm(Kn.Operand(x, s * C, sb1=L * C, inner=C, phase=s, rows=(L + s - 1) // s), Kn.Operand(wk, k * C),
            Lout, Cout, k * C, Kn.Out(y, Cout, sb1=Lout * Cout), batch=(1, B))
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), stride=s).transpose(1, 2)
    assert _rel(y, ref) < 1e-5


def test_grouped_pos_conv(cuda):
    """grouped conv k=16, pad=8, groups=2 on (B,T,D) with 64 channels per group (hf:48-90 shape, shrunk)."""
    from mtasr_b200 import kernels as Kn
    B, T, G, cg, k = 2, 150, 2, 64, 16
    D = G * cg
    x = _rand((B, T, D), cuda, 0.5, 15)
    w = _rand((D, cg, k), cuda, 0.1, 16)                         # torch grouped layout (out, in/groups, tap)
    bias = torch.randn(D, device=cuda)
    xp = Kn.pad_cast(x, k // 2, T + k)
    wk = w.view(G, cg, cg, k).permute(0, 1, 3, 2).contiguous()   # (g, out, tap, c)
    y = torch.zeros(B, T, D, device=cuda)
    Kn.gemm(Kn.Operand(xp, D, sb0=cg, sb1=(T + k) * D, inner=cg, phase=1, rows=T + k),
            Kn.Operand(wk, k * cg, sb0=cg * k * cg), T, cg, k * cg, Kn.Out(y, D, sb0=cg, sb1=T * D), batch=(G, B),
            bias=bias, bias_sb0=cg, act=Kn.ACT_GELU)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=k // 2, groups=G)[:, :, :T]
    ref = torch.nn.functional.gelu(ref).transpose(1, 2)
    assert _rel(y, ref) < 1e-5


@pytest.mark.parametrize("V", [1000, 5003])
def test_lse_and_exp_modes(cuda, V):
    from mtasr_b200 import kernels as Kn
    M, K = 333, 128
    h, w = _rand((M, K), cuda, 1.0, 17), _rand((V, K), cuda, 0.3, 18)
    bias = torch.randn(V, device=cuda) * 0.1
    nt = Kn.gemm_n_tiles(V)
    part = torch.empty(M, nt, 4, device=cuda)
    Kn.gemm(Kn.Operand(h, K), Kn.Operand(w, K), M, V, K, None, bias=bias, mode=1, lse_part=part)
    lse, am = Kn.lse_finalize(part, M, nt, want_argmax=True)
    logits = h.float() @ w.float().t() + bias
    assert (lse - torch.logsumexp(logits, -1)).abs().max() < 1e-4
    ref_am = logits.argmax(-1)
    top2 = logits.topk(2, -1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert (am[clear] == ref_am[clear]).all()
    scale = torch.rand(M, device=cuda)
    Pm = torch.empty(M, (V + 7) // 8 * 8, device=cuda, dtype=torch.bfloat16)
    Kn.gemm(Kn.Operand(h, K), Kn.Operand(w, K), M, V, K, Kn.Out(Pm, Pm.stride(0)), bias=bias, mode=2, row_vec=lse, row_scale=scale)
    ref = torch.softmax(logits, -1) * scale[:, None]
    assert _rel(Pm[:, :V], ref) < 5e-3
    # mode 1 with the optional fp16 logits output, and the streaming softmax built from it (same LSE partials as without)
    Vp = (V + 7) // 8 * 8
    lg16 = torch.full((M, Vp), float("nan"), device=cuda, dtype=torch.float16)
    part2 = torch.empty(M, nt, 4, device=cuda)
    Kn.gemm(Kn.Operand(h, K), Kn.Operand(w, K), M, V, K, Kn.Out(lg16, Vp), bias=bias, mode=1, lse_part=part2)
    lse2, am2 = Kn.lse_finalize(part2, M, nt, want_argmax=True)
    assert torch.equal(lse2, lse) and torch.equal(am2, am)
    assert (lg16[:, :V].float() - logits).abs().max() < 2e-3 * logits.abs().max()
    scale[5] = 0.0
    P2, cs = Kn.softmax_from_logits(lg16, lse, scale, V, want_colsum=True)
    ref = torch.softmax(logits, -1) * scale[:, None]
    assert _rel(P2[:, :V], ref) < 5e-3
    assert _rel(cs, ref.sum(0)) < 4e-3          # fp16 logits: 2^-11 relative on |logit| ~ 5 in the exponent
    assert P2[5].abs().max().item() == 0.0 and (Vp == V or P2[:, V:].abs().max().item() == 0.0)


@pytest.mark.parametrize("M,N,K", [(1024, 1024, 15968), (896, 128, 4000), (256, 192, 7000)])
def test_split_k_wgrad(cuda, M, N, K):
    """Weight-gradient shaped contractions with few output tiles are split along K (TMA reduce-add into a zeroed C)."""
    from mtasr_b200 import kernels as Kn
    dy, x = _rand((K, M), cuda, 1, 21), _rand((K, N), cuda, 1, 22)
    dw = torch.full((M, N + 8), float("nan"), device=cuda)[:, :N]          # strided C view (ld = N + 8)
    Kn.gemm(Kn.Operand(dy, M, major=1), Kn.Operand(x, N, major=1), M, N, K, Kn.Out(dw, dw.stride(0)))
    assert _rel(dw, dy.float().t() @ x.float()) < 1e-5


@pytest.mark.parametrize("M,N,K,amaj,bmaj", [(5120, 1024, 512, 0, 0),     # 80 pair tiles on 74 SM pairs: 6 tail tiles -> 12 half tiles
                                             (5120, 1024, 520, 0, 1),     # dgrad form (B MN-major), ragged K
                                             (5000, 1024, 512, 1, 1),     # wgrad form, ragged last M tile
                                             (384, 256 * 160, 256, 0, 0),  # single-CTA tiles (M < 512): 480 tiles on 148 SMs, 36 tail tiles
                                             (15968, 1024, 1024, 0, 0)])  # the cfg2 out-proj / FFN2 / dgrad shape (252 tiles = 3.4 rounds)
def test_tail_split_tiles(cuda, monkeypatch, M, N, K, amaj, bmaj):
    """The last, less-than-half-full round of tiles is cut into half-width tiles (gemm.cu decode_tile): results must be those
    of the un-split schedule, for every operand layout, with the fused bias + residual epilogue and for bf16 / fp32 outputs."""
    from mtasr_b200 import kernels as Kn
    a = _rand((K, M) if amaj else (M, K), cuda, 0.5, seed=7)
    b = _rand((K, N) if bmaj else (N, K), cuda, 0.5, seed=8)
    bias = torch.randn(N, device=cuda)
    res = torch.randn(M, N, device=cuda)
    A = a.float().t() if amaj else a.float()
    Bm = b.float().t() if bmaj else b.float()
    ref = A @ Bm.t() + bias + res

    def run(dtype):
        c = torch.full((M, N), float("nan"), device=cuda, dtype=dtype)
        Kn.gemm(Kn.Operand(a, a.stride(0), major=amaj), Kn.Operand(b, b.stride(0), major=bmaj), M, N, K, Kn.Out(c, N), bias=bias,
                residual=Kn.Out(res, N))
        return c

    c32, c16 = run(torch.float32), run(torch.bfloat16)
    assert torch.isfinite(c32).all() and _rel(c32, ref) < 1e-5, _rel(c32, ref)
    assert _rel(c16, ref) < 5e-3
    monkeypatch.setenv("MTASR_GEMM_NO_TAILSPLIT", "1")
    c32_ref = run(torch.float32)
    assert torch.equal(c32, c32_ref)            # same products, same accumulation order per output element
