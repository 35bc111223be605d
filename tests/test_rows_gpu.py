"""HBM-bound row kernels vs plain PyTorch fp32 references."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("D,rows", [(512, 777), (896, 777), (1024, 777), (128, 777), (64, 777), (768, 301), (8, 5), (1024, 1), (1024, 15968)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(cuda, D, rows, dt):
    from mtasr_b200 import kernels as Kn
    torch.manual_seed(0)
    x = (torch.randn(rows, D, device=cuda) * 2 + 0.5).to(dt)
    gamma, beta = torch.randn(D, device=cuda), torch.randn(D, device=cuda)
    yb, yf, mean, rstd = Kn.layernorm_fwd(x, gamma, beta, 1e-5, out_bf16=True, out_f32=True)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), gr, br, 1e-5)
    assert _rel(yf, ref) < 1e-5 and _rel(yb, ref) < 5e-3
    dy = torch.randn(rows, D, device=cuda)
    dres = torch.randn(rows, D, device=cuda)
    dxf, dxb, dg, db, dxs = Kn.layernorm_bwd(dy, x, mean, rstd, gamma, dres=dres, want_bf16=True, want_dxsum=True)
    ref.backward(dy)
    assert _rel(dxf, xr.grad + dres) < 1e-5 and _rel(dxb, xr.grad + dres) < 5e-3
    assert _rel(dg, gr.grad) < 1e-4 and _rel(db, br.grad) < 1e-4
    assert _rel(dxs, (xr.grad + dres).sum(0)) < 1e-4        # column sums of dx: the neighbouring Linear's bias gradient
    # the separate entry points of the same pass: no dres / no sums, bf16 upstream gradient, parameter gradients only
    o = Kn.layernorm_bwd(dy, x, mean, rstd, gamma, want_param_grads=False)
    assert _rel(o[0], xr.grad) < 1e-5 and o[2] is None and o[3] is None
    dyb = dy.to(torch.bfloat16)
    xr.grad = None; gr.grad = None; br.grad = None
    F.layer_norm(xr, (D,), gr, br, 1e-5).backward(dyb.float())
    o = Kn.layernorm_bwd(dyb, x, mean, rstd, gamma, want_f32=False, want_bf16=True, want_dxsum=True)
    assert o[0] is None and _rel(o[1], xr.grad) < 5e-3 and _rel(o[2], gr.grad) < 1e-4 and _rel(o[3], br.grad) < 1e-4
    assert _rel(o[4], xr.grad.sum(0)) < 1e-3
    o = Kn.layernorm_bwd(dyb, x, mean, rstd, gamma, want_f32=False, want_bf16=False)
    assert o[0] is None and o[1] is None and _rel(o[2], gr.grad) < 1e-4 and _rel(o[3], br.grad) < 1e-4
    g2, _, _, _ = Kn.layernorm_fwd(x, gamma, beta, 1e-5, post_gelu=True)
    assert _rel(g2, F.gelu(ref.detach())) < 5e-3


def test_cast_colsum_pad_glu(cuda):
    from mtasr_b200 import kernels as Kn
    torch.manual_seed(1)
    x = torch.randn(1003, device=cuda)
    assert torch.equal(Kn.cast_bf16(x), x.to(torch.bfloat16))
    m = torch.randn(999, 200, device=cuda)
    assert _rel(Kn.colsum(m), m.sum(0)) < 1e-5
    mb = m.to(torch.bfloat16)
    assert _rel(Kn.colsum(mb), mb.float().sum(0)) < 1e-5
    a = torch.randn(2, 50, 64, device=cuda)
    vlen = torch.tensor([50, 31], device=cuda, dtype=torch.int32)
    p = Kn.pad_cast(a, 8, 66, vlen)
    ref = torch.zeros(2, 66, 64, device=cuda)
    ref[0, 8:58] = a[0]; ref[1, 8:39] = a[1, :31]
    assert torch.equal(p, ref.to(torch.bfloat16))
    z = torch.randn(3, 40, 256, device=cuda)
    yb, yf = Kn.glu_fwd(z, out_f32=True)
    assert _rel(yf, F.glu(z, -1)) < 1e-6
    zr = z.clone().requires_grad_(True)
    dy = torch.randn(3, 40, 128, device=cuda)
    F.glu(zr, -1).backward(dy)
    assert _rel(Kn.glu_bwd(z, dy), zr.grad) < 5e-3


@pytest.mark.parametrize("B,H,T", [(2, 3, 197), (1, 2, 64), (2, 2, 499)])
def test_attn_softmax(cuda, B, H, T):
    from mtasr_b200 import kernels as Kn
    torch.manual_seed(2)
    Tp = (T + 7) // 8 * 8
    S = torch.randn(B, H, T, Tp, device=cuda) * 3
    gate = torch.rand(B, H, T, device=cuda) * 2
    table = torch.randn(H, 2 * T - 1, device=cuda)
    klen = torch.tensor([T] + [max(1, T - 13)] * (B - 1), device=cuda, dtype=torch.int32)
    scale = 0.125
    P = Kn.attn_softmax_fwd(S, gate, table, klen, B, H, T, Tp, scale)
    idx = (torch.arange(T, device=cuda)[None, :] - torch.arange(T, device=cuda)[:, None]) + T - 1      # [q,k]
    Sr = S[..., :T].clone().requires_grad_(True)
    gr = gate.clone().requires_grad_(True)
    tr = table.clone().requires_grad_(True)
    bias = gr[..., None] * tr[:, idx][None]
    z = Sr * scale + bias
    kmask = torch.arange(T, device=cuda)[None, :] >= klen[:, None]
    z = z.masked_fill(kmask[:, None, None, :], float("-inf"))
    ref = torch.softmax(z, -1)
    assert _rel(P[..., :T], ref) < 5e-3
    assert (P[..., T:] == 0).all()
    dP = torch.randn(B, H, T, Tp, device=cuda)
    Pf = P.float()[..., :T]
    dS, dgate, dtable = Kn.attn_softmax_bwd(P, dP, gate, table, B, H, T, Tp, scale)
    # reference backward evaluated at the SAME (bf16-rounded) probabilities, row dot normalised by sum(P) like the kernel
    dot = (Pf * dP[..., :T]).sum(-1, keepdim=True) / Pf.sum(-1, keepdim=True)
    dz = Pf * (dP[..., :T] - dot)
    assert _rel(dS[..., :T], dz * scale) < 5e-3
    assert _rel(dgate, (dz * table[:, idx][None]).sum(-1)) < 1e-4
    dt_ref = torch.zeros_like(table)
    dt_ref.index_add_(1, idx.reshape(-1), (dz * gate[..., None]).sum(0).reshape(H, -1))
    assert _rel(dtable, dt_ref) < 1e-4


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,H", [(2, 37, 2), (3, 199, 16), (1, 50, 12)])
def test_relpos_gate_fwd_bwd(cuda, B, T, H, dt):
    """gru_rel_pos gate (hf:167-176) and its gradients vs the torch formulation of the reference."""
    from mtasr_b200 import ops
    torch.manual_seed(3)
    D = H * 64
    h = (torch.randn(B, T, D, device=cuda)).to(dt)
    lin = torch.nn.Linear(64, 8).to(cuda)
    const = (torch.rand(1, H, 1, 1, device=cuda) + 0.5).requires_grad_(True)
    h1 = h.clone().requires_grad_(True)
    gate = ops.RelPosGateFn.apply(h1, lin.weight, lin.bias, const)
    h2 = h.float().clone().requires_grad_(True)
    proj = lin(h2.view(B, T, H, 64).permute(0, 2, 1, 3))                       # (B,H,T,8)
    ab = torch.sigmoid(proj.view(B, H, T, 2, 4).sum(-1))
    ref = (ab[..., 0:1] * (ab[..., 1:2] * const - 1.0) + 2.0).squeeze(-1)     # (B,H,T)
    assert _rel(gate, ref) < 1e-5
    up = torch.randn(B, H, T, device=cuda)
    g1 = torch.autograd.grad((gate * up).sum(), [h1, lin.weight, lin.bias, const])
    g2 = torch.autograd.grad((ref * up).sum(), [h2, lin.weight, lin.bias, const])
    tol = 1e-4 if dt == torch.float32 else 6e-3
    assert _rel(g1[0], g2[0]) < tol
    for a, b in zip(g1[1:], g2[1:]):
        assert _rel(a, b) < 1e-4


@pytest.mark.parametrize("B,T,H", [(2, 37, 2), (3, 199, 16), (1, 50, 12)])
def test_gate_path_inside_layernorm_bwd(cuda, B, T, H):
    """relpos_gate_bwd(want_dx=False) + layernorm_bwd(gate_ab=, gate_w8=): the rank-2-per-head gate path added inside the
    LayerNorm backward equals adding the materialised dx of the gate to the upstream gradient first."""
    from mtasr_b200 import kernels as Kn
    torch.manual_seed(11)
    D = H * 64
    x = torch.randn(B, T, D, device=cuda) * 2 + 0.3
    gamma, beta = torch.randn(D, device=cuda), torch.randn(D, device=cuda)
    h1, _, mean, rstd = Kn.layernorm_fwd(x, gamma, beta, 1e-5)
    w8, b8 = torch.randn(8, 64, device=cuda) * 0.2, torch.randn(8, device=cuda) * 0.1
    cst = torch.rand(H, device=cuda) + 0.5
    dgate = torch.randn(B, H, T, device=cuda)
    dxg, dw_a, db_a, dc_a = Kn.relpos_gate_bwd(h1, w8, b8, cst, dgate, B, T, H)
    dab, dw_b, db_b, dc_b = Kn.relpos_gate_bwd(h1, w8, b8, cst, dgate, B, T, H, want_dx=False)
    assert dab.shape == (B * T, H, 2)
    assert _rel(dw_a, dw_b) < 1e-5 and _rel(db_a, db_b) < 1e-5 and _rel(dc_a, dc_b) < 1e-5
    wa, wb = w8[:4].sum(0), w8[4:].sum(0)
    expand = dab[:, :, 0:1] * wa.view(1, 1, 64) + dab[:, :, 1:2] * wb.view(1, 1, 64)             # (B*T, H, 64)
    assert _rel(expand.reshape(B, T, D), dxg) < 1e-5
    dy = torch.randn(B, T, D, device=cuda).to(torch.bfloat16)
    dres = torch.randn(B, T, D, device=cuda)
    ref = Kn.layernorm_bwd(dy.float() + dxg, x, mean, rstd, gamma, dres=dres, want_bf16=True, want_dxsum=True)
    got = Kn.layernorm_bwd(dy, x, mean, rstd, gamma, dres=dres, want_bf16=True, want_dxsum=True, gate_ab=dab, gate_w8=w8)
    assert _rel(got[0], ref[0]) < 1e-5 and _rel(got[1], ref[1]) < 5e-3
    for a, b in zip(got[2:], ref[2:]):
        assert _rel(a, b) < 1e-4


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,N", [(1000, 1024), (15968, 3072), (77, 40), (333, 129)])
def test_colsum_shapes(cuda, M, N, dt):
    from mtasr_b200 import kernels as Kn
    torch.manual_seed(4)
    x = torch.randn(M, N, device=cuda).to(dt)
    assert _rel(Kn.colsum(x), x.float().sum(0)) < 1e-5


@pytest.mark.parametrize("terms", [3, 6])
@pytest.mark.parametrize("order", [0, 1])
def test_split_bf16_bit_exact_and_gemm_accuracy(cuda, order, terms):
    """mtasr_split_bf16: x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2) laid out per chunk in the A-side /
    B-side product order -- bit-exact against the same roundings in torch -- and the split-operand GEMM it feeds
    reproduces an fp64 product to ~4e-6 (two-way split) / fp32 rounding level (three-way)."""
    from mtasr_b200 import kernels as Kn
    torch.manual_seed(1)
    rows, c, n = 37, 64, 3
    x = torch.randn(rows, n * c, device=cuda) * 3
    y = Kn.split_bf16(x, c, order, terms).view(rows, n, terms, c)
    p1 = x.to(torch.bfloat16)
    r1 = x - p1.float()
    p2 = r1.to(torch.bfloat16)
    p3 = (r1 - p2.float()).to(torch.bfloat16)
    p1, p2, p3 = (t.view(rows, n, c) for t in (p1, p2, p3))
    if terms == 3:
        parts = (p2, p1, p1) if order == 0 else (p1, p2, p1)
    else:
        parts = (p3, p2, p1, p2, p1, p1) if order == 0 else (p1, p2, p3, p1, p2, p1)
    for j in range(terms):
        assert torch.equal(y[:, :, j], parts[j])
    if order == 0:
        M, N, Kd = 300, 200, 512
        a, b = torch.randn(M, Kd, device=cuda), torch.randn(N, Kd, device=cuda)
        out = torch.empty(M, N, device=cuda)
        Kn.gemm(Kn.Operand(Kn.split_bf16(a, Kd, 0, terms), terms * Kd), Kn.Operand(Kn.split_bf16(b, Kd, 1, terms), terms * Kd),
                M, N, terms * Kd, Kn.Out(out, N))
        ref = (a.double() @ b.double().t()).float()
        assert _rel(out, ref) < (1e-5 if terms == 3 else 1e-6), _rel(out, ref)


def test_weightnorm_fwd_bwd(cuda):
    """WeightNormFn (pos-conv weight_norm over the last dim, hf:48-66) vs torch._weight_norm and its autograd."""
    from mtasr_b200 import ops
    torch.manual_seed(3)
    v = torch.randn(96, 8, 128, device=cuda)
    g = torch.rand(1, 1, 128, device=cuda) + 0.5
    v1, g1 = v.clone().requires_grad_(True), g.clone().requires_grad_(True)
    v2, g2 = v.clone().requires_grad_(True), g.clone().requires_grad_(True)
    w1 = ops.WeightNormFn.apply(v1, g1)
    w2 = torch._weight_norm(v2, g2, 2)
    assert _rel(w1, w2) < 1e-6
    up = torch.randn_like(w2)
    d1 = torch.autograd.grad((w1 * up).sum(), [v1, g1])
    d2 = torch.autograd.grad((w2 * up).sum(), [v2, g2])
    assert _rel(d1[0], d2[0]) < 1e-5 and _rel(d1[1], d2[1]) < 1e-5


def test_host_prefetcher_and_scalar_reader(cuda):
    """mtasr_b200.io: staged copies arrive intact on the compute stream; pipelined scalar read-back returns each value."""
    from mtasr_b200.io import HostPrefetcher, ScalarReader
    pre, rd = HostPrefetcher(cuda), ScalarReader(depth=2)
    host = [torch.randn(4, 1000).pin_memory(), torch.arange(4000, dtype=torch.int64).view(4, 1000).pin_memory()]
    pend = []
    for i in range(5):
        st = pre.stage([t + i for t in host])
        a, b = pre.take(st)
        assert torch.equal(a.cpu(), host[0] + i) and torch.equal(b.cpu(), host[1] + i)
        pend.append(rd.submit(a.sum() * 0 + i))
        if len(pend) == 2:
            assert pend.pop(0).result() == float(i - 1)
    assert pend.pop(0).result() == 4.0


def test_bf16_twin_registry_hit_and_miss_paths(cuda):
    """ops._publish_twin / _cast_or_twin (the bf16 copy of a residual-stream gradient published by the LayerNorm backward for
    the next block's GEMMs): a hit needs the very same tensor object, unmodified; anything else -- another tensor, an in-place
    update (gradient accumulation, a hook), a twin of another size, a second consumer (retain_graph / PCGrad re-entry: the
    entry is popped by its first use) -- must fall back to a fresh cast of the values actually passed."""
    from mtasr_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(0)
    a = torch.randn(64, 128, device=cuda, generator=g)
    wrong = torch.full((64, 128), 7.0, device=cuda, dtype=torch.bfloat16)      # a twin with recognisable content
    ops._publish_twin(a, wrong)
    assert torch.equal(ops._cast_or_twin(a), wrong)                            # hit: same object, same version
    assert torch.equal(ops._cast_or_twin(a), a.to(torch.bfloat16))             # popped by the first consumer: second use casts
    ops._publish_twin(a, wrong)
    a.add_(1.0)                                                                # version bump (accumulation / hook)
    assert torch.equal(ops._cast_or_twin(a), a.to(torch.bfloat16))
    ops._publish_twin(a, wrong)
    b = a.clone()                                                              # another object with equal values
    assert torch.equal(ops._cast_or_twin(b), b.to(torch.bfloat16))
    ops._publish_twin(a, wrong[:32])                                           # size mismatch
    assert torch.equal(ops._cast_or_twin(a), a.to(torch.bfloat16))
    ops._publish_twin(a, wrong)
    del a                                                                      # the weak reference dies with the tensor
    c = torch.randn(64, 128, device=cuda, generator=g)                        # may reuse the id: the dead weakref is no match
    assert torch.equal(ops._cast_or_twin(c), c.to(torch.bfloat16))
