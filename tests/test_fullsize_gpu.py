"""Size-independent properties of the kernels at the FULL sizes of BASELINE.json (cfg2: B=32, T=499, D=1024, H=16,
V=128259; cfg5: T=1499), where the CPU oracle cannot run in seconds:

  * softmax rows sum to one  -> attention with V = 1 returns 1 on every valid row, whatever Q, K, gate and table are;
  * linearity of the backward in the upstream gradient (attention, CTC head);
  * checksum of checksums    -> the dense CTC gradient softmax - occupancy sums to zero over the vocabulary of every valid
                                frame, hence the bias gradient of the head sums to ~0 and is exactly 0 on padded frames;
  * idempotence / sortedness -> collapse(collapse(x)) == collapse(x), no blank / pad / equal neighbours survive;
  * closed form              -> an LSTM with zero recurrent weights is a pointwise function of its input gates;
  * sampled-row equality     -> the vocabulary GEMM's LSE / argmax against torch on a random sample of rows.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


@pytest.mark.parametrize("B,H,T", [(32, 16, 499), (4, 16, 1499)])
def test_attention_rows_are_stochastic_and_backward_is_linear(cuda, B, H, T):
    from mtasr_b200 import kernels as Kn
    g = torch.Generator(device=cuda).manual_seed(7)
    D = H * 64
    qkv = (torch.randn(B * T, 3 * D, device=cuda, generator=g) * 0.8).to(torch.bfloat16)
    qkv[:, 2 * D:] = 1.0                                                  # V = 1  =>  O = sum_k P[q,k] = 1
    gate = torch.rand(B, H, T, device=cuda, generator=g) * 2
    table = torch.randn(H, 2 * T - 1, device=cuda, generator=g)
    klen = torch.randint(T // 2, T + 1, (B,), device=cuda, generator=g, dtype=torch.int32)
    klen[0] = T
    out, lse = Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125)
    assert (out.float() - 1.0).abs().max().item() < 1e-2                  # bf16 P, fp32 accumulation, bf16 output
    assert torch.isfinite(lse).all()
    # backward: with V = 1, dP = rowsum(dO) for every key, so dZ = P (dP - delta) = 0: no gradient reaches q, k, the gate
    # or the table, and dV[k] = sum_q P[q,k] dO[q] sums over the keys to sum_q dO[q]
    dout = (torch.randn(B * T, D, device=cuda, generator=g) * 0.5).to(torch.bfloat16)
    dqkv, dgate, dtable = Kn.attn_bwd(qkv, out, dout, lse, gate, table, klen, B, H, T, 0.125)
    dq, dk, dv = dqkv[:, :D].float(), dqkv[:, D:2 * D].float(), dqkv[:, 2 * D:].float()
    scale_ref = dv.abs().mean().item()
    assert dq.abs().mean().item() < 0.05 * scale_ref and dk.abs().mean().item() < 0.05 * scale_ref   # bf16 O != 1 exactly
    dv_sum = dv.view(B, T, D).sum(1)
    do_sum = dout.float().view(B, T, D).sum(1)
    assert _rel(dv_sum, do_sum) < 1e-2
    for b in range(B):                                                    # padded keys receive nothing
        kl = int(klen[b])
        if kl < T:
            assert dqkv.view(B, T, 3 * D)[b, kl:, D:].abs().max().item() == 0.0
    # linearity in the upstream gradient
    dqkv2, dgate2, dtable2 = Kn.attn_bwd(qkv, out, (dout.float() * 2).to(torch.bfloat16), lse, gate, table, klen, B, H, T, 0.125)
    assert _rel(dqkv2[:, 2 * D:], 2 * dv) < 1e-2


def test_ctc_head_full_vocab_checksums(cuda):
    """Full LLaMA-3 vocabulary (V = 128259), T = 499: gradients of the fused head."""
    from mtasr_b200.ctc import CTC
    torch.manual_seed(3)
    V, D, B, T = 128259, 1024, 4, 499
    head = CTC(V, D).to(cuda)
    hs = torch.randn(B, T, D, device=cuda, requires_grad=True)
    hlens = torch.tensor([T, T - 50, T // 2, 300], device=cuda)
    Lmax = 60
    ys = torch.randint(0, V - 2, (B, Lmax), device=cuda)
    ylens = torch.tensor([60, 33, 20, 0], device=cuda)
    nll = head.per_utterance_nll(hs, hlens, ys, ylens)
    assert torch.isfinite(nll).all() and (nll >= 0).all()
    up = torch.tensor([1.0, 0.5, 2.0, 1.5], device=cuda)
    dh, dw, db = torch.autograd.grad((nll * up).sum(), [hs, head.ctc_lo.weight, head.ctc_lo.bias], retain_graph=True)
    # sum over the vocabulary of (softmax - occupancy) is zero on every valid frame => the bias gradient sums to ~0
    assert abs(db.sum().item()) < 2e-3 * db.abs().sum().item()
    # padded frames carry no gradient at all
    for b in range(B):
        if int(hlens[b]) < T:
            assert dh[b, int(hlens[b]):].abs().max().item() == 0.0
    # blank dominates an empty target: its bias gradient is negative (occupancy 1 per frame), everything else positive mass
    assert db[V - 1].item() < 0
    # linearity in the upstream gradient (same saved forward)
    dh2, dw2, db2 = torch.autograd.grad((nll * up * 3).sum(), [hs, head.ctc_lo.weight, head.ctc_lo.bias])
    assert _rel(dh2, 3 * dh) < 1e-2 and _rel(db2, 3 * db) < 1e-2 and _rel(dw2[:4096], 3 * dw[:4096]) < 1e-2
    # sampled rows of the vocabulary GEMM: argmax against torch in fp32 wherever the top-2 margin is clear
    with torch.no_grad():
        rows = torch.randint(0, B * T, (64,), device=cuda)
        logits = hs.detach().view(B * T, D)[rows].to(torch.bfloat16).float() @ head.ctc_lo.weight.to(torch.bfloat16).float().t() \
            + head.ctc_lo.bias
        top2 = logits.topk(2, -1).values
        clear = (top2[:, 0] - top2[:, 1]) > 1e-3
        am = head.argmax(hs.detach()).view(-1)[rows]
        assert torch.equal(am[clear], logits.argmax(-1)[clear])


@pytest.mark.parametrize("B,T", [(64, 1499), (32, 499)])
def test_collapse_idempotent_at_full_length(cuda, B, T):
    from mtasr_b200.greedy import ctc_remove_duplicates_and_blank
    g = torch.Generator(device=cuda).manual_seed(11)
    blank, pad = 128258, 128257
    ids = torch.randint(0, 40, (B, T), device=cuda, generator=g)
    ids[ids >= 30] = blank                                                # ~25 % blanks
    ids[ids == 29] = pad
    ids[0] = blank                                                        # all-blank row -> length 0
    ids[1, ::2] = 5
    ids[1, 1::2] = blank                                                  # A,blank,A,...  -> "A" under the reference's rule
    out, lens = ctc_remove_duplicates_and_blank(ids, blank_id=blank, pad_id=pad)
    assert lens[0] == 0 and lens[1] == 1 and out[1, 0].item() == 5
    assert out.shape == (B, max(lens))
    for b in range(B):
        row = out[b, :lens[b]]
        assert (row != blank).all() and (row != pad).all()
        if lens[b] > 1:
            assert (row[1:] != row[:-1]).all()                            # no equal neighbours survive
        assert (out[b, lens[b]:] == pad).all()
    out2, lens2 = ctc_remove_duplicates_and_blank(out, blank_id=blank, pad_id=pad)
    assert lens2 == lens and torch.equal(out2, out)                       # idempotent


def test_lstm_zero_recurrence_closed_form(cuda):
    """Hs = 896, B = 32, T = 499 (the unrolled kernel instantiations): W_hh = 0 makes every step pointwise."""
    from mtasr_b200 import kernels as Kn
    torch.manual_seed(5)
    B, T, Hs = 32, 499, 896
    xg = torch.randn(B, T, 4 * Hs, device=cuda)
    W = torch.zeros(4 * Hs, 2 * Hs, device=cuda, dtype=torch.bfloat16)
    h, hf, c, gates = Kn.lstm_fwd(xg, W[:, Hs:], 2 * Hs, want_h_f32=True)
    i, f, g_, o = (xg[..., k * Hs:(k + 1) * Hs] for k in range(4))
    i, f, o = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o)
    g_ = torch.tanh(g_)
    c_ref = torch.zeros(B, Hs, device=cuda)
    worst = 0.0
    for t in range(0, T):
        c_ref = f[:, t] * c_ref + i[:, t] * g_[:, t]
        if t % 50 == 0 or t == T - 1:
            worst = max(worst, (hf[:, t] - o[:, t] * torch.tanh(c_ref)).abs().max().item(), (c[:, t] - c_ref).abs().max().item())
    assert worst < 2e-5, worst
    # backward with zero recurrence: dgates of the last step depend on dh_T only
    dh = torch.randn(B, T, Hs, device=cuda)
    dg = Kn.lstm_bwd(dh, gates, c, W[:, Hs:], 2 * Hs).float()
    tc = torch.tanh(c[:, -1])
    dao = dh[:, -1] * tc * o[:, -1] * (1 - o[:, -1])
    assert _rel(dg[:, -1, 3 * Hs:], dao) < 1e-2
