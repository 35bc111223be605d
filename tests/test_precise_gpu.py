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
    from oracle.model_ref import RefCTC
    from mtasr_b200 import precise
    from mtasr_b200.ctc import CTC
    torch.manual_seed(4)
    head = CTC(V, D).to(cuda)
    with torch.no_grad():
        head.ctc_lo.weight.mul_(3.0)
    ref = RefCTC(V, D).to(cuda)
    ref.ctc_lo.load_state_dict(head.ctc_lo.state_dict())
    hs = torch.randn(B, T, D, device=cuda)
    hlens = torch.tensor([T, T - 7, T // 2, 5][:B], device=cuda)
    Lmax = 14
    ys = torch.randint(0, V - 2, (B, Lmax), device=cuda)
    ys[:, 3] = ys[:, 2]
    ylens = torch.tensor([Lmax, 9, 0, 12][:B], device=cuda)       # row 2: empty target; row 3: infeasible (5 frames)
    h1 = hs.clone().requires_grad_(True)
    h2 = hs.clone().requires_grad_(True)
    up = torch.rand(B, device=cuda) + 0.5
    with precise.precision("fp32"):
        n1 = head.per_utterance_nll(h1, hlens, ys, ylens)
        g1 = torch.autograd.grad((n1 * up).sum(), [h1, head.ctc_lo.weight, head.ctc_lo.bias])
        am = head.argmax(hs)
        scalar = head(hs, hlens, ys, ylens)
    n2 = ref.per_utt_nll(h2, hlens, ys, ylens)
    g2 = torch.autograd.grad((n2 * up).sum(), [h2, ref.ctc_lo.weight, ref.ctc_lo.bias])
    errs = dict(nll=rel(n1, n2), dh=rel(g1[0], g2[0]), dw=rel(g1[1], g2[1]), db=rel(g1[2], g2[2]))
    with torch.no_grad():
        logits = ref.ctc_lo(hs)
        am_ref = logits.argmax(-1)
        top2 = logits.topk(2, -1).values
        errs["min_top2_margin"] = (top2[..., 0] - top2[..., 1]).min().item()
        errs["argmax_mismatches"] = int((am != am_ref).sum().item())
    _log(test="ctc_head_fp32_mode", V=V, **errs)
    assert errs["nll"] < 1e-5, errs
    assert errs["dh"] < 1e-4 and errs["dw"] < 1e-4 and errs["db"] < 1e-4, errs
    if B > 3:
        assert n1[3].item() == 0.0 and g1[0][3].abs().max().item() == 0.0
    assert abs(scalar.item() - ref(hs, hlens, ys, ylens).item()) < 1e-5 * abs(ref(hs, hlens, ys, ylens).item())
    assert torch.equal(am, am_ref), errs                           # every frame, no margin mask


@pytest.mark.parametrize("B,T,D,Hs", [(3, 37, 128, 96), (4, 120, 1024, 896)])
def test_separator_fp32_mode_vs_oracle(cuda, B, T, D, Hs):
    from oracle.model_ref import RefSeparator
    from mtasr_b200 import precise
    from mtasr_b200.separator import Separator
    torch.manual_seed(2)
    sep = Separator(D, Hs, 2).to(cuda).eval()
    ref = RefSeparator(D, Hs, 2).to(cuda).eval()
    ref.load_state_dict(sep.state_dict())
    x = torch.randn(B, T, D, device=cuda)
    x1 = x.clone().requires_grad_(True)
    x2 = x.clone().requires_grad_(True)
    w = [torch.randn(B, T, D, device=cuda) for _ in range(2)]
    with precise.precision("fp32"):
        y1 = sep(x1)
        g1 = torch.autograd.grad(sum((a * b).sum() for a, b in zip(y1, w)), [x1] + list(sep.parameters()))
    y2 = ref(x2)
    g2 = torch.autograd.grad(sum((a * b).sum() for a, b in zip(y2, w)), [x2] + list(ref.parameters()))
    names = ["x"] + [n for n, _ in sep.named_parameters()]
    errs = {n: rel(a, b) for n, a, b in zip(names, g1, g2)}
    fwd = max(rel(y1[0], y2[0]), rel(y1[1], y2[1]))
    _log(test="separator_fp32_mode", Hs=Hs, fwd=fwd, worst=max(errs.values()), grads={k: float(f"{v:.3g}") for k, v in errs.items()})
    assert y1[0].dtype == torch.float32
    assert fwd < 1e-4, fwd
    assert max(errs.values()) < 2e-4, errs


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
