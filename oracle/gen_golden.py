"""Pin the oracle against THE REFERENCE ITSELF and write tests/golden/*.npz.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python oracle/gen_golden.py
For every fixture the real reference modules (ref:models/modeling_wavlm.py, separator.py, ctc.py, losses.py,
utils/split_labels_by_sc.py, ctc_prompt.py, mt_ctctoken_builder.py and the AST-extracted
ctc_remove_duplicates_and_blank of ref:models/modeling_speech_encoder_decoder_llama.py:902-972, whose module
does not import under transformers 5.x) are executed on seeded inputs; the oracle restatements are asserted
equal; inputs, weights and the REFERENCE's outputs are stored.  Test infrastructure only.
"""
import ast
import os
import sys
import textwrap
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from oracle import ctc_ref, host_ref  # noqa: E402
from oracle.model_ref import (RefCTC, RefSeparator, RefWavLMModel, make_config, ref_hybrid_ctc,  # noqa: E402
                              synth_batch)

import models.modeling_wavlm as ref_wavlm  # noqa: E402
from models.ctc import CTC as RefOrigCTC  # noqa: E402
from models.ctc_prompt import build_multi_ctc_prefix_from_heads  # noqa: E402
from models.losses import HybridLoss  # noqa: E402
from models.mt_ctctoken_builder import MultiSpkCTCTokenBuilder  # noqa: E402
from models.separator import Separator as RefOrigSeparator  # noqa: E402
from utils.split_labels_by_sc import split_k_speakers_and_lengths  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def extract_collapse():
    src = open(os.path.join(REF, "models/modeling_speech_encoder_decoder_llama.py")).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "ctc_remove_duplicates_and_blank":
            code = textwrap.dedent(ast.get_source_segment(src, node))
            ns = {}
            exec("import torch\nfrom typing import *\nfrom torch.nn.utils.rnn import pad_sequence\n" + code, ns)
            return ns["ctc_remove_duplicates_and_blank"]
    raise RuntimeError("collapse not found")


def npz(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def sd_arrays(prefix, module):
    return {prefix + k: v for k, v in module.state_dict().items()}


def gen_model(kind, seed, B, S, n_spk, vocab, hidden_sep):
    torch.manual_seed(seed)
    cfg = make_config(kind)
    ref = ref_wavlm.WavLMModel(cfg).eval()
    with torch.no_grad():
        for lyr in ref.encoder.layers:
            lyr.attention.gru_rel_pos_const.uniform_(0.5, 1.5)
            lyr.attention.gru_rel_pos_linear.bias.normal_(0, 0.5)
        ref.encoder.layers[0].attention.rel_attn_embed.weight.normal_(0, 0.5)
        for n, p in ref.named_parameters():
            if "layer_norm" in n or "pre_ln" in n:
                p.add_(0.1 * torch.randn_like(p))
    sep = RefOrigSeparator(cfg.hidden_size, hidden_sep, n_spk).eval()
    heads = torch.nn.ModuleList(RefOrigCTC(vocab, cfg.hidden_size) for _ in range(n_spk))
    with torch.no_grad():
        for h in heads:
            h.ctc_lo.weight.mul_(3.0)
    loss_mod = HybridLoss(mode="ctc", blank_id=vocab - 1)
    wav, mask, labels, lab_lens = synth_batch(B, S, n_spk, vocab, seed=seed, varlen=True, tok_per_sec=(8, 16))

    out = ref(wav, attention_mask=mask)
    last, enc, down, feats = out[0], out[1], out[2], out[3]
    sep_out = sep(enc)
    fmask = ref._get_feature_vector_attention_mask_x0(enc.shape[1], mask)
    loss = loss_mod(talker_ctc=heads, sep_hidden_states=sep_out, encoder_attention_mask_ctc=fmask,
                    label_spks=labels, label_spks_lengths=lab_lens, talker_numbers=n_spk)
    params = {**{"encoder." + k: v for k, v in ref.named_parameters()},
              **{"separator." + k: v for k, v in sep.named_parameters()},
              **{"serialized_ctc." + k: v for k, v in heads.named_parameters()}}
    names = [k for k, v in params.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    argmax = [h.argmax(x) for h, x in zip(heads, sep_out)]

    # --- oracle restatement must agree with the reference on the same weights -------------------------------
    o_enc = RefWavLMModel(cfg).eval(); o_enc.load_state_dict(ref.state_dict())
    o_sep = RefSeparator(cfg.hidden_size, hidden_sep, n_spk).eval(); o_sep.load_state_dict(sep.state_dict())
    o_heads = [RefCTC(vocab, cfg.hidden_size) for _ in range(n_spk)]
    for oh, h in zip(o_heads, heads):
        oh.load_state_dict({k: v for k, v in h.state_dict().items()})
    ol, oe, od, of = o_enc(wav, mask)
    osep = o_sep(oe)
    oloss, _ = ref_hybrid_ctc(o_heads, osep, o_enc.frame_mask_x0(oe.shape[1], mask), labels, lab_lens)
    for a, b, nm in [(ol, last, "last"), (oe, enc, "enc"), (od, down, "down"), (of, feats, "feats"),
                     (osep[0], sep_out[0], "sep0"), (oloss, loss, "loss")]:
        err = (a - b).abs().max().item()
        assert err < 1e-5, (nm, err)
    # numpy fp64 CTC restatement vs the reference CTC module (value and gradient wrt logits)
    hl = fmask.sum(1)
    for h, x, y, yl in zip(heads, sep_out, labels, lab_lens):
        logits = h.ctc_lo(x).detach().requires_grad_(True)
        lp = logits.transpose(0, 1).log_softmax(2)
        nll = h.ctc_loss(lp, torch.cat([y[i, :l] for i, l in enumerate(yl)]), hl, yl)
        (g,) = torch.autograd.grad(nll.sum(), logits)
        n64, g64 = ctc_ref.ctc_loss_and_grad(logits.detach().numpy(), hl, y.numpy(), yl, vocab - 1)
        assert np.abs(n64 - nll.detach().numpy()).max() < 1e-3 * max(1.0, np.abs(n64).max()), "ctc nll"
        assert np.abs(g64 - g.numpy()).max() < 1e-4, "ctc grad"
    print(f"[{kind}] oracle == reference; loss {loss.item():.6f}")

    fx = dict(kind=kind, seed=seed, n_spk=n_spk, vocab=vocab, hidden_sep=hidden_sep, wav=wav, mask=mask,
              last=last, enc=enc, down=down, feats=feats, loss=loss, frame_mask=fmask)
    for i in range(n_spk):
        fx[f"labels{i}"] = labels[i]; fx[f"lab_lens{i}"] = lab_lens[i]
        fx[f"sep{i}"] = sep_out[i]; fx[f"argmax{i}"] = argmax[i]
    for k, v in params.items():
        fx["p:" + k] = v
    for k, gval in zip(names, grads):
        if gval is not None:
            fx["g:" + k] = gval
    npz(f"model_{kind}.npz", **fx)


def gen_ctc():
    g = torch.Generator().manual_seed(7)
    B, T, V = 6, 40, 23
    logits = torch.randn(B, T, V, generator=g) * 2
    hlens = torch.tensor([40, 33, 40, 9, 40, 5])
    ys = torch.randint(0, V - 2, (B, 12), generator=g)
    ys[0, 3] = ys[0, 2]; ys[2, :6] = 4                       # repeats
    ylens = torch.tensor([12, 7, 6, 8, 0, 5])                # row 3: infeasible (T=9 < needed) ; row 4: empty target
    ys[5, :5] = torch.tensor([1, 1, 1, 1, 1])                # needs 9 frames, has 5 -> infeasible
    tgt = torch.cat([ys[i, :l] for i, l in enumerate(ylens)])
    w = torch.arange(1, B + 1, dtype=torch.float32)
    n64, g64 = ctc_ref.ctc_loss_and_grad(logits.numpy(), hlens, ys.numpy(), ylens, V - 1)
    for dt, tol_n, tol_g in ((torch.float64, 1e-10, 1e-10), (torch.float32, 1e-4, 5e-4)):
        lg = logits.to(dt).requires_grad_(True)
        lp = lg.transpose(0, 1).log_softmax(2)
        nll = torch.nn.CTCLoss(reduction="none", zero_infinity=True, blank=V - 1)(lp, tgt, hlens, ylens)
        (grad,) = torch.autograd.grad((nll * w.to(dt)).sum(), lg)
        assert np.abs(n64 - nll.detach().numpy()).max() < tol_n, dt
        assert np.abs(g64 * w.numpy()[:, None, None] - grad.numpy()).max() < tol_g, dt
        if dt == torch.float64:
            nll_f64, grad_f64 = nll.detach(), grad
    # brute force on a tiny case
    lp_small = ctc_ref.log_softmax(np.random.RandomState(0).randn(5, 4))
    for y in ([0], [1, 1], [0, 2], []):
        a = ctc_ref.ctc_alpha_beta(lp_small, y, 3)[0]
        b = ctc_ref.ctc_brute_force(lp_small, y, 3)
        assert abs(a - b) < 1e-10, (y, a, b)
    print("[ctc] numpy oracle == torch.nn.CTCLoss == brute force")
    npz("ctc_small.npz", logits=logits, hlens=hlens, ys=ys, ylens=ylens, nll=nll_f64, upstream=w, grad=grad_f64, blank=V - 1,
        nll_f32=nll, grad_f32=grad)


def gen_host():
    collapse = extract_collapse()
    g = torch.Generator().manual_seed(11)
    blank, pad = 18, 17
    am = torch.randint(0, 19, (7, 50), generator=g)
    am[1] = blank; am[2, :] = 5; am[3, ::2] = blank; am[4, 10:] = pad
    am[5] = torch.tensor(([3, blank, 3, 3, 4, blank, blank, 4, 3, pad] * 5))
    for cab in (True, False):
        padded, lens = collapse(None, am, blank_id=blank, pad_id=pad, collapse_across_blanks=cab)
        rows, l2 = host_ref.collapse(am.tolist(), blank, pad)
        assert lens == l2 and padded.tolist() == host_ref.pad_rows(rows, pad), "collapse"
    empty, elens = collapse(None, torch.full((3, 9), blank), blank_id=blank, pad_id=pad)
    assert tuple(empty.shape) == (3, 0)

    # label splitter (ref:utils/split_labels_by_sc.py)
    sc, padt = 30, 31
    labels = torch.tensor([[1, 2, 3, sc, 4, 5, padt, padt, -100],
                           [7, sc, 8, 9, 10, 11, 12, padt, padt],
                           [5, 6, 7, 8, sc, 9, padt, -100, -100]])
    labs, lens_ = split_k_speakers_and_lengths(labels, 2, sc, padt, ignore_id=-100, end_token_id=padt, allow_empty_segment=False)
    olabs, olens = host_ref.split_labels(labels.tolist(), 2, sc, padt, -100, padt, False)
    assert [x.tolist() for x in labs] == olabs and [x.tolist() for x in lens_] == olens

    # prefix builder (ref:models/ctc_prompt.py)
    class Dec(torch.nn.Module):
        def __init__(self):
            super().__init__(); self.emb = torch.nn.Embedding(32, 4)
        def get_input_embeddings(self): return self.emb
    h0 = torch.tensor([[1, 2, padt], [3, padt, padt], [padt, padt, padt]])
    h1 = torch.tensor([[4, padt], [5, 6], [7, padt]])
    emb, m, ids = build_multi_ctc_prefix_from_heads([h0, h1], Dec(), padt, None)
    oids, omask = host_ref.prefix_ids([h0.tolist(), h1.tolist()], padt, None)
    assert ids.tolist() == oids and m.tolist() == omask

    # token builder segmentation (ref:models/mt_ctctoken_builder.py)
    torch.manual_seed(3)
    ctc = RefOrigCTC(9, 16)
    with torch.no_grad():
        ctc.ctc_lo.weight.mul_(20)
    x = torch.randn(3, 30, 16)
    fm = torch.ones(3, 30, dtype=torch.bool); fm[1, 21:] = False
    tb = MultiSpkCTCTokenBuilder()
    mem, mmask, conf = tb([x, x.flip(1)], fm, [ctc, ctc])
    path = ctc.argmax(x)
    segs = [host_ref.token_segments(path[b].tolist(), fm[b].tolist(), 8) for b in range(3)]
    L0 = max(len(s) for s in segs)
    for b in range(3):
        for j, s in enumerate(segs[b]):
            assert torch.allclose(mem[b, j], x[b, s].mean(0), atol=1e-6)
        assert (~mmask[b, :L0]).sum().item() == len(segs[b])
    print("[host] collapse / split / prefix / segments restatements == reference")
    npz("host_small.npz", argmax=am, blank=blank, pad=pad, collapsed=collapse(None, am, blank_id=blank, pad_id=pad)[0],
        collapsed_lens=np.array(collapse(None, am, blank_id=blank, pad_id=pad)[1]),
        split_labels=labels, split_sc=sc, split_pad=padt, split0=labs[0], split1=labs[1], split_len0=lens_[0], split_len1=lens_[1],
        prefix_h0=h0, prefix_h1=h1, prefix_ids=ids, prefix_mask=m,
        tb_x=x, tb_mask=fm, tb_w=ctc.ctc_lo.weight, tb_b=ctc.ctc_lo.bias, tb_mem=mem, tb_memmask=mmask, tb_conf=conf)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_ctc()
    gen_host()
    gen_model("tiny_large", 101, B=3, S=6000, n_spk=2, vocab=40, hidden_sep=96)
    gen_model("tiny_base", 202, B=2, S=5200, n_spk=3, vocab=33, hidden_sep=64)
