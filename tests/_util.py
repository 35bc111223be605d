"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b, mask=None):
    """||a-b||_2 / ||b||_2, optionally over the rows selected by a (B,T) boolean mask (valid frames only)."""
    a, b = a.detach().float(), b.detach().float()
    if mask is not None:
        a, b = a[mask], b[mask]
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def load_model_golden(kind):
    g = np.load(os.path.join(GOLDEN, f"model_{kind}.npz"))
    params = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p:")}
    grads = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("g:")}
    return g, params, grads


def sub_state(params, prefix):
    return {k[len(prefix):]: v for k, v in params.items() if k.startswith(prefix)}


def build_ours(cfg, n_spk, hidden_sep, vocab, params=None, device="cuda"):
    """Instantiate the product modules (WavLMModel, Separator, CTC heads, HybridLoss), optionally with given weights."""
    from mtasr_b200.ctc import CTC
    from mtasr_b200.losses import HybridLoss
    from mtasr_b200.modeling_wavlm import WavLMModel
    from mtasr_b200.separator import Separator
    enc = WavLMModel(cfg)
    sep = Separator(cfg.hidden_size, hidden_sep, n_spk)
    heads = torch.nn.ModuleList(CTC(vocab, cfg.hidden_size) for _ in range(n_spk))
    if params is not None:
        enc.load_state_dict(sub_state(params, "encoder."), strict=True)
        sep.load_state_dict(sub_state(params, "separator."), strict=True)
        heads.load_state_dict(sub_state(params, "serialized_ctc."), strict=False)
    enc, sep, heads = enc.to(device).eval(), sep.to(device).eval(), heads.to(device).eval()
    enc.freeze_feature_encoder()
    return enc, sep, heads, HybridLoss(mode="ctc", blank_id=vocab - 1)


def build_oracle(cfg, n_spk, hidden_sep, vocab, ours=None, device="cuda"):
    """The oracle's restatement of the reference modules (fp32 torch), sharing the weights of `ours` when given."""
    from oracle.model_ref import RefCTC, RefSeparator, RefWavLMModel
    enc = RefWavLMModel(cfg)
    sep = RefSeparator(cfg.hidden_size, hidden_sep, n_spk)
    heads = torch.nn.ModuleList(RefCTC(vocab, cfg.hidden_size) for _ in range(n_spk))
    if ours is not None:
        o_enc, o_sep, o_heads = ours
        enc.load_state_dict(o_enc.state_dict(), strict=True)
        sep.load_state_dict(o_sep.state_dict(), strict=True)
        for h, oh in zip(heads, o_heads):
            h.ctc_lo.load_state_dict(oh.ctc_lo.state_dict())
    return enc.to(device).eval(), sep.to(device).eval(), heads.to(device).eval()


def perturb_(enc, seed=0):
    """Make the gate / rel-pos / LayerNorm parameters non-trivial (random init leaves them at 1/0)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for lyr in enc.encoder.layers:
            a = lyr.attention
            a.gru_rel_pos_const.copy_((torch.rand(a.gru_rel_pos_const.shape, generator=g) + 0.5).to(a.gru_rel_pos_const.device))
            b = a.gru_rel_pos_linear.bias
            b.copy_((torch.randn(b.shape, generator=g) * 0.5).to(b.device))
        w = enc.encoder.layers[0].attention.rel_attn_embed.weight
        w.copy_((torch.randn(w.shape, generator=g) * 0.5).to(w.device))
        for n, p in enc.named_parameters():
            if "layer_norm" in n:
                p.add_((0.1 * torch.randn(p.shape, generator=g)).to(p.device))


def oracle_run(o_enc, o_sep, o_heads, wav, mask, labels, lens, autocast=False, want_grads=True):
    """Oracle forward (+ backward) of the serialized-CTC loss.  autocast=True gives the yardstick "what the reference
    itself loses in bf16" (torch.autocast on encoder + separator; the CTC block stays fp32 like
    ref:models/losses.py:265-268)."""
    from oracle.model_ref import ref_hybrid_ctc
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            last, enc, down, feats = o_enc(wav, mask)
            seps = o_sep(enc)
        fm = o_enc.frame_mask_x0(enc.shape[1], mask)
        loss, _ = ref_hybrid_ctc(o_heads, [s.float() for s in seps], fm, labels, lens)
        grads = None
        if want_grads:
            named = named_params(o_enc, o_sep, o_heads)
            names = [k for k, v in named.items() if v.requires_grad]
            gs = torch.autograd.grad(loss, [named[k] for k in names], allow_unused=True)
            grads = {k: v for k, v in zip(names, gs) if v is not None}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
    return dict(last=last.float(), enc=enc.float(), down=down.float(), feats=feats.float(), seps=[s.float() for s in seps],
                loss=loss, fm=fm, grads=grads)


def named_params(enc, sep, heads):
    return {**{"encoder." + k: v for k, v in enc.named_parameters()},
            **{"separator." + k: v for k, v in sep.named_parameters()},
            **{"serialized_ctc." + k: v for k, v in heads.named_parameters()}}


def grad_errors(got, ref):
    """global relative L2 error over all parameters in `ref`, and per-parameter errors."""
    num = den = 0.0
    per = {}
    for k, r in ref.items():
        gv = got.get(k)
        r = r.float()
        if gv is None:
            gv = torch.zeros_like(r)
        gv = gv.float().to(r.device)
        num += (gv - r).pow(2).sum().item()
        den += r.pow(2).sum().item()
        if r.norm().item() > 1e-8:
            per[k] = ((gv - r).norm() / r.norm()).item()
    return (num / max(den, 1e-30)) ** 0.5, per
