"""`mtasr_b200.composite.SpeechEncoderDecoderModelLlama` against THE REFERENCE'S OWN composite model.

The fixture (oracle/gen_golden_composite.py) holds the reference class's state_dict, inputs and outputs, produced by
ref:models/modeling_speech_encoder_decoder_llama.py itself under the transformers-5.x shim (mtasr_b200.compat) on CPU in
fp32.  Here the same weights are loaded (strict) into the B200-native composite and every entry point the trainer / the
inference scripts use is compared:  forward in 'ctc' and 'hybrid' mode (loss, per-head CTC, decoder logits), forward_ctc,
the 'ctcprompt' bridge, and greedy decoding -- cached (row f2) and with the reference's per-token recomputation.
fp32 mode: tight tolerances and exact token ids; bf16 throughput mode: the encoder's bf16 tolerance."""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, rel

pytestmark = pytest.mark.gpu


@pytest.fixture()
def setup(cuda):
    from oracle.model_ref import make_composite_config
    from mtasr_b200.composite import SpeechEncoderDecoderModelLlama
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLDEN, "composite_tiny.npz"))
    cfg = make_composite_config(int(g["meta"][0]))
    cfg.decoder.max_position_embeddings = 512
    model = SpeechEncoderDecoderModelLlama(cfg)
    model.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p:")}, strict=True)
    model = model.to(cuda).eval()
    model.freeze_feature_encoder()
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    yield model, g, t
    torch.backends.cuda.matmul.allow_tf32 = prev


def test_forward_losses_logits_and_ctc_tokens_fp32_mode(setup):
    from mtasr_b200 import precise
    model, g, t = setup
    wav, mask, labels = t("wav"), t("mask"), t("labels")
    with torch.no_grad(), precise.precision("fp32"):
        model.losses.mode = "ctc"
        out = model(inputs=wav, attention_mask=mask, labels=labels.clone())
        assert abs(out.loss.item() - float(g["ctc_loss"])) < 2e-5 * abs(float(g["ctc_loss"]))
        assert rel(torch.stack(list(out.ctc_per_head)), t("ctc_per_head")) < 2e-5
        model.losses.mode = "hybrid"
        out = model(inputs=wav, attention_mask=mask, labels=labels.clone())
        assert abs(out.loss.item() - float(g["hybrid_loss"])) < 5e-5 * abs(float(g["hybrid_loss"]))
        assert out.logits.shape == tuple(g["hybrid_logits"].shape)
        assert rel(out.logits, t("hybrid_logits")) < 2e-4
        assert rel(out.encoder_last_hidden_state, t("enc_last")) < 1e-4
        ids = model.forward_ctc(inputs=wav, attention_mask=mask)
        assert ids.cpu().tolist() == g["forward_ctc"].tolist()          # greedy CTC tokens: bit-exact


def test_ctcprompt_bridge_and_greedy_decoding_fp32_mode(setup):
    """ref :644-668 + row f2: the prompt prefix is built from the greedy CTC transcripts; cached decoding (separator, heads,
    collapse, prefix computed once) returns the tokens of the reference's per-token recomputation."""
    from mtasr_b200 import kernels as K
    from mtasr_b200 import precise
    model, g, t = setup
    wav, mask, labels = t("wav"), t("mask"), t("labels")
    model.ctc_bridge, model.ctc_bridge_type = True, "ctcprompt"
    model.losses.mode = "hybrid"
    steps = int(g["meta"][3])
    with torch.no_grad(), precise.precision("fp32"):
        out = model(inputs=wav, attention_mask=mask, labels=labels.clone())
        assert out.encoder_last_hidden_state.shape == tuple(g["prompt_enc_last"].shape)     # same prefix length
        assert rel(out.encoder_last_hidden_state, t("prompt_enc_last")) < 1e-4
        assert abs(out.loss.item() - float(g["prompt_hybrid_loss"])) < 5e-5 * abs(float(g["prompt_hybrid_loss"]))
        assert rel(out.logits, t("prompt_logits")) < 2e-4
        l0 = K.launch_count()
        slow = model.generate(wav, attention_mask=mask, max_new_tokens=steps, recompute=True)
        l_slow = K.launch_count() - l0
        l0 = K.launch_count()
        fast = model.generate(wav, attention_mask=mask, max_new_tokens=steps)
        l_fast = K.launch_count() - l0
    assert slow.cpu().tolist() == g["prompt_greedy_ids"].tolist()
    assert fast.cpu().tolist() == g["prompt_greedy_ids"].tolist()
    # the hot-path kernels (separator, N vocabulary GEMMs, collapse) run once instead of once per token
    assert l_slow > l_fast + (steps - 1) * 20, (l_slow, l_fast)


def test_bf16_mode_forward_backward_and_cached_decoding(setup):
    model, g, t = setup
    wav, mask, labels = t("wav"), t("mask"), t("labels")
    model.losses.mode = "hybrid"
    out = model(inputs=wav, attention_mask=mask, labels=labels.clone())
    assert abs(out.loss.item() - float(g["hybrid_loss"])) < 2e-2 * abs(float(g["hybrid_loss"]))
    assert rel(out.logits, t("hybrid_logits")) < 5e-2
    assert len(out.ctc_per_head) == 2 and out.ctc_per_head[0].shape == (wav.shape[0],) and out.ctc_per_head[0].requires_grad
    out.loss.backward()
    for name in ("encoder.encoder.layers.0.attention.q_proj.weight", "separator.lstm.cells.0.W.weight", "serialized_ctc.1.ctc_lo.weight",
                 "decoder.model.layers.0.self_attn.q_proj.weight", "encoder.adapter.layers.0.conv.weight"):
        p = dict(model.named_parameters())[name]
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum().item() > 0, name
    # PCGrad-style per-head backward passes through the same graph (ref:src/trainer_seq2seq.py:1082-1110)
    out = model(inputs=wav, attention_mask=mask, labels=labels.clone())
    shared = [p for p in list(model.encoder.parameters()) + list(model.separator.parameters()) if p.requires_grad]
    gs = [torch.autograd.grad(h.mean(), shared, retain_graph=True, allow_unused=True) for h in out.ctc_per_head]
    assert any(x is not None for x in gs[0]) and any(x is not None for x in gs[1])
    model.ctc_bridge, model.ctc_bridge_type = True, "ctcprompt"
    with torch.no_grad():
        a = model.generate(wav, attention_mask=mask, max_new_tokens=4, recompute=True)
        b = model.generate(wav, attention_mask=mask, max_new_tokens=4)
    assert a.shape == b.shape == (wav.shape[0], 5)
    assert (a == b).float().mean().item() > 0.8          # identical speech context; only the decoder's cached vs full attention differs
