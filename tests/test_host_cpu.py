"""Host-side logic of the product package that needs no GPU: integer helpers, masks, parameter names, loss wiring."""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, load_model_golden, sub_state


def test_split_and_prefix_match_reference_golden():
    from mtasr_b200 import greedy
    g = np.load(os.path.join(GOLDEN, "host_small.npz"))
    sc, pad = int(g["split_sc"]), int(g["split_pad"])
    labs, lens = greedy.split_k_speakers_and_lengths(torch.from_numpy(g["split_labels"]), 2, sc, pad, ignore_id=-100,
                                                     end_token_id=pad, allow_empty_segment=False)
    assert [l.tolist() for l in labs] == [g["split0"].tolist(), g["split1"].tolist()]
    assert [l.tolist() for l in lens] == [g["split_len0"].tolist(), g["split_len1"].tolist()]
    assert labs[0].dtype == torch.int64 and lens[0].dtype == torch.int64

    class Dec(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(32, 4)

        def get_input_embeddings(self):
            return self.emb

    dec = Dec()
    emb, mask, ids = greedy.build_multi_ctc_prefix_from_heads([torch.from_numpy(g["prefix_h0"]), torch.from_numpy(g["prefix_h1"])],
                                                              dec, pad, None)
    assert ids.tolist() == g["prefix_ids"].tolist() and mask.tolist() == g["prefix_mask"].tolist()
    assert torch.equal(emb, dec.emb(ids))
    with pytest.raises(RuntimeError):        # a sample with no token from any head: the reference's torch.cat([]) raises
        greedy.build_multi_ctc_prefix_from_heads([torch.tensor([[1, pad], [pad, pad]]), torch.tensor([[2], [pad]])], dec, pad, None)


def test_split_errors_and_random_against_oracle():
    from mtasr_b200 import greedy
    from oracle import host_ref
    with pytest.raises(ValueError):          # wrong separator count
        greedy.split_k_speakers_and_lengths(torch.tensor([[1, 2, 3]]), 2, 9, 0)
    with pytest.raises(ValueError):          # empty segment
        greedy.split_k_speakers_and_lengths(torch.tensor([[9, 1, 2]]), 2, 9, 0, allow_empty_segment=False)
    rs = np.random.RandomState(0)
    for K in (2, 3):
        rows = []
        for _ in range(17):
            segs = [rs.randint(1, 8, size=rs.randint(1, 6)).tolist() for _ in range(K)]
            row = []
            for i, s in enumerate(segs):
                row += s + ([9] if i < K - 1 else [])
            rows.append(row + [0] * rs.randint(0, 3))
        L = max(len(r) for r in rows)
        mat = torch.tensor([r + [-100] * (L - len(r)) for r in rows])
        labs, lens = greedy.split_k_speakers_and_lengths(mat, K, 9, 0, ignore_id=-100, end_token_id=0, allow_empty_segment=False)
        olabs, olens = host_ref.split_labels(mat.tolist(), K, 9, 0, -100, 0, False)
        assert [l.tolist() for l in labs] == olabs and [l.tolist() for l in lens] == olens


def test_state_dict_names_and_masks_match_reference():
    """Reference checkpoints must load: the fixture's parameter names come from the reference's own modules."""
    from mtasr_b200.ctc import CTC
    from mtasr_b200.modeling_wavlm import WavLMModel, relpos_bucket
    from mtasr_b200.separator import Separator
    from oracle.model_ref import make_config
    for kind in ("tiny_large", "tiny_base"):
        g, params, _ = load_model_golden(kind)
        cfg = make_config(kind)
        enc = WavLMModel(cfg)
        enc.load_state_dict(sub_state(params, "encoder."), strict=True)
        sep = Separator(cfg.hidden_size, int(g["hidden_sep"]), int(g["n_spk"]))
        sep.load_state_dict(sub_state(params, "separator."), strict=True)
        head = CTC(int(g["vocab"]), cfg.hidden_size)
        head.load_state_dict(sub_state(params, "serialized_ctc.0."), strict=True)
        assert head.ctc_loss.blank == int(g["vocab"]) - 1 and head.ctc_lo.weight.shape == (int(g["vocab"]), cfg.hidden_size)
        mask = torch.from_numpy(g["mask"])
        T = g["enc"].shape[1]
        assert torch.equal(enc._get_feature_vector_attention_mask_x0(T, mask), torch.from_numpy(g["frame_mask"]))
        assert enc._get_feature_vector_attention_mask_x4(g["down"].shape[1], mask).shape == g["down"].shape[:2]
        assert enc.main_input_name == "input_values" and enc.get_output_embeddings() is None
        with pytest.raises(Exception):      # the product path never silently runs on the CPU
            enc(torch.from_numpy(g["wav"]), attention_mask=mask)
    # T5-style bucket of hf:253-271 (restated in the survey): exact below 80, log-spaced to 800, 160 per sign
    rel = torch.arange(-1200, 1201)
    b = relpos_bucket(rel, 320, 800)
    assert b.min().item() == 0 and b.max().item() == 319
    assert b[1200].item() == 0 and b[1200 + 5].item() == 160 + 5 and b[1200 - 79].item() == 79
    assert b[0].item() == 159 and b[-1].item() == 319
    from transformers.models.wavlm.modeling_wavlm import WavLMAttention
    att = WavLMAttention(embed_dim=128, num_heads=2, num_buckets=320, max_distance=800, has_relative_position_bias=True)
    q = torch.arange(50)[:, None]
    k = torch.arange(50)[None, :]
    assert torch.equal(att._relative_positions_bucket(k - q), relpos_bucket(k - q, 320, 800))


def test_hybrid_loss_wiring_and_errors():
    from mtasr_b200.losses import HybridLoss, build_perm
    assert build_perm(3, None, 0, 100) == [0, 1, 2]
    assert build_perm(3, "swap01", 0, 100) == [1, 0, 2]
    assert build_perm(3, "reverse", 0, 100) == [2, 1, 0]
    assert build_perm(3, "rotate", 250, 100) == [2, 0, 1]
    with pytest.raises(ValueError):
        build_perm(2, "bogus", 0, 1)
    with pytest.raises(AssertionError):
        HybridLoss(mode="nope")
    loss = HybridLoss(mode="ctc", blank_id=10)
    with pytest.raises(ValueError):
        loss(talker_ctc=None)

    class FakeHead(torch.nn.Module):           # scalar-returning head like CTC(reduce=True)
        def forward(self, hs, hlens, ys, yl):
            return hs.sum() * 0 + float(yl.sum())

    B, T, D = 3, 5, 4
    heads = [FakeHead(), FakeHead()]
    seps = [torch.zeros(B, T, D), torch.zeros(B, T, D)]
    fm = torch.ones(B, T, dtype=torch.bool)
    ys = [torch.ones(B, 2, dtype=torch.long), torch.ones(B, 3, dtype=torch.long)]
    yl = [torch.tensor([2, 2, 2]), torch.tensor([3, 3, 1])]
    out = loss(talker_ctc=heads, sep_hidden_states=seps, encoder_attention_mask_ctc=fm, label_spks=ys, label_spks_lengths=yl,
               talker_numbers=2)
    assert out.item() == pytest.approx((6 + 7) / 2)
    assert len(loss.last_ctc_per_head) == 2 and loss.last_ctc_per_head[0].shape == (B,)
    with pytest.raises(AssertionError):
        loss(talker_ctc=heads, sep_hidden_states=seps, encoder_attention_mask_ctc=fm, label_spks=ys, label_spks_lengths=yl,
             talker_numbers=3)
    chk = HybridLoss(mode="ctc", blank_id=1, enable_blank_check=True)
    with pytest.raises(AssertionError):        # target id >= blank id
        chk(talker_ctc=heads, sep_hidden_states=seps, encoder_attention_mask_ctc=fm, label_spks=ys, label_spks_lengths=yl,
            talker_numbers=2)
    hyb = HybridLoss(alpha=0.7, mode="hybrid")

    class DO:
        logits = torch.zeros(B, 2, 6)
    v = hyb(decoder_outputs=DO(), labels=torch.zeros(B, 2, dtype=torch.long), decoder_vocab_size=6, talker_ctc=heads,
            sep_hidden_states=seps, encoder_attention_mask_ctc=fm, label_spks=ys, label_spks_lengths=yl, talker_numbers=2)
    assert v.item() == pytest.approx(0.7 * np.log(6) + 0.3 * 6.5, rel=1e-5)


def test_algorithmic_flops_match_survey():
    from mtasr_b200.configs import V_LLAMA3_CTC, algorithmic_flops, wavlm_config
    r = algorithmic_flops(wavlm_config("large"), 160000, 2, 896, V_LLAMA3_CTC, adapter_backward=True)
    assert r["frames"] == 499
    assert abs(r["fwd"] / 1e9 - 668.9) < 1.0 and abs(r["total"] / 1e9 - 1908.4) < 2.0     # SURVEY 8d / BASELINE.md 3
    r3 = algorithmic_flops(wavlm_config("large"), 240000, 3, 896, V_LLAMA3_CTC, adapter_backward=True)
    assert r3["frames"] == 749 and abs(r3["fwd"] / 1e9 - 1221.6) < 2.0


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_spec_augment_matches_reference_golden(case):
    """`_mask_hidden_states` (ref:models/modeling_wavlm.py:358-402) is host-side numpy RNG + a boolean scatter and
    stays in Python in the product; from the same numpy seed it must reproduce the REFERENCE's masked tensor bit for bit
    (fixture written by oracle/gen_golden_specaug.py running the reference's own method), in training mode with sampled
    spans (time and feature axis), with explicit `mask_time_indices`, and be the identity in eval mode."""
    from oracle.model_ref import make_config
    from mtasr_b200.modeling_wavlm import WavLMModel
    g = np.load(os.path.join(GOLDEN, "spec_augment.npz"))
    p, ml, mm, pf, seed = g[f"{case}_cfg"].tolist()
    cfg = make_config("tiny_large", hidden_size=32, output_hidden_size=32, intermediate_size=64, mask_time_prob=p,
                      mask_time_length=int(ml), mask_time_min_masks=int(mm), mask_feature_prob=pf, mask_feature_length=8,
                      mask_feature_min_masks=1)
    model = WavLMModel(cfg)
    with torch.no_grad():
        model.masked_spec_embed.copy_(torch.from_numpy(g[f"{case}_embed"]))
    h = torch.from_numpy(g[f"{case}_h"])
    fm = torch.from_numpy(g[f"{case}_fm"])
    h0 = h.clone()
    model.train()
    np.random.seed(int(seed))
    out = model._mask_hidden_states(h, attention_mask=fm)
    assert torch.equal(out, torch.from_numpy(g[f"{case}_out"]))
    assert torch.equal(h, h0)                                   # the product never writes into the caller's tensor
    masked = (out != h0).any(-1)
    assert masked.any()
    if pf == 0.0:
        assert not (masked & ~fm).any()                         # time spans stay inside the valid frames
    if f"{case}_out_given" in g.files:
        given = torch.from_numpy(g[f"{case}_given"])
        out_g = model._mask_hidden_states(h, mask_time_indices=given)
        assert torch.equal(out_g, torch.from_numpy(g[f"{case}_out_given"]))
        assert torch.equal(out_g[given], model.masked_spec_embed.detach().expand(int(given.sum()), -1))
    model.eval()
    assert torch.equal(model._mask_hidden_states(h, attention_mask=fm), h0)
    cfg.apply_spec_augment = False
    model.train()
    assert torch.equal(model._mask_hidden_states(h, attention_mask=fm), h0)


def test_composite_state_dict_matches_reference_and_unsupported_options_raise():
    """Row f1 on the host: the B200 composite has exactly the reference composite's parameter names and shapes (fixture written
    by the reference class itself), and LLM-side options outside the hot path are refused loudly."""
    from oracle.model_ref import make_composite_config
    from mtasr_b200.composite import SpeechEncoderDecoderModelLlama, shift_tokens_right
    g = np.load(os.path.join(GOLDEN, "composite_tiny.npz"))
    cfg = make_composite_config(int(g["meta"][0]))
    model = SpeechEncoderDecoderModelLlama(cfg)
    sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("p:")}
    assert set(sd) == set(model.state_dict())
    model.load_state_dict(sd, strict=True)
    assert model.ctc_blank_id == int(g["meta"][0]) + 3 and model.serialized_ctc[0].ctc_loss.blank == int(g["meta"][0]) + 2
    assert shift_tokens_right(torch.tensor([[5, 6, -100]]), 41, 1).tolist() == [[1, 5, 6]]
    for opt in ({"instruct": True}, {"decoder_cross_attention": True}, {"ctc_bridge": True, "ctc_bridge_type": "raw"}):
        with pytest.raises(NotImplementedError):
            SpeechEncoderDecoderModelLlama(make_composite_config(40, **opt))
    with pytest.raises(Exception):           # the product path has no CPU fallback
        model(inputs=torch.zeros(1, 4000), labels=torch.tensor([[3, 40, 4]]))


def test_vocabulary_row_pitch_is_line_aligned():
    """The (rows, V) logits / softmax matrices of the CTC head use a row pitch that is a multiple of 64 two-byte elements
    (every row starts on a 128-byte line) and never narrower than V; MTASR_VP_ALIGN is rounded to the
    16-byte stride rule of tensor maps."""
    from mtasr_b200 import ops
    for V in (17, 515, 4099, 128259, 128320):
        p = ops._vocab_pitch(V)
        assert p >= V and p % ops._VP_ALIGN == 0 and p - V < ops._VP_ALIGN
    assert ops._VP_ALIGN % 8 == 0
    if os.environ.get("MTASR_VP_ALIGN") is None:
        assert ops._VP_ALIGN == 64 and ops._vocab_pitch(128259) == 128320


def test_length_arithmetic_closed_form_equals_the_layer_loop():
    """`WavLMModel._length_chain` (one floor division) against the per-layer loop of hf:640-659 / hf:655-657 for every
    input length, the WavLM conv stack with 0..3 adapter steps, and random kernel / stride stacks (negative intermediate
    lengths included: floor semantics)."""
    from mtasr_b200.modeling_wavlm import WavLMModel
    rs = np.random.RandomState(0)
    L = torch.arange(-64, 500000, dtype=torch.int64)
    stacks = [list(zip((10, 3, 3, 3, 3, 2, 2), (5, 2, 2, 2, 2, 2, 2))) + [(1, 2)] * a for a in range(4)]
    stacks += [[(int(rs.randint(1, 13)), int(rs.randint(1, 7))) for _ in range(rs.randint(1, 9))] for _ in range(50)]
    for st in stacks:
        n = L.clone()
        for k, s in st:
            n = torch.div(n - k, s, rounding_mode="floor") + 1
        assert torch.equal(WavLMModel._length_chain(L, st), n), st
