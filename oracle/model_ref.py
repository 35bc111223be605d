"""Torch-CPU restatement of the reference's hot-path MODULE WIRING.  Test infrastructure only.

The arithmetic of the WavLM blocks lives in the third-party ``transformers`` package (5.5.0 in this image;
unpinned by the reference) which IS present on the GPU box, so the oracle instantiates those
classes directly and restates only what the reference adds on top of them:

  RefWavLMModel   ref:models/modeling_wavlm.py:318-465 (two-output adapter ref:223-254, 6-field output ref:71-99)
  RefSeparator    ref:models/separator.py:6-166
  RefCTC          ref:models/ctc.py:23-65,129-193 (ctc_type="builtin" only: the only reachable branch)
  ref_hybrid_ctc  ref:models/losses.py:213-293,345-353 (mode='ctc', identity permutation)

`oracle/gen_golden.py` checks every class here against the real reference modules (same state_dict,
same inputs) before the golden fixtures are written.
"""
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from transformers import WavLMConfig
from transformers.models.wavlm.modeling_wavlm import (
    WavLMAdapterLayer,
    WavLMEncoder,
    WavLMEncoderStableLayerNorm,
    WavLMFeatureEncoder,
    WavLMFeatureProjection,
    WavLMPreTrainedModel,
)


class RefAdapter(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.output_hidden_size != config.hidden_size:
            self.proj = nn.Linear(config.hidden_size, config.output_hidden_size)
            self.proj_layer_norm = nn.LayerNorm(config.output_hidden_size)
        else:
            self.proj = self.proj_layer_norm = None
        self.layers = nn.ModuleList(WavLMAdapterLayer(config) for _ in range(config.num_adapter_layers))
        self.layerdrop = config.layerdrop

    def forward(self, h):
        if self.proj is not None:
            h = self.proj_layer_norm(self.proj(h))
        h = h.transpose(1, 2)
        tap = None
        for i, layer in enumerate(self.layers):
            if not self.training or (np.random.random() > self.layerdrop):
                h = layer(h)
            if i == 1:
                tap = h
        return h.transpose(1, 2), tap.transpose(1, 2)


class RefWavLMModel(WavLMPreTrainedModel):
    def __init__(self, config: WavLMConfig):
        super().__init__(config)
        self.config = config
        self.feature_extractor = WavLMFeatureEncoder(config)
        self.feature_projection = WavLMFeatureProjection(config)
        if config.mask_time_prob > 0.0 or config.mask_feature_prob > 0.0:
            self.masked_spec_embed = nn.Parameter(torch.Tensor(config.hidden_size).uniform_())
        self.encoder = WavLMEncoderStableLayerNorm(config) if config.do_stable_layer_norm else WavLMEncoder(config)
        self.adapter = RefAdapter(config) if config.add_adapter else None
        self.post_init()

    def frame_mask_x0(self, T, attention_mask):
        n = attention_mask.cumsum(dim=-1)[:, -1]
        for k, s in zip(self.config.conv_kernel, self.config.conv_stride):
            n = torch.div(n - k, s, rounding_mode="floor") + 1
        return torch.arange(T, device=n.device)[None, :] < n[:, None]

    def forward(self, input_values, attention_mask=None, mask_time_indices=None):
        feats = self.feature_extractor(input_values).transpose(1, 2)
        if attention_mask is not None:
            attention_mask = self._get_feature_vector_attention_mask(feats.shape[1], attention_mask, add_adapter=False)
        h, feats = self.feature_projection(feats)
        if mask_time_indices is not None:
            h[mask_time_indices] = self.masked_spec_embed.to(h.dtype)
        enc = self.encoder(h, attention_mask=attention_mask, return_dict=True)[0]
        last, down = self.adapter(enc)
        return last, enc, down, feats


class RefLSTMCell(nn.Module):
    def __init__(self, i, h):
        super().__init__()
        self.W = nn.Linear(i + h, 4 * h)


class RefLSTM(nn.Module):
    def __init__(self, hidden, layers):
        super().__init__()
        self.cells = nn.ModuleList(RefLSTMCell(hidden, hidden) for _ in range(layers))


class RefSeparator(nn.Module):
    """Eval-mode (dropout = identity) restatement with the reference's parameter names."""

    def __init__(self, in_dim, hidden, n_spk, num_layers=2, eps=1e-3):
        super().__init__()
        self.pre_proj = nn.Linear(in_dim, hidden)
        self.pre_ln = nn.LayerNorm(hidden)
        self.lstm = RefLSTM(hidden, num_layers)
        self.post_ln = nn.LayerNorm(hidden)
        self.sep_branches = nn.ModuleList(
            nn.Sequential(nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, in_dim), nn.ReLU(), nn.LayerNorm(in_dim))
            for _ in range(n_spk))
        self.hidden = hidden

    def forward(self, x) -> List[torch.Tensor]:
        y = self.pre_ln(F.relu(self.pre_proj(x)))
        B, T, _ = y.shape
        for cell in self.lstm.cells:
            h = y.new_zeros(B, self.hidden)
            c = y.new_zeros(B, self.hidden)
            outs = []
            for t in range(T):
                g = cell.W(torch.cat([y[:, t], h], -1))
                i, f, gg, o = g.chunk(4, -1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                outs.append(h)
            y = torch.stack(outs, 1)
        y = self.post_ln(y)
        return [br(y) for br in self.sep_branches]


class RefCTC(nn.Module):
    def __init__(self, odim, eprojs):
        super().__init__()
        self.ctc_lo = nn.Linear(eprojs, odim)
        self.blank = odim - 1

    def per_utt_nll(self, hs, hlens, ys_pad, ys_lens):
        lp = self.ctc_lo(hs).transpose(0, 1).log_softmax(2).float()
        tgt = torch.cat([ys_pad[i, :l] for i, l in enumerate(ys_lens)])
        return F.ctc_loss(lp, tgt, hlens, ys_lens, blank=self.blank, reduction="none", zero_infinity=True)

    def forward(self, hs, hlens, ys_pad, ys_lens):
        return (self.per_utt_nll(hs, hlens, ys_pad, ys_lens).sum() / hs.size(0)).to(hs.dtype)

    def argmax(self, hs):
        return torch.argmax(self.ctc_lo(hs), dim=2)


def ref_hybrid_ctc(heads, sep_hidden, frame_mask, label_spks, label_lens):
    """mean over heads of (sum_b nll_b / B); also returns the per-head (B,)-expanded list."""
    hlens = frame_mask.sum(1).long()
    B = hlens.numel()
    per = [h(x.float(), hlens, y, yl).unsqueeze(0).expand(B) for h, x, y, yl in zip(heads, sep_hidden, label_spks, label_lens)]
    return torch.stack([p.mean() for p in per]).mean(), per


def make_config(kind: str, **over) -> WavLMConfig:
    """kind: 'large' | 'base_plus' | 'tiny_large' | 'tiny_base' (hub config.json is unreachable offline;
    values follow SURVEY Appendix A and ref:utils/create_from_pretrained.py:202-212)."""
    common = dict(add_adapter=True, feat_proj_dropout=0.0, final_dropout=0.0, layerdrop=0.0, mask_time_prob=0.0,
                  hidden_dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    if kind == "large":
        kw = dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                  feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=False)
    elif kind == "base_plus":
        kw = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                  feat_extract_norm="group", do_stable_layer_norm=False, conv_bias=False)
    elif kind == "tiny_large":
        kw = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                  conv_dim=(64,) * 7, feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=True,
                  num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=2, output_hidden_size=128)
    elif kind == "tiny_base":
        kw = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                  conv_dim=(64,) * 7, feat_extract_norm="group", do_stable_layer_norm=False, conv_bias=False,
                  num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=2, output_hidden_size=128)
    else:
        raise ValueError(kind)
    kw.update(common)
    kw.update(over)
    return WavLMConfig(**kw)


def synth_batch(B: int, S: int, n_spk: int, vocab: int, seed: int = 1234, varlen: bool = False,
                tok_per_sec=(2, 6)):
    """LibriMix-shaped synthetic batch: sum of low-passed noise 'speakers', per-utterance
    zero-mean/unit-variance, right-zero-padded; targets U[0, vocab-3) with pad=vocab-2, blank=vocab-1."""
    g = torch.Generator().manual_seed(seed)
    wav = torch.zeros(B, S)
    for _ in range(n_spk):
        n = torch.randn(B, S, generator=g)
        lp = torch.empty_like(n)
        acc = torch.zeros(B)
        # 1-pole low-pass, vectorised as a cumulative filter in blocks (exact recursion is too slow in Python)
        k = 0.85
        w = k ** torch.arange(64, dtype=torch.float32)
        lp = F.conv1d(F.pad(n[:, None], (63, 0)), w.flip(0)[None, None])[:, 0] * (1 - k)
        wav += lp * (0.5 + 0.5 * torch.rand(B, 1, generator=g))
    lens = torch.full((B,), S, dtype=torch.long)
    if varlen:
        lens = (S * (0.6 + 0.4 * torch.rand(B, generator=g))).long()
        lens[0] = S
    mask = (torch.arange(S)[None] < lens[:, None]).long()
    wav = wav * mask
    mean = wav.sum(1, keepdim=True) / lens[:, None]
    var = (((wav - mean) * mask) ** 2).sum(1, keepdim=True) / lens[:, None]
    wav = (wav - mean) / torch.sqrt(var + 1e-7) * mask
    sec = S / 16000.0
    lo, hi = max(1, int(tok_per_sec[0] * sec)), max(2, int(tok_per_sec[1] * sec))
    pad_id, blank = vocab - 2, vocab - 1
    labels, lab_lens = [], []
    for _ in range(n_spk):
        L = torch.randint(lo, hi + 1, (B,), generator=g)
        y = torch.full((B, int(L.max())), pad_id, dtype=torch.long)
        for b in range(B):
            y[b, : L[b]] = torch.randint(0, vocab - 3, (int(L[b]),), generator=g)
            if L[b] >= 3:
                y[b, 2] = y[b, 1]            # exercise the repeated-label rule
        labels.append(y)
        lab_lens.append(L)
    return wav, mask, labels, lab_lens


def make_composite_config(v_txt: int = 40, separator_hidden: int = 96, n_spk: int = 2, enc_kind: str = "tiny_large", dec_hidden: int = 128,
                          dec_layers: int = 2, dec_heads: int = 2, dec_kv_heads: int = 2, dec_inter: int = 256, **over):
    """SpeechEncoderDecoderConfig the way ref:utils/create_from_pretrained.py:202-270 assembles it (non-instruct): text ids
    0..v_txt-1, <sc> = v_txt, <pad> = v_txt + 1 (decoder vocabulary v_txt + 2; the CTC heads add the blank as the last id)."""
    from transformers import LlamaConfig
    from transformers.models.speech_encoder_decoder.configuration_speech_encoder_decoder import SpeechEncoderDecoderConfig
    enc_cfg = make_config(enc_kind)
    dec_cfg = LlamaConfig(vocab_size=v_txt + 2, hidden_size=dec_hidden, intermediate_size=dec_inter, num_hidden_layers=dec_layers,
                          num_attention_heads=dec_heads, num_key_value_heads=dec_kv_heads, max_position_embeddings=2048,
                          pad_token_id=v_txt + 1, bos_token_id=1, eos_token_id=2)
    dec_cfg.instruct = False
    dec_cfg.cross_attention_hidden_size = None
    dec_cfg.sc_token_id = v_txt
    dec_cfg.ignore_token_id = -100
    cfg = SpeechEncoderDecoderConfig.from_encoder_decoder_configs(enc_cfg, dec_cfg)
    cfg.talker_ctc = True
    cfg.talker_numbers = n_spk
    cfg.separator_hidden = separator_hidden
    cfg.ctc_alpha = 0.7
    cfg.train_mode = "ctc"
    cfg.pad_token_id = v_txt + 1
    cfg.sc_token_id = v_txt
    cfg.ignore_token_id = -100
    cfg.eos_token_id = 2
    cfg.decoder_start_token_id = 1
    cfg.instruct = False
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg
