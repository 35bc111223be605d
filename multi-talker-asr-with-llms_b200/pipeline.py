"""The hot-path slice of `SpeechEncoderDecoderModelLlama` as one module: encoder -> separator -> frame mask ->
N CTC heads -> HybridLoss(mode='ctc'), plus the CTC-only greedy decode.

It restates exactly the wiring of ref:models/modeling_speech_encoder_decoder_llama.py:548-566 (encoder call and
positional outputs), :579-585 (frame-rate mask via `_get_feature_vector_attention_mask_x0`), :686-694 (label split
at <sc>), :772-789 (loss call) and :873-900 (`forward_ctc`), with the same sub-module names (`encoder`, `separator`,
`serialized_ctc`) so reference checkpoints / `extract_sep_ctc` / `load_sep_ctc_from_partial` key prefixes line up.
The LLaMA decoder side of that class is outside this path.
"""
from typing import List, Optional

import torch
import torch.nn as nn
from transformers.models.wavlm.configuration_wavlm import WavLMConfig

from .ctc import CTC
from .greedy import forward_ctc as _forward_ctc
from .greedy import split_k_speakers_and_lengths
from .losses import HybridLoss
from .modeling_wavlm import WavLMModel
from .separator import Separator


class SerializedCTCPath(nn.Module):
    def __init__(self, config: WavLMConfig, talker_numbers: int = 2, separator_hidden: int = 896, vocab_size: int = 128258,
                 pad_token_id: int = 128257, sc_token_id: Optional[int] = None, ctc_alpha: float = 0.7):
        super().__init__()
        self.talker_numbers = int(talker_numbers)
        self.encoder = WavLMModel(config)
        self.separator = Separator(config.hidden_size, separator_hidden, self.talker_numbers)
        odim = vocab_size + 1                                    # ref ...llama.py:187-193: blank = last id
        self.serialized_ctc = nn.ModuleList(CTC(odim, config.hidden_size) for _ in range(self.talker_numbers))
        self.losses = HybridLoss(alpha=ctc_alpha, mode="ctc", blank_id=odim - 1)
        self.ctc_blank_id = odim                                 # the reference passes blank_id = ctc_blank_id - 1
        self.pad_token_id = pad_token_id
        self.sc_token_id = sc_token_id
        self.ctc_per_head: Optional[List[torch.Tensor]] = None

    def encode(self, input_values, attention_mask=None):
        out = self.encoder(input_values, attention_mask=attention_mask)
        enc_h = out[1]                                           # encoder_hidden_state (T frames) feeds the separator
        sep = self.separator(enc_h)
        if attention_mask is not None:
            fmask = self.encoder._get_feature_vector_attention_mask_x0(enc_h.shape[1], attention_mask)
        else:
            fmask = torch.ones(enc_h.shape[:2], dtype=torch.bool, device=enc_h.device)
        return out, sep, fmask

    def forward(self, input_values, attention_mask=None, labels=None, label_spks=None, label_spks_lengths=None):
        """Serialized-CTC training loss.  Give either SOT `labels` (B,L) containing <sc> separators (split here like
        the reference does) or the already split per-speaker `label_spks` / `label_spks_lengths`."""
        _, sep, fmask = self.encode(input_values, attention_mask)
        if label_spks is None:
            if labels is None or self.sc_token_id is None:
                raise ValueError("need labels + sc_token_id, or label_spks + label_spks_lengths")
            label_spks, label_spks_lengths = split_k_speakers_and_lengths(
                labels, self.talker_numbers, self.sc_token_id, self.pad_token_id, ignore_id=-100,
                end_token_id=self.pad_token_id, allow_empty_segment=False)
        loss = self.losses(talker_ctc=self.serialized_ctc, sep_hidden_states=sep, encoder_attention_mask_ctc=fmask,
                           label_spks=label_spks, label_spks_lengths=label_spks_lengths,
                           talker_numbers=self.talker_numbers)
        self.ctc_per_head = self.losses.last_ctc_per_head
        return loss

    def release_graph(self) -> None:
        """Drop the references that keep the last step's autograd graph alive (the per-head losses kept for PCGrad)."""
        self.ctc_per_head = None
        self.losses.last_ctc_per_head = None

    @torch.no_grad()
    def forward_ctc(self, input_values, attention_mask=None) -> torch.Tensor:
        out = self.encoder(input_values, attention_mask=attention_mask)
        return _forward_ctc(out[1], self.separator, self.serialized_ctc, blank_id=self.ctc_blank_id - 1,
                            pad_id=self.pad_token_id)
