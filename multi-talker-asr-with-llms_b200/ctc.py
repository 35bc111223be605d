"""Drop-in `CTC` head for ref:models/ctc.py (ESPnet-derived): same constructor / forward / helper methods.

`forward` never builds the (B,T,V) logits: the vocabulary projection, log-softmax and CTC lattice are fused
(ops.CTCHeadFn).  V is the LLM vocabulary + 1 (~128k), so the reference's path writes and re-reads an 8 GB fp32
tensor per head several times; here only W (V x D bf16) and H are read.
"""
import logging
from typing import Optional

import torch
import torch.nn.functional as F

from . import ops
from . import precise


class CTC(torch.nn.Module):
    def __init__(self, odim: int, encoder_output_size: int, dropout_rate: float = 0.0, ctc_type: str = "builtin",
                 reduce: bool = True, ignore_nan_grad: Optional[bool] = None, zero_infinity: bool = True,
                 brctc_risk_strategy: str = "exp", brctc_group_strategy: str = "end", brctc_risk_factor: float = 0.0):
        super().__init__()
        eprojs = encoder_output_size
        self.dropout_rate = dropout_rate
        self.ctc_lo = torch.nn.Linear(eprojs, odim)
        self.ctc_type = ctc_type
        if ignore_nan_grad is not None:
            zero_infinity = ignore_nan_grad
        if not zero_infinity:
            raise NotImplementedError("mtasr_b200.CTC implements zero_infinity=True only (the reference default)")
        # kept for attribute compatibility (`.ctc_loss.blank`); the arithmetic does not go through it
        self.ctc_loss = torch.nn.CTCLoss(reduction="none", zero_infinity=zero_infinity, blank=odim - 1)
        self.reduce = reduce

    def per_utterance_nll(self, hs_pad, hlens, ys_pad, ys_lens) -> torch.Tensor:
        """(B,) negative log-likelihoods, infeasible rows zeroed -- torch.nn.CTCLoss(reduction='none', zero_infinity=True)."""
        ys_lens = torch.as_tensor(ys_lens, device=hs_pad.device)
        hlens = torch.as_tensor(hlens, device=hs_pad.device)
        Lmax = int(ys_pad.shape[1])
        fn = precise.CTCHeadF32Fn if precise.get_precision() == "fp32" else ops.CTCHeadFn
        return fn.apply(hs_pad, self.ctc_lo.weight, self.ctc_lo.bias, hlens, ys_pad[:, :Lmax].to(torch.int64),
                                   ys_lens, self.ctc_loss.blank)

    def forward(self, hs_pad, hlens, ys_pad, ys_lens):
        """hs_pad (B,Tmax,D), hlens (B), ys_pad (B,Lmax), ys_lens (B) -> loss (ref:models/ctc.py:129-160)."""
        if self.ctc_type != "builtin":
            raise NotImplementedError(f"mtasr_b200.CTC supports ctc_type='builtin' only (got {self.ctc_type!r}); the other "
                                      "reference branches are unreachable or broken (SURVEY 8a/a13)")
        hs = F.dropout(hs_pad, p=self.dropout_rate)          # NB: reference applies it with training=True regardless of mode
        nll = self.per_utterance_nll(hs, hlens, ys_pad, ys_lens)
        size = hs_pad.size(0)
        loss = nll.sum() / size if self.reduce else nll / size
        return loss.to(device=hs_pad.device, dtype=hs_pad.dtype)

    def softmax(self, hs_pad):
        return F.softmax(self.logits(hs_pad), dim=2)

    def log_softmax(self, hs_pad):
        return F.log_softmax(self.logits(hs_pad), dim=2)

    def argmax(self, hs_pad):
        """(B,Tmax) int64 argmax over the vocabulary, fused into the projection epilogue (no logits tensor)."""
        if precise.get_precision() == "fp32":
            return precise.ctc_head_argmax(hs_pad, self.ctc_lo.weight, self.ctc_lo.bias)
        return ops.ctc_head_argmax(hs_pad, self.ctc_lo.weight, self.ctc_lo.bias)

    def logits(self, hs_pad):
        if precise.get_precision() == "fp32":
            return precise.ctc_head_logits(hs_pad, self.ctc_lo.weight, self.ctc_lo.bias)
        return ops.ctc_head_logits(hs_pad, self.ctc_lo.weight, self.ctc_lo.bias)
