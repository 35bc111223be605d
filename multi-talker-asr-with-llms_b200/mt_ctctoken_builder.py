"""Drop-in `MultiSpkCTCTokenBuilder` for ref:models/mt_ctctoken_builder.py (same constructor / forward contract).

Token-level acoustic memory from the separator branches and their CTC heads: greedy path -> segments (maximal runs of one
non-blank token, emitted when a blank follows or the valid region ends) -> per-segment mean of the branch features and
confidence 1 - mean(p_blank), concatenated over speakers.  The reference walks (speaker, utterance, frame) in Python with a
host round trip per frame; here the path and the blank posterior come from the fused vocabulary GEMM (no (B,T,V) tensor),
the segmentation is one small kernel and the segment means are one kernel each way.  ONE host read (the segment counts,
which fix the output shape) per speaker.
"""
from typing import List, Tuple

import torch
from torch import Tensor, nn

from . import kernels as K
from . import ops


class MultiSpkCTCTokenBuilder(nn.Module):
    def __init__(self):
        super().__init__()

    @staticmethod
    def _blank_id(ctc_module: nn.Module) -> int:
        if hasattr(ctc_module, "blank_id"):
            return int(ctc_module.blank_id)
        if hasattr(ctc_module, "ctc_loss") and hasattr(ctc_module.ctc_loss, "blank"):
            return int(ctc_module.ctc_loss.blank)
        return int(ctc_module.ctc_lo.weight.shape[0]) - 1

    def _build_one_speaker(self, sep_hidden: Tensor, enc_mask: Tensor, ctc_module: nn.Module) -> Tuple[Tensor, Tensor, Tensor]:
        B, T, D = sep_hidden.shape
        blank = self._blank_id(ctc_module)
        with torch.no_grad():
            path, pblank = ops.ctc_head_path_and_pblank(sep_hidden, ctc_module.ctc_lo.weight, ctc_module.ctc_lo.bias, blank)
            ss, se, n = K.ctc_segments(path, enc_mask.bool(), blank)
            lengths = n.tolist()
        max_L = max(lengths) if lengths else 0
        if max_L == 0:
            return (sep_hidden.new_zeros((B, 0, D)), torch.ones(B, 0, dtype=torch.bool, device=sep_hidden.device),
                    sep_hidden.new_zeros((B, 0)))
        feats, conf = ops.SegmentMeanFn.apply(sep_hidden, pblank, ss, se, n, max_L)
        tok_mask = torch.arange(max_L, device=sep_hidden.device)[None, :] >= n.to(torch.long)[:, None]   # True = padding
        return feats, tok_mask, conf.to(sep_hidden.dtype)

    def forward(self, sep_hidden_list: List[Tensor], encoder_attention_mask_ctc: Tensor,
                ctc_modules: List[nn.Module]) -> Tuple[Tensor, Tensor, Tensor]:
        K_spk = len(sep_hidden_list)
        assert len(ctc_modules) == K_spk, "need one CTC module per separator branch"
        mems, masks, confs = [], [], []
        for k in range(K_spk):
            f, m, c = self._build_one_speaker(sep_hidden_list[k], encoder_attention_mask_ctc, ctc_modules[k])
            mems.append(f)
            masks.append(m)
            confs.append(c)
        return torch.cat(mems, dim=1), torch.cat(masks, dim=1), torch.cat(confs, dim=1)
