"""Drop-in `HybridLoss` for ref:models/losses.py:135-370 (same constructor, forward keywords and side effects).

The CTC branch (ref:models/losses.py:213-293) calls the B200-native `CTC` heads, whose forward fuses the vocabulary
projection, log-softmax and the alpha/beta lattice (ops.CTCHeadFn).  The attention branch is the decoder's
cross-entropy (LLM side, outside this path) and stays `torch.nn.CrossEntropyLoss`.  The reference's
PIT branch is unreachable (`do_pit = False`, ref:models/losses.py:240) and is not reproduced.
"""
from typing import List, Optional

import torch
import torch.nn as nn


def build_perm(N: int, mode: Optional[str], step: int, rotate_every: int) -> List[int]:
    """Head permutation policy of ref:models/losses.py:8-26 (`perm_mode` is None in every reference run)."""
    base = list(range(N))
    if mode is None:
        return base
    if mode == "swap01":
        assert N >= 2
        base[0], base[1] = base[1], base[0]
        return base
    if mode == "reverse":
        return base[::-1]
    if mode == "rotate":
        k = (step // max(1, rotate_every)) % N
        return base[k:] + base[:k]
    raise ValueError(f"Unknown perm_mode: {mode}")


class HybridLoss(nn.Module):
    def __init__(self, alpha: float = 0.7, mode: str = "hybrid", blank_id: Optional[int] = None,
                 enable_blank_check: bool = False, log_every_steps: int = 0, rotate_every: int = 100,
                 use_pit: bool = False, pit_until: int = 1_000, pit_every: int = 1, pit_max_perms: Optional[int] = None):
        super().__init__()
        assert mode in ("attention", "ctc", "hybrid"), "mode must be 'attention', 'ctc', or 'hybrid'"
        self.alpha = alpha
        self.mode = mode
        self.ce_loss = nn.CrossEntropyLoss()
        self.perm_mode = None
        self.rotate_every = rotate_every
        self.blank_id = blank_id
        self.enable_blank_check = enable_blank_check
        self.log_every_steps = int(log_every_steps)
        self.log_dict = {}
        self.use_pit = use_pit
        self.pit_until = pit_until
        self.pit_every = pit_every
        self.pit_max_perms = pit_max_perms
        self.last_ctc_per_head = None

    def forward(self, decoder_outputs=None, labels=None, decoder_vocab_size=None, talker_ctc=None, sep_hidden_states=None,
                encoder_attention_mask_ctc=None, label_spks=None, label_spks_lengths=None, cross_att_layer_gate=None,
                cross_att_layer_gate_loss=None, cross_att_layer_gate_ratio=0.8, talker_numbers=1, shared_params=None,
                return_dict=True):
        loss_attn = 0.0
        loss_ctc = 0.0
        ctc_per_head = None

        if self.mode in ("attention", "hybrid"):
            if decoder_outputs is None or labels is None or decoder_vocab_size is None:
                raise ValueError("decoder_outputs, labels, decoder_vocab_size must be provided for attention loss")
            logits = decoder_outputs.logits if return_dict else decoder_outputs[0]
            loss_attn = self.ce_loss(logits.reshape(-1, decoder_vocab_size), labels.reshape(-1))

        if self.mode in ("ctc", "hybrid"):
            if (talker_ctc is None or sep_hidden_states is None or encoder_attention_mask_ctc is None
                    or label_spks is None or label_spks_lengths is None):
                raise ValueError("CTC related inputs must be provided for CTC loss")
            N = int(talker_numbers)
            assert len(talker_ctc) == N, f"len(talker_ctc)={len(talker_ctc)} != talker_numbers={N}"
            assert len(sep_hidden_states) == len(label_spks) == len(label_spks_lengths) == N, \
                "Mismatch among heads/labels/lengths"
            hlens = encoder_attention_mask_ctc.sum(dim=1).long()
            B = hlens.size(0)
            for i in range(N):
                x, y, yl = sep_hidden_states[i], label_spks[i], label_spks_lengths[i]
                assert x.size(0) == y.size(0) == yl.size(0) == B, f"batch dim mismatch @head {i}"
                assert yl.dtype in (torch.int32, torch.int64), f"length dtype must be int @head {i}"

            step = int(getattr(self, "global_step", 0))
            if self.enable_blank_check and self.blank_id is not None and step % max(1, self.log_every_steps or 1000) == 0:
                with torch.no_grad():   # one host sync for all heads instead of the reference's 2N `.item()`s
                    mx = torch.stack([y.max() if y.numel() else y.new_tensor(-1) for y in label_spks]).tolist()
                    tot = torch.stack([yl.sum() for yl in label_spks_lengths]).tolist()
                for i in range(N):
                    if tot[i] > 0:
                        assert int(mx[i]) < self.blank_id, \
                            f"[CTC blank check] head {i}: target id {int(mx[i])} >= blank_id {self.blank_id}"

            perm = build_perm(N, self.perm_mode, step=step, rotate_every=self.rotate_every)
            sep_hidden_states = [sep_hidden_states[j] for j in perm]
            label_spks = [label_spks[j] for j in perm]
            label_spks_lengths = [label_spks_lengths[j] for j in perm]
            ctc_per_head = []
            for i, head in enumerate(talker_ctc):
                li = head(sep_hidden_states[i].float(), hlens, label_spks[i], label_spks_lengths[i])
                if li.dim() == 0:
                    li = li.unsqueeze(0).expand(B)
                elif li.dim() > 1:
                    li = li.reshape(-1)
                ctc_per_head.append(li)
            loss_ctc = torch.stack([l.mean() for l in ctc_per_head]).mean()

        if self.mode == "attention":
            total = loss_attn
            self.last_ctc_per_head = None
        elif self.mode == "ctc":
            total = loss_ctc
            self.last_ctc_per_head = ctc_per_head
        else:
            total = self.alpha * loss_attn + (1.0 - self.alpha) * loss_ctc
            self.last_ctc_per_head = ctc_per_head
        return total
