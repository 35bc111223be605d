"""`HybridLoss`: the loss object the composite model calls once per step (drop-in for ref:models/losses.py:135-370).

What the path needs from it is small: turn the frame mask into per-utterance lengths, run head `i` on separator stream `i`
and speaker-`i` targets, average.  The per-head work -- vocabulary projection, log-softmax and the CTC lattice, fused --
lives in `ctc.CTC` / `ops.CTCHeadFn`; nothing here touches a (B, T, V) tensor.

Interface kept from the reference because its callers depend on it:
  * constructor keywords and `forward` keywords (ref:models/losses.py:136-200; the composite model passes everything by
    keyword, ref:...llama.py:772-789);
  * `mode` in {"attention", "ctc", "hybrid"}; hybrid = alpha * CE + (1 - alpha) * CTC (ref:345-353);
  * the CTC term is mean over heads of mean over the batch of each head's per-utterance values (ref:283-293);
  * side effect `last_ctc_per_head`: list of N (B,) tensors that still carry their graph -- PCGrad back-propagates each of
    them separately (ref:src/trainer_seq2seq.py:1082-1110); None in attention mode;
  * `perm_mode` / `build_perm` (ref:8-26): the head permutation policy (None in every reference run);
  * the optional blank check raises AssertionError when a target id collides with the blank (ref:246-259) -- here with one
    host read for all heads instead of two per head.
The PIT branch of the reference is unreachable (`do_pit = False`, ref:240) and is not reproduced; the `use_pit` / `pit_*`
keywords are accepted and stored only.  The decoder cross-entropy belongs to the LLM side (outside this path, SURVEY 2
#11) and is plain `torch.nn.CrossEntropyLoss`.
"""
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

_MODES = ("attention", "ctc", "hybrid")


def build_perm(N: int, mode: Optional[str], step: int, rotate_every: int) -> List[int]:
    """Which separator stream feeds head i (ref:models/losses.py:8-26)."""
    ident = list(range(N))
    if mode is None:
        return ident
    if mode == "reverse":
        return ident[::-1]
    if mode == "swap01":
        if N < 2:
            raise AssertionError("swap01 needs at least two heads")
        return [1, 0] + ident[2:]
    if mode == "rotate":
        shift = (step // max(1, rotate_every)) % N
        return ident[shift:] + ident[:shift]
    raise ValueError(f"Unknown perm_mode: {mode}")


class HybridLoss(nn.Module):
    def __init__(self, alpha: float = 0.7, mode: str = "hybrid", blank_id: Optional[int] = None,
                 enable_blank_check: bool = False, log_every_steps: int = 0, rotate_every: int = 100,
                 use_pit: bool = False, pit_until: int = 1_000, pit_every: int = 1, pit_max_perms: Optional[int] = None):
        super().__init__()
        if mode not in _MODES:
            raise AssertionError("mode must be 'attention', 'ctc', or 'hybrid'")
        self.alpha, self.mode = alpha, mode
        self.blank_id, self.enable_blank_check = blank_id, enable_blank_check
        self.log_every_steps, self.rotate_every = int(log_every_steps), rotate_every
        self.use_pit, self.pit_until, self.pit_every, self.pit_max_perms = use_pit, pit_until, pit_every, pit_max_perms
        self.perm_mode: Optional[str] = None
        self.ce_loss = nn.CrossEntropyLoss()
        self.log_dict = {}
        self.last_ctc_per_head: Optional[List[torch.Tensor]] = None

    # ------------------------------------------------------------------------------------------------ the two terms
    def _attention_term(self, decoder_outputs, labels, vocab, return_dict):
        if decoder_outputs is None or labels is None or vocab is None:
            raise ValueError("decoder_outputs, labels, decoder_vocab_size must be provided for attention loss")
        logits = decoder_outputs.logits if return_dict else decoder_outputs[0]
        return self.ce_loss(logits.reshape(-1, vocab), labels.reshape(-1))

    def _check_targets_below_blank(self, targets: Sequence[torch.Tensor], lengths: Sequence[torch.Tensor]) -> None:
        with torch.no_grad():
            stats = torch.stack([torch.stack([(y.max() if y.numel() else y.new_tensor(-1)).to(torch.int64), n.sum().to(torch.int64)])
                                 for y, n in zip(targets, lengths)]).tolist()             # ONE device->host read
        for i, (top, total) in enumerate(stats):
            if total > 0 and top >= self.blank_id:
                raise AssertionError(f"[CTC blank check] head {i}: target id {top} >= blank_id {self.blank_id}")

    def _ctc_term(self, heads, streams, frame_mask, targets, lengths, n_heads):
        if any(v is None for v in (heads, streams, frame_mask, targets, lengths)):
            raise ValueError("CTC related inputs must be provided for CTC loss")
        N = int(n_heads)
        sizes = (len(heads), len(streams), len(targets), len(lengths))
        if sizes != (N, N, N, N):
            raise AssertionError(f"heads/streams/labels/lengths = {sizes} do not all match talker_numbers={N}")
        frames = frame_mask.sum(dim=1).long()
        B = frames.numel()
        for i in range(N):
            if not (streams[i].size(0) == targets[i].size(0) == lengths[i].size(0) == B):
                raise AssertionError(f"batch dim mismatch @head {i}")
            if lengths[i].dtype not in (torch.int32, torch.int64):
                raise AssertionError(f"length dtype must be int @head {i}")
        step = int(getattr(self, "global_step", 0))
        if self.enable_blank_check and self.blank_id is not None and step % max(1, self.log_every_steps or 1000) == 0:
            self._check_targets_below_blank(targets, lengths)
        order = build_perm(N, self.perm_mode, step=step, rotate_every=self.rotate_every)
        per_head = []
        for head, src in zip(heads, order):
            # the heads run in fp32-class arithmetic whatever the autocast state (the reference switches autocast off and
            # casts to float around this call, ref:models/losses.py:265-268)
            value = head(streams[src].float(), frames, targets[src], lengths[src])
            per_head.append(value.reshape(-1) if value.dim() else value.expand(B))
        return torch.stack([v.mean() for v in per_head]).mean(), per_head

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, decoder_outputs=None, labels=None, decoder_vocab_size=None, talker_ctc=None, sep_hidden_states=None,
                encoder_attention_mask_ctc=None, label_spks=None, label_spks_lengths=None, cross_att_layer_gate=None,
                cross_att_layer_gate_loss=None, cross_att_layer_gate_ratio=0.8, talker_numbers=1, shared_params=None,
                return_dict=True):
        want_attn, want_ctc = self.mode != "ctc", self.mode != "attention"
        attn = self._attention_term(decoder_outputs, labels, decoder_vocab_size, return_dict) if want_attn else None
        ctc, per_head = (self._ctc_term(talker_ctc, sep_hidden_states, encoder_attention_mask_ctc, label_spks, label_spks_lengths,
                                        talker_numbers) if want_ctc else (None, None))
        self.last_ctc_per_head = per_head
        if not want_ctc:
            return attn
        if not want_attn:
            return ctc
        return self.alpha * attn + (1.0 - self.alpha) * ctc
