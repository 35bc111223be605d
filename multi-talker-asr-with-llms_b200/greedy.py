"""Greedy-CTC consumers of the path: collapse, CTC-only decode, LLM prompt prefix and the label splitter.

  ctc_remove_duplicates_and_blank   ref:models/modeling_speech_encoder_decoder_llama.py:902-972
  forward_ctc                       ref:models/modeling_speech_encoder_decoder_llama.py:873-900
  build_multi_ctc_prefix_from_heads ref:models/ctc_prompt.py:5-120
  split_k_speakers_and_lengths      ref:utils/split_labels_by_sc.py:5-97

The reference walks every row as a Python list on the host (`seq.detach().cpu().tolist()` per row, `.item()` per
sample).  Here the collapse is one warp-per-row stream compaction on the device (csrc/ctc.cu) and the whole call
costs ONE host read-back (the per-row lengths, which the reference API returns as list[int] and which fix the
output shape).  Integer results are bit-exact with the reference.
"""
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import kernels as K


def ctc_remove_duplicates_and_blank(argmax_tensor: torch.Tensor, blank_id: int = 128258, pad_id: int = 128257,
                                    collapse_across_blanks: bool = True) -> Tuple[torch.Tensor, List[int]]:
    """(B,T) argmax ids -> ((B,Lmax) right-padded with pad_id, list of lengths).  Drops pad and blank, then drops a
    token equal to the last KEPT token; both `collapse_across_blanks` settings of the reference reduce to this rule
    (its "classic" branch compares with processed[-1], which is the last kept token too)."""
    if argmax_tensor.dim() != 2:
        raise ValueError("argmax_tensor must be (B, T)")
    out, lens = K.ctc_collapse(argmax_tensor.to(torch.int64), int(blank_id), int(pad_id))
    lengths = lens.tolist()
    lmax = max(lengths) if lengths else 0
    return out[:, :lmax].contiguous(), lengths


def forward_ctc(encoder_hidden_state: torch.Tensor, separator: nn.Module, serialized_ctc: Sequence[nn.Module],
                blank_id: int, pad_id: int) -> torch.Tensor:
    """CTC-only greedy decode: separator -> per-head fused vocab-GEMM argmax -> collapse -> concat over heads."""
    sep = separator(encoder_hidden_state)
    outs = []
    for head, x in zip(serialized_ctc, sep):
        ids, _ = ctc_remove_duplicates_and_blank(head.argmax(x), blank_id=blank_id, pad_id=pad_id)
        outs.append(ids)
    return torch.cat(outs, dim=1)


def build_multi_ctc_prefix_from_heads(ctc_transcription_list: List[torch.Tensor], decoder: nn.Module, pad_id: int,
                                      max_prefix_len_per_head: Optional[int] = 64):
    """Concatenate each sample's non-pad ids over the heads (each head truncated to `max_prefix_len_per_head`),
    right-pad to the batch maximum (at least 1) and look the ids up in the decoder's embedding table.
    Returns (prefix_embeds (B,L,d), prefix_mask (B,L) bool, prefix_ids (B,L) int64)."""
    assert len(ctc_transcription_list) > 0, "ctc_transcription_list must not be empty."
    B = ctc_transcription_list[0].size(0)
    dev = ctc_transcription_list[0].device
    for i, t in enumerate(ctc_transcription_list):
        assert t.size(0) == B, f"CTC head {i} has different batch size: {t.size(0)} vs {B}"
    if hasattr(decoder, "model") and hasattr(decoder.model, "embed_tokens"):
        embed = decoder.model.embed_tokens
    else:
        embed = decoder.get_input_embeddings()
    dec_pad = getattr(getattr(decoder, "config", None), "pad_token_id", None)
    if dec_pad is not None and dec_pad != pad_id:
        print(f"[WARN] pad_id mismatch: decoder.config.pad_token_id={dec_pad}, but function pad_id={pad_id}. "
              f"Using pad_id={pad_id} for prefix_ids.")
    keeps = []
    for ids in ctc_transcription_list:
        keep = ids != pad_id
        if max_prefix_len_per_head is not None:
            keep = keep & (keep.cumsum(1) <= max_prefix_len_per_head)
        keeps.append(keep)
    ids_all = torch.cat([t.to(torch.int64) for t in ctc_transcription_list], dim=1)       # (B, sum L_k)
    keep_all = torch.cat(keeps, dim=1)
    pos = keep_all.cumsum(1) - 1
    lengths = keep_all.sum(1)
    lens_host = lengths.tolist()                                                           # the one host sync
    if ids_all.shape[1] == 0 or min(lens_host) == 0:
        # the reference concatenates an empty list for such a sample (ref:models/ctc_prompt.py:97-104)
        raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
    L = max(1, max(lens_host))
    prefix_ids = torch.full((B, L + 1), pad_id, dtype=torch.long, device=dev)
    prefix_ids.scatter_(1, torch.where(keep_all, pos, torch.full_like(pos, L)), ids_all)  # dropped ids -> spill column
    prefix_ids = prefix_ids[:, :L].contiguous()
    prefix_mask = torch.arange(L, device=dev)[None, :] < lengths[:, None]
    prefix_ids = torch.where(prefix_mask, prefix_ids, torch.full_like(prefix_ids, pad_id))
    return embed(prefix_ids), prefix_mask, prefix_ids


@torch.no_grad()
def split_k_speakers_and_lengths(labels: torch.Tensor, k_speakers: int, sep_id: int, pad_token_id: int,
                                 ignore_id: Optional[int] = -100, end_token_id: Optional[int] = -100,
                                 allow_empty_segment: bool = True):
    """Split SOT label rows at `sep_id` into K per-speaker right-padded targets + lengths.

    CUDA labels (the training path: the collator's labels are on the device by the time the composite model splits them)
    never leave the device: one thread-per-row kernel does the cut / split / trim (csrc/ctc.cu split_labels_kernel, SURVEY
    row f3) and the only host traffic is ONE read of K+3 integers -- the per-head maximum lengths, which fix the output
    shapes the reference API promises, and the error word (the reference raises ValueError for a wrong separator count or
    an empty segment; it reads several `.item()`s per sample).  CPU labels are split on the host with numpy."""
    dev = labels.device
    if labels.is_cuda:
        B = labels.shape[0]
        out, lens, status = K.split_labels(labels.detach(), k_speakers, sep_id, pad_token_id, ignore_id, end_token_id,
                                           allow_empty_segment)
        host = torch.cat([status.to(torch.int64), lens.max(dim=1).values if B else lens.new_zeros(k_speakers)]).tolist()
        bad_row, kind, info = host[0], host[1], host[2]
        if bad_row < B:
            if kind == 1:
                raise ValueError(f"[split_k_speakers_and_lengths_strict] Sample index {bad_row}: found {info} separators "
                                 f"(token id={sep_id}) but expected {k_speakers - 1}. labels[b].shape={tuple(labels[bad_row].shape)}")
            raise ValueError(f"[split_k_speakers_and_lengths_strict] Sample {bad_row}, speaker-slot {info} resulted in an "
                             f"empty segment while allow_empty_segment=False.")
        return ([out[i, :, :host[3 + i]].contiguous() for i in range(k_speakers)], [lens[i] for i in range(k_speakers)])
    rows = labels.detach().to("cpu", torch.int64).numpy()
    B = rows.shape[0]
    segs: List[List[np.ndarray]] = [[] for _ in range(k_speakers)]
    for b in range(B):
        row = rows[b]
        if end_token_id is not None:
            hit = np.flatnonzero(row == end_token_id)
            if hit.size:
                row = row[: hit[0]]
        seps = np.flatnonzero(row == sep_id)
        if seps.size != k_speakers - 1:
            raise ValueError(f"[split_k_speakers_and_lengths_strict] Sample index {b}: found {seps.size} separators "
                             f"(token id={sep_id}) but expected {k_speakers - 1}. labels[b].shape={tuple(row.shape)}")
        bounds = np.concatenate(([-1], seps, [row.size]))
        for i in range(k_speakers):
            seg = row[bounds[i] + 1: bounds[i + 1]]
            if ignore_id is not None:
                seg = seg[seg != ignore_id]
            if pad_token_id is not None and seg.size:
                keep = np.flatnonzero(seg != pad_token_id)
                seg = seg[: keep[-1] + 1] if keep.size else seg[:0]
            if seg.size == 0 and not allow_empty_segment:
                raise ValueError(f"[split_k_speakers_and_lengths_strict] Sample {b}, speaker-slot {i} resulted in an "
                                 f"empty segment while allow_empty_segment=False.")
            segs[i].append(seg)
    labs, lens = [], []
    for i in range(k_speakers):
        ln = np.array([s.size for s in segs[i]], dtype=np.int64)
        m = int(ln.max()) if B else 0
        mat = np.full((B, m), pad_token_id, dtype=np.int64)
        for b, s in enumerate(segs[i]):
            mat[b, : s.size] = s
        labs.append(torch.from_numpy(mat).to(dev, non_blocking=True))
        lens.append(torch.from_numpy(ln).to(dev, non_blocking=True))
    return labs, lens
