"""Pure-Python restatements of the integer/host helpers on the hot path.
Test infrastructure only (see oracle/__init__.py).

  collapse()            ref:models/modeling_speech_encoder_decoder_llama.py:902-972
  split_labels()        ref:utils/split_labels_by_sc.py:5-97
  prefix_ids()          ref:models/ctc_prompt.py:5-120 (ids/mask part; embedding lookup left to caller)
  token_segments()      ref:models/mt_ctctoken_builder.py:56-157 (segment boundaries)
  feat_lengths()        hf:640-659 / ref:models/modeling_wavlm.py:508-533
"""
from typing import List, Optional, Sequence, Tuple


def collapse(rows: Sequence[Sequence[int]], blank_id: int, pad_id: int) -> Tuple[List[List[int]], List[int]]:
    """Non-classic greedy collapse: drop pad, drop blank, drop a token equal to the LAST KEPT token
    (so A,blank,A -> A).  Both `collapse_across_blanks` branches of the reference reduce to this."""
    out, lens = [], []
    for row in rows:
        kept: List[int] = []
        for tok in row:
            tok = int(tok)
            if tok == pad_id or tok == blank_id:
                continue
            if kept and kept[-1] == tok:
                continue
            kept.append(tok)
        out.append(kept)
        lens.append(len(kept))
    return out, lens


def pad_rows(rows: List[List[int]], pad_id: int) -> List[List[int]]:
    m = max((len(r) for r in rows), default=0)
    return [list(r) + [pad_id] * (m - len(r)) for r in rows]


def split_labels(labels: Sequence[Sequence[int]], k_speakers: int, sep_id: int, pad_token_id: int,
                 ignore_id: Optional[int] = -100, end_token_id: Optional[int] = -100,
                 allow_empty_segment: bool = True):
    per_spk = [[] for _ in range(k_speakers)]
    for b, row in enumerate(labels):
        row = [int(v) for v in row]
        if end_token_id is not None and end_token_id in row:
            row = row[: row.index(end_token_id)]
        seps = [i for i, v in enumerate(row) if v == sep_id]
        if len(seps) != k_speakers - 1:
            raise ValueError(f"sample {b}: found {len(seps)} separators, expected {k_speakers - 1}")
        starts = [0] + [i + 1 for i in seps]
        ends = seps + [len(row)]
        for i, (s, e) in enumerate(zip(starts, ends)):
            seg = row[s:e]
            if ignore_id is not None:
                seg = [v for v in seg if v != ignore_id]
            if pad_token_id is not None:
                while seg and seg[-1] == pad_token_id:
                    seg.pop()
            if not seg and not allow_empty_segment:
                raise ValueError(f"sample {b}, speaker-slot {i}: empty segment")
            per_spk[i].append(seg)
    labs = [pad_rows(rows, pad_token_id) for rows in per_spk]
    lens = [[len(r) for r in rows] for rows in per_spk]
    return labs, lens


def prefix_ids(heads: Sequence[Sequence[Sequence[int]]], pad_id: int, max_per_head: Optional[int] = None):
    """heads[k][b] = padded id row.  Returns ids (B, Ltot) and mask (B, Ltot); raises like torch.cat on a
    sample with no token from any head (ref:models/ctc_prompt.py:97-104)."""
    B = len(heads[0])
    rows = []
    for b in range(B):
        cat: List[int] = []
        n_parts = 0
        for h in heads:
            v = [int(x) for x in h[b] if int(x) != pad_id]
            if max_per_head is not None:
                v = v[:max_per_head]
            if v:
                n_parts += 1
                cat += v
        if n_parts == 0:
            raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
        rows.append(cat)
    m = max(1, max(len(r) for r in rows))
    ids = [r + [pad_id] * (m - len(r)) for r in rows]
    mask = [[True] * len(r) + [False] * (m - len(r)) for r in rows]
    return ids, mask


def token_segments(path: Sequence[int], valid: Sequence[bool], blank_id: int) -> List[List[int]]:
    """Frames of every maximal run of one non-blank token; blank closes a segment; stop at first masked frame."""
    segs: List[List[int]] = []
    cur: List[int] = []
    prev = None
    for t, tok in enumerate(path):
        if not valid[t]:
            break
        tok = int(tok)
        if tok == blank_id:
            if cur:
                segs.append(cur)
                cur = []
            prev = None
            continue
        if prev is None or tok != prev:
            # NB the reference does NOT flush an open segment when the token changes without a blank:
            # it simply restarts `current_indices` (ref:models/mt_ctctoken_builder.py:120-124).
            cur = [t]
            prev = tok
        else:
            cur.append(t)
    if cur:
        segs.append(cur)
    return segs


def feat_lengths(n: int, kernels=(10, 3, 3, 3, 3, 2, 2), strides=(5, 2, 2, 2, 2, 2, 2)) -> int:
    for k, s in zip(kernels, strides):
        n = (n - k) // s + 1
    return n
