"""SURVEY rows f3 / f4 on the device: the label splitter kernel against the reference-generated golden and the python oracle
(incl. its ValueError cases), and `pcgrad_backward` against a literal restatement of the reference's PCGrad training step
(ref:src/trainer_seq2seq.py:1071-1141: K + 1 backward passes, host-side `if dot < 0`)."""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, build_ours, load_model_golden, named_params, rel

pytestmark = pytest.mark.gpu


def test_split_labels_on_device_matches_reference_golden_and_oracle(cuda):
    from oracle import host_ref
    from mtasr_b200 import greedy
    from mtasr_b200 import kernels as K
    g = np.load(os.path.join(GOLDEN, "host_small.npz"))
    sc, pad = int(g["split_sc"]), int(g["split_pad"])
    l0 = K.launch_count()
    labs, lens = greedy.split_k_speakers_and_lengths(torch.from_numpy(g["split_labels"]).to(cuda), 2, sc, pad, ignore_id=-100,
                                                     end_token_id=pad, allow_empty_segment=False)
    assert K.launch_count() == l0 + 1                               # the device kernel ran (no host loop)
    assert all(t.is_cuda for t in labs + lens)
    assert [l.cpu().tolist() for l in labs] == [g["split0"].tolist(), g["split1"].tolist()]
    assert [l.cpu().tolist() for l in lens] == [g["split_len0"].tolist(), g["split_len1"].tolist()]
    assert labs[0].dtype == torch.int64 and lens[0].dtype == torch.int64
    rs = np.random.RandomState(1)
    for Kspk in (2, 3):
        for allow_empty in (True, False):
            rows = []
            for _ in range(33):
                segs = [rs.randint(1, 8, size=rs.randint(0 if allow_empty else 1, 6)).tolist() for _ in range(Kspk)]
                row = []
                for i, s in enumerate(segs):
                    row += s + ([9] if i < Kspk - 1 else [])
                rows.append(row + [0] * rs.randint(0, 3))
            L = max(len(r) for r in rows) + 2
            mat = torch.tensor([r + [-100] * (L - len(r)) for r in rows])
            olabs, olens = host_ref.split_labels(mat.tolist(), Kspk, 9, 0, -100, 0, allow_empty)
            labs, lens = greedy.split_k_speakers_and_lengths(mat.to(cuda), Kspk, 9, 0, ignore_id=-100, end_token_id=0,
                                                             allow_empty_segment=allow_empty)
            assert [l.cpu().tolist() for l in labs] == olabs and [l.cpu().tolist() for l in lens] == olens
            # non-contiguous view of a wider matrix (row stride != L)
            wide = torch.cat([mat, mat], 1).to(cuda)
            labs2, lens2 = greedy.split_k_speakers_and_lengths(wide[:, :L], Kspk, 9, 0, ignore_id=-100, end_token_id=0,
                                                               allow_empty_segment=allow_empty)
            assert [l.cpu().tolist() for l in labs2] == olabs
    # pad inside a segment is kept, trailing pads trimmed, -100 dropped anywhere; no end token / no pad given
    mat = torch.tensor([[5, 0, 6, 0, 0, 9, -100, 7, -100, 0, 8]]).to(cuda)
    labs, lens = greedy.split_k_speakers_and_lengths(mat, 2, 9, 0, ignore_id=-100, end_token_id=None)
    assert labs[0].cpu().tolist() == [[5, 0, 6]] and labs[1].cpu().tolist() == [[7, 0, 8]]
    with pytest.raises(ValueError, match="found 0 separators"):
        greedy.split_k_speakers_and_lengths(torch.tensor([[1, 2, 3], [1, 9, 3]]).to(cuda), 2, 9, 0)
    with pytest.raises(ValueError, match=r"Sample 1, speaker-slot 0"):
        greedy.split_k_speakers_and_lengths(torch.tensor([[1, 9, 2], [9, 1, 2]]).to(cuda), 2, 9, 0, allow_empty_segment=False)


def _reference_pcgrad_step(loss, ctc_per_head, shared, everything):
    """ref:src/trainer_seq2seq.py:1071-1141, restated: per-head grads of the shared parameters, sequential projection with a
    host-side sign test, full backward, shared grads overwritten."""
    grads = []
    for Li in ctc_per_head:
        gi = torch.autograd.grad(Li.mean(), shared, retain_graph=True, allow_unused=True)
        grads.append([g if g is not None else torch.zeros_like(p) for g, p in zip(gi, shared)])
    Kh = len(grads)
    n_conflicts = 0
    for i in range(Kh):
        for j in range(Kh):
            if i == j:
                continue
            dot = sum((a * b).sum() for a, b in zip(grads[i], grads[j]))
            if dot < 0:
                n_conflicts += 1
                norm2 = sum((b * b).sum() for b in grads[j]) + 1e-12
                alpha = dot / norm2
                grads[i] = [a - alpha * b for a, b in zip(grads[i], grads[j])]
    proj = [sum(grads[i][k] for i in range(Kh)) for k in range(len(shared))]
    for p in everything:
        p.grad = None
    loss.backward()
    for p, g in zip(shared, proj):
        p.grad = g.detach()
    return n_conflicts


@pytest.mark.parametrize("conflict", [False, True])
def test_pcgrad_backward_matches_reference_procedure(cuda, conflict):
    from mtasr_b200.pcgrad import pcgrad_backward
    g, params, _ = load_model_golden("tiny_large")
    n_spk, vocab, hs = int(g["n_spk"]), int(g["vocab"]), int(g["hidden_sep"])
    from oracle.model_ref import make_config
    enc, sep, heads, loss_mod = build_ours(make_config("tiny_large"), n_spk, hs, vocab, params)
    for p in enc.adapter.parameters():
        p.requires_grad_(False)
    wav, mask = torch.from_numpy(g["wav"]).to(cuda), torch.from_numpy(g["mask"]).to(cuda)
    fm = torch.from_numpy(g["frame_mask"]).to(cuda)
    labels = [torch.from_numpy(g[f"labels{i}"]).to(cuda) for i in range(n_spk)]
    lens = [torch.from_numpy(g[f"lab_lens{i}"]).to(cuda) for i in range(n_spk)]
    shared = [p for p in list(enc.parameters()) + list(sep.parameters()) if p.requires_grad]
    other = [p for p in heads.parameters() if p.requires_grad]

    def forward():
        out = enc(wav, attention_mask=mask)
        seps = sep(out[1])
        loss = loss_mod(talker_ctc=heads, sep_hidden_states=seps, encoder_attention_mask_ctc=fm, label_spks=labels,
                        label_spks_lengths=lens, talker_numbers=n_spk)
        per_head = list(loss_mod.last_ctc_per_head)
        if conflict:      # make head 1 pull against head 0 on the shared parameters
            per_head = [per_head[0], 0.2 * per_head[1] - per_head[0]]
        return loss, per_head

    loss, per_head = forward()
    n_conf = _reference_pcgrad_step(loss, per_head, shared, shared + other)
    want = [p.grad.clone() for p in shared + other]
    assert (n_conf > 0) == conflict
    for p in shared + other:
        p.grad = None
    loss, per_head = forward()
    pcgrad_backward(loss, per_head, shared, other)
    got = [p.grad for p in shared + other]
    num = sum((a.float() - b.float()).pow(2).sum().item() for a, b in zip(got, want))
    den = sum(b.float().pow(2).sum().item() for b in want)
    assert (num / den) ** 0.5 < 1e-4, (num / den) ** 0.5          # identical kernels; fp32 atomics reorder a few sums
    for a, b in zip(got, want):
        if b.numel() >= 256 and b.norm().item() > 1e-6:
            assert rel(a, b) < 5e-3
