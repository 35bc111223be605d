"""The oracle against the golden vectors produced by the reference itself (oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import torch

from _util import GOLDEN, load_model_golden, sub_state


def test_ctc_numpy_oracle_vs_reference_golden():
    from oracle import ctc_ref
    g = np.load(os.path.join(GOLDEN, "ctc_small.npz"))
    nll, grad = ctc_ref.ctc_loss_and_grad(g["logits"], g["hlens"], g["ys"], g["ylens"], int(g["blank"]))
    assert np.abs(nll - g["nll"]).max() < 1e-10
    assert np.abs(grad * g["upstream"][:, None, None] - g["grad"]).max() < 1e-10
    assert nll[5] == 0.0 and np.abs(grad[5]).max() == 0.0                        # infeasible row (zero_infinity)
    assert nll[4] > 0.0                                                           # empty target: all-blank path


def test_ctc_brute_force_small():
    from oracle import ctc_ref
    lp = ctc_ref.log_softmax(np.random.RandomState(1).randn(6, 5))
    for y in ([0], [1, 1], [0, 2], [3, 0, 3], []):
        a = ctc_ref.ctc_alpha_beta(lp, y, 4)[0]
        b = ctc_ref.ctc_brute_force(lp, y, 4)
        assert abs(a - b) < 1e-10, (y, a, b)


def test_host_oracle_vs_reference_golden():
    from oracle import host_ref
    g = np.load(os.path.join(GOLDEN, "host_small.npz"))
    rows, lens = host_ref.collapse(g["argmax"].tolist(), int(g["blank"]), int(g["pad"]))
    assert lens == g["collapsed_lens"].tolist()
    assert host_ref.pad_rows(rows, int(g["pad"])) == g["collapsed"].tolist()
    labs, ln = host_ref.split_labels(g["split_labels"].tolist(), 2, int(g["split_sc"]), int(g["split_pad"]), -100,
                                     int(g["split_pad"]), False)
    assert labs == [g["split0"].tolist(), g["split1"].tolist()] and ln == [g["split_len0"].tolist(), g["split_len1"].tolist()]
    ids, mask = host_ref.prefix_ids([g["prefix_h0"].tolist(), g["prefix_h1"].tolist()], int(g["split_pad"]), None)
    assert ids == g["prefix_ids"].tolist() and mask == g["prefix_mask"].tolist()
    # token-builder segmentation: segment means of the reference (tb_mem) from the oracle's segment boundaries
    x = torch.from_numpy(g["tb_x"])
    path = (x @ torch.from_numpy(g["tb_w"]).t() + torch.from_numpy(g["tb_b"])).argmax(-1)
    for b in range(x.shape[0]):
        segs = host_ref.token_segments(path[b].tolist(), g["tb_mask"][b].tolist(), 8)
        for j, s in enumerate(segs):
            assert np.allclose(g["tb_mem"][b, j], x[b, s].mean(0).numpy(), atol=1e-6)


def test_model_oracle_reproduces_reference_golden():
    """RefWavLMModel / RefSeparator / RefCTC / ref_hybrid_ctc with the fixture's weights give the reference's outputs."""
    from oracle.model_ref import RefCTC, RefSeparator, RefWavLMModel, make_config, ref_hybrid_ctc
    for kind in ("tiny_large", "tiny_base"):
        g, params, _ = load_model_golden(kind)
        n_spk, vocab, hs = int(g["n_spk"]), int(g["vocab"]), int(g["hidden_sep"])
        cfg = make_config(kind)
        enc = RefWavLMModel(cfg).eval(); enc.load_state_dict(sub_state(params, "encoder."))
        sep = RefSeparator(cfg.hidden_size, hs, n_spk).eval(); sep.load_state_dict(sub_state(params, "separator."))
        heads = [RefCTC(vocab, cfg.hidden_size) for _ in range(n_spk)]
        for i, h in enumerate(heads):
            h.load_state_dict(sub_state(params, f"serialized_ctc.{i}."))
        wav, mask = torch.from_numpy(g["wav"]), torch.from_numpy(g["mask"])
        with torch.no_grad():
            last, e, down, feats = enc(wav, mask)
            seps = sep(e)
            fm = enc.frame_mask_x0(e.shape[1], mask)
            loss, per = ref_hybrid_ctc(heads, seps, fm, [torch.from_numpy(g[f"labels{i}"]) for i in range(n_spk)],
                                       [torch.from_numpy(g[f"lab_lens{i}"]) for i in range(n_spk)])
        assert torch.equal(fm, torch.from_numpy(g["frame_mask"]))
        for a, b in ((last, "last"), (e, "enc"), (down, "down"), (feats, "feats"), (seps[0], "sep0")):
            assert (a - torch.from_numpy(g[b])).abs().max().item() < 1e-4, (kind, b)
        assert abs(loss.item() - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
        for i, h in enumerate(heads):
            assert torch.equal(h.argmax(seps[i]), torch.from_numpy(g[f"argmax{i}"]))
