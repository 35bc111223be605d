"""graphs.GraphedTrainStep: forward + backward of the serialized-CTC path captured as one CUDA graph must reproduce the eager
step -- for new inputs copied into the static buffers, after an optimizer update of the weights (the operand caches are
bypassed inside the graph), and with training-mode dropout (fresh masks per replay, seeded by torch's CUDA generator)."""
import pytest
import torch

from _util import build_ours, load_model_golden, named_params, rel

pytestmark = pytest.mark.gpu


def _setup(cuda, p_drop=0.0):
    from oracle.model_ref import make_config
    g, params, _ = load_model_golden("tiny_large")
    n_spk, vocab, hs = int(g["n_spk"]), int(g["vocab"]), int(g["hidden_sep"])
    over = dict(hidden_dropout=p_drop, activation_dropout=p_drop, attention_dropout=p_drop) if p_drop else {}
    enc, sep, heads, loss_mod = build_ours(make_config("tiny_large", **over), n_spk, hs, vocab, params)
    wav, mask = torch.from_numpy(g["wav"]).to(cuda), torch.from_numpy(g["mask"]).to(cuda)
    fm = torch.from_numpy(g["frame_mask"]).to(cuda)
    labels = [torch.from_numpy(g[f"labels{i}"]).to(cuda) for i in range(n_spk)]
    lens = [torch.from_numpy(g[f"lab_lens{i}"]).to(cuda) for i in range(n_spk)]

    def fn(w, m, y0, y1, n0, n1):
        out = enc(w, attention_mask=m)
        seps = sep(out[1])
        fmask = enc._get_feature_vector_attention_mask_x0(out[1].shape[1], m)
        return loss_mod(talker_ctc=heads, sep_hidden_states=seps, encoder_attention_mask_ctc=fmask, label_spks=[y0, y1],
                        label_spks_lengths=[n0, n1], talker_numbers=n_spk)

    named = named_params(enc, sep, heads)
    params_l = [v for v in named.values() if v.requires_grad]

    def release():      # HybridLoss keeps the per-head losses WITH their graph (PCGrad): a stale graph pins the accumulators
        loss_mod.last_ctc_per_head = None

    return enc, fn, [wav, mask, labels[0], labels[1], lens[0], lens[1]], named, params_l, release


def _eager(fn, inputs, params_l, release):
    for p in params_l:
        p.grad = None
    loss = fn(*inputs)
    loss.backward()
    out = loss.detach().clone(), [None if p.grad is None else p.grad.clone() for p in params_l]
    del loss
    release()
    return out


def _compare(ga, gb, tol=1e-4):
    for a, b in zip(ga, gb):
        assert (a is None) == (b is None)
        if a is not None and a.numel() >= 256:
            assert rel(a, b) < tol, rel(a, b)       # identical kernels; atomics reorder a few sums


def test_graphed_step_matches_eager_and_tracks_weight_updates(cuda):
    from mtasr_b200.graphs import GraphedTrainStep
    enc, fn, inputs, named, params_l, release = _setup(cuda)
    l_e, g_e = _eager(fn, inputs, params_l, release)
    step = GraphedTrainStep(fn, inputs, params_l, release=release)
    assert step.launches_per_replay > 50
    l_g = step(*inputs).clone()
    g_g = [None if p.grad is None else p.grad.clone() for p in params_l]
    assert abs(l_g.item() - l_e.item()) < 1e-5 * abs(l_e.item())
    _compare(g_g, g_e)
    # new inputs: other waveform (copied into the static buffers)
    inputs2 = list(inputs)
    inputs2[0] = inputs[0].flip(0).contiguous()
    inputs2[1] = inputs[1].flip(0).contiguous()
    l_e2, g_e2 = _eager(fn, inputs2, params_l, release)
    l_g2 = step(*inputs2).clone()
    assert abs(l_e2.item() - l_e.item()) > 1e-6 * abs(l_e.item())
    assert abs(l_g2.item() - l_e2.item()) < 1e-5 * abs(l_e2.item())
    _compare([None if p.grad is None else p.grad.clone() for p in params_l], g_e2)
    # "optimizer step": every trainable weight moves; the replay must see the new values (caches bypassed in the graph)
    with torch.no_grad():
        for p in params_l:
            p.add_(0.01 * torch.randn_like(p))
    l_e3, g_e3 = _eager(fn, inputs, params_l, release)
    l_g3 = step(*inputs).clone()
    assert abs(l_e3.item() - l_e.item()) > 1e-5 * abs(l_e.item())
    assert abs(l_g3.item() - l_e3.item()) < 1e-5 * abs(l_e3.item())
    _compare([None if p.grad is None else p.grad.clone() for p in params_l], g_e3)
    # a caller that clears the gradients between steps (optimizer.zero_grad(): set_to_none) finds them again after the replay
    # -- NOT the tensors of an eager step that ran in between (which is what _eager leaves behind above)
    for p in params_l:
        p.grad = None
    step(*inputs)
    assert all(p.grad is g for p, g in zip(step.params, step.static_grads))
    _compare([None if p.grad is None else p.grad.clone() for p in params_l], g_e3)
    with pytest.raises(ValueError):
        step(*[t[:1] for t in inputs])


def test_graphed_step_with_dropout_draws_fresh_masks(cuda):
    from mtasr_b200.graphs import GraphedTrainStep
    enc, fn, inputs, named, params_l, release = _setup(cuda, p_drop=0.1)
    enc.train()
    torch.manual_seed(3)
    step = GraphedTrainStep(fn, inputs, params_l, release=release)
    a = step(*inputs).item()
    b = step(*inputs).item()
    assert a != b                                    # the generator advances per replay
    enc.eval()
