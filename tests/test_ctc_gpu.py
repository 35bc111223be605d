"""CTC alpha/beta kernels, collapse and LSE finalize vs the fp64 oracle and the reference-generated golden vectors.
Tolerances (BASELINE north_star): CTC loss and gradients <= 1e-5 relative in fp32; greedy tokens bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ctc_dense(Kn, logits, hlens, ys, ylens, blank, upstream):
    B, T, V = logits.shape
    Lmax = int(ylens.max())
    Lp = (Lmax + 1 + 7) // 8 * 8
    glog = Kn.ctc_gather_cols(logits, ys, ylens, Lp, blank)
    lse = torch.logsumexp(logits, -1)
    nll, nll_raw, alpha, coff = Kn.ctc_alpha_fwd(glog, lse, ys, hlens, ylens, Lmax)
    dG, rowscale = Kn.ctc_beta_bwd(glog, lse, ys, hlens, ylens, Lmax, alpha, coff, nll_raw, upstream)
    dlogits = torch.exp(logits - lse[..., None]) * rowscale[..., None]
    Kn.ctc_scatter_cols(dG, ys, ylens, dlogits, blank)
    return nll, dlogits


def test_ctc_golden(cuda, golden_dir):
    from mtasr_b200 import kernels as Kn
    g = np.load(os.path.join(golden_dir, "ctc_small.npz"))
    t = lambda k, dt=None: torch.from_numpy(g[k]).to(cuda) if dt is None else torch.from_numpy(g[k]).to(cuda).to(dt)
    nll, dl = _ctc_dense(Kn, t("logits"), t("hlens"), t("ys"), t("ylens"), int(g["blank"]), t("upstream"))
    ref_nll, ref_grad = t("nll"), t("grad")
    assert torch.allclose(nll.double(), ref_nll.double(), rtol=1e-5, atol=1e-5), (nll, ref_nll)
    rel = ((dl - ref_grad).norm() / ref_grad.norm()).item()
    assert rel < 1e-5, rel
    assert nll[5].item() == 0.0 and dl[5].abs().max().item() == 0.0      # infeasible row: zero_infinity


@pytest.mark.parametrize("B,T,V,Lmax,seed", [(4, 50, 30, 12, 0), (3, 200, 500, 60, 1), (2, 300, 64, 120, 2), (2, 20, 11, 0, 3),
                                              (2, 560, 300, 255, 4),      # maximum label length (16 states per lane)
                                              (3, 1, 9, 1, 5), (3, 3, 9, 2, 6),   # fewer frames than one ring chunk
                                              (2, 67, 40, 30, 7)])        # frame count not a multiple of the chunk
def test_ctc_vs_fp64_oracle(cuda, B, T, V, Lmax, seed):
    from mtasr_b200 import kernels as Kn
    from oracle import ctc_ref
    rs = np.random.RandomState(seed)
    logits = (rs.randn(B, T, V) * 2).astype(np.float32)
    hlens = rs.randint(max(1, T // 2), T + 1, size=B); hlens[0] = T
    ylens = rs.randint(0, Lmax + 1, size=B) if Lmax else np.zeros(B, dtype=np.int64)
    if Lmax:
        ylens[0] = Lmax
    ys = rs.randint(0, V - 1, size=(B, max(Lmax, 1)))
    if Lmax >= 3:
        ys[:, 2] = ys[:, 1]
    up = rs.rand(B).astype(np.float32) + 0.5
    n64, g64 = ctc_ref.ctc_loss_and_grad(logits, hlens, ys, ylens, V - 1)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    nll, dl = _ctc_dense(Kn, dev(logits), dev(hlens.astype(np.int64)), dev(ys.astype(np.int64)), dev(ylens.astype(np.int64)), V - 1, dev(up))
    assert np.allclose(nll.cpu().numpy(), n64, rtol=1e-5, atol=1e-5)
    ref = g64 * up[:, None, None]
    rel = np.linalg.norm(dl.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert rel < 1e-5, rel


def test_collapse_golden_and_random(cuda, golden_dir):
    from mtasr_b200 import kernels as Kn
    from oracle import host_ref
    g = np.load(os.path.join(golden_dir, "host_small.npz"))
    am = torch.from_numpy(g["argmax"]).to(cuda)
    out, lens = Kn.ctc_collapse(am, int(g["blank"]), int(g["pad"]))
    assert lens.cpu().tolist() == g["collapsed_lens"].tolist()
    Lm = int(lens.max())
    assert out[:, :Lm].cpu().numpy().tolist() == g["collapsed"].tolist()
    rs = np.random.RandomState(0)
    ids = rs.randint(0, 6, size=(37, 499)).astype(np.int64)       # tiny vocab -> many repeats / blanks / pads
    out, lens = Kn.ctc_collapse(torch.from_numpy(ids).to(cuda), 5, 4)
    rows, l2 = host_ref.collapse(ids.tolist(), 5, 4)
    assert lens.cpu().tolist() == l2
    o = out.cpu().numpy()
    for b, r in enumerate(rows):
        assert o[b, : len(r)].tolist() == r and (o[b, len(r):] == 4).all()
