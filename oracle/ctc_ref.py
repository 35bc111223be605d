"""float64 numpy restatement of the CTC loss used by the reference.

Follows torch.nn.CTCLoss(reduction='none', zero_infinity=True, blank=V-1) as
called from ref:models/ctc.py:44-46,51-65 (ATen LossCTC semantics):
  ext[s] = blank for even s, y[s//2] for odd s, S' = 2L+1
  alpha_0[0] = lp[0, blank], alpha_0[1] = lp[0, y_0]
  alpha_t[s] = lp[t, ext[s]] + LSE(alpha_{t-1}[s], alpha_{t-1}[s-1],
                                   alpha_{t-1}[s-2] if ext[s] != blank and ext[s] != ext[s-2])
  nll = -LSE(alpha_{T-1}[S'-1], alpha_{T-1}[S'-2])
  d nll / d logits[t, v] = softmax[t, v] - exp(LSE_{s: ext[s]=v}(alpha_t[s]+beta_t[s]) + nll - lp[t, v])
  rows t >= T_b get zero gradient; nll = +inf -> loss 0, grad 0 (zero_infinity).
Test infrastructure only (see oracle/__init__.py).
"""
import itertools
import numpy as np

NEG = -np.inf


def _lse(*xs):
    m = max(xs)
    if m == NEG:
        return NEG
    return m + np.log(sum(np.exp(x - m) for x in xs))


def log_softmax(logits):
    logits = np.asarray(logits, dtype=np.float64)
    m = logits.max(axis=-1, keepdims=True)
    z = logits - m
    return z - np.log(np.exp(z).sum(axis=-1, keepdims=True))


def ctc_alpha_beta(lp, y, blank):
    """lp: (T, V) log-probs of ONE utterance (already cut to its length); y: (L,) labels.
    returns nll, alpha (T,S'), beta (T,S')."""
    T, V = lp.shape
    L = len(y)
    S = 2 * L + 1
    ext = [blank if s % 2 == 0 else int(y[s // 2]) for s in range(S)]
    alpha = np.full((T, S), NEG)
    beta = np.full((T, S), NEG)
    if T == 0:
        return (0.0 if L == 0 else np.inf), alpha, beta
    alpha[0, 0] = lp[0, blank]
    if S > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        for s in range(S):
            terms = [alpha[t - 1, s]]
            if s >= 1:
                terms.append(alpha[t - 1, s - 1])
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                terms.append(alpha[t - 1, s - 2])
            alpha[t, s] = lp[t, ext[s]] + _lse(*terms)
    ends = [alpha[T - 1, S - 1]] + ([alpha[T - 1, S - 2]] if S > 1 else [])
    nll = -_lse(*ends)
    beta[T - 1, S - 1] = lp[T - 1, ext[S - 1]]
    if S > 1:
        beta[T - 1, S - 2] = lp[T - 1, ext[S - 2]]
    for t in range(T - 2, -1, -1):
        for s in range(S):
            terms = [beta[t + 1, s]]
            if s + 1 < S:
                terms.append(beta[t + 1, s + 1])
            if s + 2 < S and ext[s] != blank and ext[s] != ext[s + 2]:
                terms.append(beta[t + 1, s + 2])
            beta[t, s] = lp[t, ext[s]] + _lse(*terms)
    return nll, alpha, beta


def ctc_loss_and_grad(logits, hlens, ys, ylens, blank, zero_infinity=True):
    """logits: (B, T, V) float; returns nll (B,), dlogits (B, T, V) = d nll_b / d logits[b]
    (i.e. per-utterance gradient with unit upstream), both float64."""
    logits = np.asarray(logits, dtype=np.float64)
    B, T, V = logits.shape
    nll = np.zeros(B)
    grad = np.zeros_like(logits)
    for b in range(B):
        Tb, Lb = int(hlens[b]), int(ylens[b])
        lp = log_softmax(logits[b, :Tb])
        y = [int(v) for v in ys[b][:Lb]]
        n, alpha, beta = ctc_alpha_beta(lp, y, blank)
        if not np.isfinite(n):
            nll[b] = 0.0 if zero_infinity else n
            continue
        nll[b] = n
        S = 2 * Lb + 1
        ext = [blank if s % 2 == 0 else y[s // 2] for s in range(S)]
        for t in range(Tb):
            occ = {}
            for s in range(S):
                ab = alpha[t, s] + beta[t, s]
                occ[ext[s]] = _lse(occ.get(ext[s], NEG), ab)
            g = np.exp(lp[t])
            for v, lab in occ.items():
                if lab > NEG:
                    g[v] -= np.exp(lab + n - lp[t, v])
            grad[b, t] = g
    return nll, grad


def ctc_brute_force(lp, y, blank):
    """Enumerate every alignment (tiny T only): -log sum_paths prod p."""
    T, V = lp.shape
    tot = NEG
    for path in itertools.product(range(V), repeat=T):
        col = []
        prev = None
        for p in path:
            if p != prev and p != blank:
                col.append(p)
            prev = p
        if col == list(y):
            tot = _lse(tot, sum(lp[t, p] for t, p in enumerate(path)))
    return -tot
