"""CTC alpha / beta+grad kernels at the cfg2 lattice shape (B=32, T=499, L~U(20,60)): CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as Kn

dev = torch.device("cuda:0")
B, T = int(os.environ.get("B", 32)), int(os.environ.get("T", 499))
g = torch.Generator().manual_seed(0)
ylens = torch.randint(20, 61, (B,), generator=g)
Lmax = int(ylens.max())
Lp = (Lmax + 1 + 63) // 64 * 64
ys = torch.randint(0, 1000, (B, Lmax), generator=g).to(dev)
ylens = ylens.to(dev)
hlens = torch.full((B,), T, dtype=torch.int64, device=dev)
glog = torch.randn(B, T, Lp, device=dev)
lse = torch.logsumexp(glog, -1) + 2.0
gout = torch.ones(B, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


nll, nll_raw, alpha, coff = Kn.ctc_alpha_fwd(glog, lse, ys, hlens, ylens, Lmax)
fa = timeit(lambda: Kn.ctc_alpha_fwd(glog, lse, ys, hlens, ylens, Lmax))
fb = timeit(lambda: Kn.ctc_beta_bwd(glog, lse, ys, hlens, ylens, Lmax, alpha, coff, nll_raw, gout))
print(f"B={B} T={T} Lmax={Lmax} Lp={Lp}: alpha {fa * 1e3:.1f} us ({fa * 1e3 / T:.3f} us/frame)  beta+grad {fb * 1e3:.1f} us ({fb * 1e3 / T:.3f} us/frame)  nll[0]={nll[0].item():.3f}")
