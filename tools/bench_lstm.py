"""LSTM recurrence micro-benchmark (CUDA events): persistent forward / backward kernels at the cfg2 separator shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as Kn

dev = torch.device("cuda:0")
B, T, Hs = int(os.environ.get("B", 32)), int(os.environ.get("T", 499)), 896
torch.manual_seed(0)
xg = torch.randn(B, T, 4 * Hs, device=dev) * 0.5
W = (torch.randn(4 * Hs, 2 * Hs, device=dev) * 0.03).to(torch.bfloat16)
whh = W[:, Hs:]


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


h, hf, c, gates = Kn.lstm_fwd(xg, whh, 2 * Hs, want_h_f32=True)
dh = torch.randn(B, T, Hs, device=dev)
ms_f = timeit(lambda: Kn.lstm_fwd(xg, whh, 2 * Hs, want_h_f32=True))
ms_b = timeit(lambda: Kn.lstm_bwd(dh, gates, c, whh, 2 * Hs))
print(f"B={B} T={T} Hs={Hs}: fwd {ms_f:.3f} ms ({ms_f / T * 1e3:.2f} us/step)  bwd {ms_b:.3f} ms ({ms_b / T * 1e3:.2f} us/step)")
