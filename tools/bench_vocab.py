"""Vocabulary-sized GEMMs of the CTC head (cfg2: M = 32 * 499 rows, V = 128259, D = 1024) at two row pitches of the
(rows, V) logits / softmax matrices: the minimal one (multiple of 8 elements, the tensor-map stride rule) and a multiple of
64 elements (rows start on 128-byte lines).  CUDA events, L2 flushed between launches.

    python tools/bench_vocab.py            -> table on stdout + gpurun_out/bench_vocab.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as Kn

dev = torch.device("cuda:0")
M, V, D = 32 * 499, 128259, 1024
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.05).to(torch.bfloat16)
h, w = rnd(M, D), rnd(V, D)
bias = torch.zeros(V, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
nt = Kn.gemm_n_tiles(V)
part = torch.empty(M, nt, 4, device=dev)
for align in (8, 64):
    Vp = (V + align - 1) // align * align
    lg = torch.empty(M, Vp, device=dev, dtype=torch.float16)
    fl = 2 * M * V * D
    ms = timeit(lambda: Kn.gemm(Kn.Operand(h, D), Kn.Operand(w, D), M, V, D, Kn.Out(lg, Vp), bias=bias, mode=1, lse_part=part))
    res[f"fwd_lse_logits16 pitch%{align}"] = (ms, fl / ms / 1e9)
    lse, _ = Kn.lse_finalize(part, M, nt)
    rs = torch.full((M,), 1.0 / M, device=dev)
    ms = timeit(lambda: Kn.softmax_from_logits(lg, lse.view(-1), rs, V, want_colsum=True))
    res[f"softmax_from_logits pitch%{align}"] = (ms, 4.0 * M * V / ms / 1e6)            # GB/s in the second column
    P, _ = Kn.softmax_from_logits(lg, lse.view(-1), rs, V, want_colsum=True)
    dh = torch.empty(M, D, device=dev, dtype=torch.float32)
    ms = timeit(lambda: Kn.gemm(Kn.Operand(P, Vp), Kn.Operand(w, D, major=1), M, D, V, Kn.Out(dh, D)))
    res[f"dgrad P.W pitch%{align}"] = (ms, fl / ms / 1e9)
    dw = torch.empty(V, D, device=dev, dtype=torch.float32)
    ms = timeit(lambda: Kn.gemm(Kn.Operand(P, Vp, major=1), Kn.Operand(h, D, major=1), V, D, M, Kn.Out(dw, D)))
    res[f"wgrad P^T.H pitch%{align}"] = (ms, fl / ms / 1e9)
    del lg, P
for k, (ms, r) in res.items():
    print(f"{k:36s} {ms:8.3f} ms  {r:9.1f} {'GB/s' if 'softmax' in k else 'TFLOP/s'}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_vocab.json", "w"))
