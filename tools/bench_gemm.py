"""Micro-benchmark of the tcgen05 GEMM on the hot-path shapes (CUDA events, L2 flushed between runs)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as Kn

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def rnd(*shape):
    return (torch.randn(*shape, device=dev) * 0.1).to(torch.bfloat16)


res = []
M = 32 * 499
for name, N, K in [("qkv", 3072, 1024), ("out", 1024, 1024), ("ffn1", 4096, 1024), ("ffn2", 1024, 4096)]:
    x, w = rnd(M, K), rnd(N, K)
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: Kn.gemm(Kn.Operand(x, K), Kn.Operand(w, K), M, N, K, Kn.Out(y, N)))
    res.append((name + "_fwd", ms, 2 * M * N * K / ms / 1e9))
    dy = rnd(M, N)
    dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: Kn.gemm(Kn.Operand(dy, N), Kn.Operand(w, K, major=1), M, K, N, Kn.Out(dx, K)))
    res.append((name + "_dgrad", ms, 2 * M * N * K / ms / 1e9))
    dw = torch.empty(N, K, device=dev, dtype=torch.float32)
    ms = timeit(lambda: Kn.gemm(Kn.Operand(dy, N, major=1), Kn.Operand(x, K, major=1), N, K, M, Kn.Out(dw, K)))
    res.append((name + "_wgrad", ms, 2 * M * N * K / ms / 1e9))
# epilogue-heavy variants as the model issues them
x, w1, b1 = rnd(M, 1024), rnd(4096, 1024), torch.randn(4096, device=dev)
a = torch.empty(M, 4096, device=dev, dtype=torch.bfloat16); u = torch.empty_like(a)
ms = timeit(lambda: Kn.gemm(Kn.Operand(x, 1024), Kn.Operand(w1, 1024), M, 4096, 1024, Kn.Out(a, 4096), bias=b1, act=Kn.ACT_GELU, aux=u))
res.append(("ffn1_gelu_aux", ms, 2 * M * 4096 * 1024 / ms / 1e9))
w2, b2 = rnd(1024, 4096), torch.randn(1024, device=dev)
resid = torch.randn(M, 1024, device=dev); y32 = torch.empty(M, 1024, device=dev)
ms = timeit(lambda: Kn.gemm(Kn.Operand(a, 4096), Kn.Operand(w2, 4096), M, 1024, 4096, Kn.Out(y32, 1024), bias=b2, residual=Kn.Out(resid, 1024)))
res.append(("ffn2_res_f32", ms, 2 * M * 4096 * 1024 / ms / 1e9))
wo = rnd(1024, 1024)
ms = timeit(lambda: Kn.gemm(Kn.Operand(x, 1024), Kn.Operand(wo, 1024), M, 1024, 1024, Kn.Out(y32, 1024), bias=b2, residual=Kn.Out(resid, 1024)))
res.append(("out_res_f32", ms, 2 * M * 1024 * 1024 / ms / 1e9))
dyb = rnd(M, 1024); du = torch.empty(M, 4096, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: Kn.gemm(Kn.Operand(dyb, 1024), Kn.Operand(w2, 4096, major=1), M, 4096, 1024, Kn.Out(du, 4096), act=Kn.ACT_GELU_BWD, residual=Kn.Out(u, 4096)))
res.append(("ffn2_dgrad_gelubwd", ms, 2 * M * 4096 * 1024 / ms / 1e9))
Bq, Hh, Tt, dd = 32, 16, 499, 64
Tp = 504
qkv = rnd(Bq * Tt, 3 * 1024)
S = torch.empty(Bq, Hh, Tt, Tp, device=dev)
ms = timeit(lambda: Kn.gemm(Kn.Operand(qkv, 3072, sb0=dd, sb1=Tt * 3072, rows=Tt), Kn.Operand(qkv, 3072, sb0=dd, sb1=Tt * 3072, offset=1024, rows=Tt),
                            Tt, Tt, dd, Kn.Out(S, Tp, sb0=Tt * Tp, sb1=Hh * Tt * Tp), batch=(Hh, Bq)))
res.append(("attn_qk_f32", ms, 2 * Bq * Hh * Tt * Tt * dd / ms / 1e9))
Pm = torch.zeros(Bq, Hh, Tt, Tp, device=dev, dtype=torch.bfloat16)
O = torch.empty(Bq * Tt, 1024, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: Kn.gemm(Kn.Operand(Pm, Tp, sb0=Tt * Tp, sb1=Hh * Tt * Tp), Kn.Operand(qkv, 3072, major=1, sb0=dd, sb1=Tt * 3072, offset=2048, rows=Tt),
                            Tt, dd, Tt, Kn.Out(O, 1024, sb0=dd, sb1=Tt * 1024), batch=(Hh, Bq)))
res.append(("attn_pv", ms, 2 * Bq * Hh * Tt * Tt * dd / ms / 1e9))
V, K = 128259, 1024
h, w = rnd(M, K), rnd(V, K)
nt = Kn.gemm_n_tiles(V)
part = torch.empty(M, nt, 4, device=dev)
ms = timeit(lambda: Kn.gemm(Kn.Operand(h, K), Kn.Operand(w, K), M, V, K, None, mode=1, lse_part=part), iters=5)
res.append(("vocab_lse", ms, 2 * M * V * K / ms / 1e9))
lse, _ = Kn.lse_finalize(part, M, nt)
Vp = (V + 7) // 8 * 8
P = torch.empty(M, Vp, device=dev, dtype=torch.bfloat16)
ms = timeit(lambda: Kn.gemm(Kn.Operand(h, K), Kn.Operand(w, K), M, V, K, Kn.Out(P, Vp), mode=2, row_vec=lse), iters=5)
res.append(("vocab_exp", ms, 2 * M * V * K / ms / 1e9))
dh = torch.empty(M, K, device=dev, dtype=torch.float32)
ms = timeit(lambda: Kn.gemm(Kn.Operand(P, Vp), Kn.Operand(w, K, major=1), M, K, V, Kn.Out(dh, K)), iters=5)
res.append(("vocab_dgrad", ms, 2 * M * V * K / ms / 1e9))
dw = torch.empty(V, K, device=dev, dtype=torch.float32)
ms = timeit(lambda: Kn.gemm(Kn.Operand(P, Vp, major=1), Kn.Operand(h, K, major=1), V, K, M, Kn.Out(dw, K)), iters=5)
res.append(("vocab_wgrad", ms, 2 * M * V * K / ms / 1e9))
x, w = rnd(M, 1024), rnd(4096, 1024)
ms = timeit(lambda: torch.matmul(x, w.t()))
res.append(("torch_ffn1_fwd", ms, 2 * M * 4096 * 1024 / ms / 1e9))
for r in res:
    print(f"{r[0]:16s} {r[1]:8.3f} ms  {r[2]:8.1f} TFLOP/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_gemm.json", "w"))
