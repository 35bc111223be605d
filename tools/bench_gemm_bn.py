"""block_n sweep on the per-layer GEMM shapes (CUDA events, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as Kn
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]
def rnd(*shape): return (torch.randn(*shape, device=dev) * 0.1).to(torch.bfloat16)
M = 32 * 499
for name, N, Kd, f32res in [("out_res_f32", 1024, 1024, True), ("ffn2_res_f32", 1024, 4096, True), ("out_dgrad", 1024, 1024, False), ("qkv_fwd", 3072, 1024, False), ("ffn1_fwd", 4096, 1024, False)]:
    x, w, b = rnd(M, Kd), rnd(N, Kd), torch.randn(N, device=dev)
    for bn in (0, 128, 256):
        if f32res:
            resid = torch.randn(M, N, device=dev); y = torch.empty(M, N, device=dev)
            fn = lambda: Kn.gemm(Kn.Operand(x, Kd), Kn.Operand(w, Kd), M, N, Kd, Kn.Out(y, N), bias=b, residual=Kn.Out(resid, N), block_n=bn)
        else:
            y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            fn = lambda: Kn.gemm(Kn.Operand(x, Kd), Kn.Operand(w, Kd), M, N, Kd, Kn.Out(y, N), bias=b, block_n=bn)
        ms = timeit(fn)
        print(f"{name:14s} bn={bn:3d}  {ms * 1e3:7.1f} us  {2 * M * N * Kd / ms / 1e9:7.0f} TFLOP/s", flush=True)
