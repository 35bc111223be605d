"""HBM-bound row kernels at the cfg2 shapes (B*T = 15968 rows): CUDA-event timing + achieved GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as Kn, ops

dev = torch.device("cuda:0")
M, D, H, B, T = 32 * 499, 1024, 16, 32, 499
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


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


def rep(name, ms, nbytes):
    print(f"{name:28s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.0f} GB/s")


x32 = torch.randn(M, D, device=dev)
xb = x32.to(torch.bfloat16)
g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
rep("ln_fwd f32->bf16", timeit(lambda: Kn.layernorm_fwd(x32, g, b, 1e-5)), M * D * 6)
yb, _, mean, rstd = Kn.layernorm_fwd(x32, g, b, 1e-5)
dyb = torch.randn(M, D, device=dev).to(torch.bfloat16)
dres = torch.randn(M, D, device=dev)
rep("ln_bwd (dy bf16, x f32)->f32", timeit(lambda: Kn.layernorm_bwd(dyb, x32, mean, rstd, g, dres=dres)), M * D * 14)
rep("ln_bwd + bf16 twin + 3 sums", timeit(lambda: Kn.layernorm_bwd(dyb, x32, mean, rstd, g, dres=dres, want_bf16=True, want_dxsum=True)), M * D * 16)
rep("cast f32->bf16", timeit(lambda: Kn.cast_bf16(x32)), M * D * 6)
big = torch.randn(M, 4096, device=dev).to(torch.bfloat16)
rep("colsum bf16 (M,4096)", timeit(lambda: Kn.colsum(big)), M * 4096 * 2)
rep("colsum bf16 (M,1024)", timeit(lambda: Kn.colsum(xb)), M * D * 2)
w = torch.randn(8, 64, device=dev); bb = torch.randn(8, device=dev); cst = torch.rand(1, H, 1, 1, device=dev)
wab, bab, c1 = w.contiguous(), bb.contiguous(), cst.reshape(H).contiguous()   # raw (8,64) / (8,) projection: summed in-kernel
h3 = xb.view(B, T, D)
rep("gate fwd", timeit(lambda: Kn.relpos_gate_fwd(h3, wab, bab, c1, B, T, H)), M * D * 2)
dg = torch.randn(B, H, T, device=dev)
rep("gate bwd", timeit(lambda: Kn.relpos_gate_bwd(h3, wab, bab, c1, dg, B, T, H)), M * D * 6)
wav = torch.randn(32, 160000, device=dev)
w0 = torch.randn(512, 1, 10, device=dev) * 0.1
rep("conv0 + LN + GELU", timeit(lambda: Kn.conv0_fwd(wav, w0, None, torch.ones(512, device=dev), torch.zeros(512, device=dev), 1e-5, 10, 5, True)),
    32 * 31999 * 512 * 2 + 32 * 160000 * 4)
