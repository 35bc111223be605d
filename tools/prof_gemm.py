"""One GEMM variant a few times (for `ncu --set full -k regex:gemm_bf16`) or an epilogue-cost ladder (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as Kn

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "ladder"
M, N, K = 32 * 499, 4096, 1024
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.1).to(torch.bfloat16)
x, w, b = rnd(M, K), rnd(N, K), torch.randn(N, device=dev)
yb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
yf = torch.empty(M, N, device=dev, dtype=torch.float32)
u = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
rb = rnd(M, N)
A, B = Kn.Operand(x, K), Kn.Operand(w, K)
variants = {
    "plain_bf16": lambda: Kn.gemm(A, B, M, N, K, Kn.Out(yb, N)),
    "plain_f32": lambda: Kn.gemm(A, B, M, N, K, Kn.Out(yf, N)),
    "bias": lambda: Kn.gemm(A, B, M, N, K, Kn.Out(yb, N), bias=b),
    "bias_gelu": lambda: Kn.gemm(A, B, M, N, K, Kn.Out(yb, N), bias=b, act=Kn.ACT_GELU),
    "bias_relu": lambda: Kn.gemm(A, B, M, N, K, Kn.Out(yb, N), bias=b, act=Kn.ACT_RELU),
    "bias_gelu_aux": lambda: Kn.gemm(A, B, M, N, K, Kn.Out(yb, N), bias=b, act=Kn.ACT_GELU, aux=u),
    "res_bf16": lambda: Kn.gemm(A, B, M, N, K, Kn.Out(yb, N), residual=Kn.Out(rb, N)),
    "gelu_bwd": lambda: Kn.gemm(A, B, M, N, K, Kn.Out(yb, N), act=Kn.ACT_GELU_BWD, residual=Kn.Out(rb, N)),
}
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


if which == "ladder":
    for n, fn in variants.items():
        ms = timeit(fn)
        print(f"{n:16s} {ms:8.3f} ms {2 * M * N * K / ms / 1e9:8.1f} TFLOP/s")
else:
    for _ in range(5):
        variants[which]()
    torch.cuda.synchronize()
