"""Fused attention fwd/bwd at the cfg2 shape: CUDA-event timing, or a few plain launches for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as Kn

dev = torch.device("cuda:0")
B, H, T = int(os.environ.get("B", 32)), 16, int(os.environ.get("T", 499))
g = torch.Generator(device=dev).manual_seed(0)
qkv = (torch.randn(B * T, 3 * H * 64, device=dev, generator=g) * 0.8).to(torch.bfloat16)
gate = torch.rand(B, H, T, device=dev, generator=g) * 2
table = torch.randn(H, 2 * T - 1, device=dev, generator=g)
klen = torch.full((B,), T, device=dev, dtype=torch.int32)
out, lse = Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125)
dout = (torch.randn(B * T, H * 64, device=dev, generator=g) * 0.5).to(torch.bfloat16)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


if len(sys.argv) > 1 and sys.argv[1] == "once":
    for _ in range(3):
        Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125)
        Kn.attn_bwd(qkv, out, dout, lse, gate, table, klen, B, H, T, 0.125)
    torch.cuda.synchronize()
else:
    f = timeit(lambda: Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125))
    b = timeit(lambda: Kn.attn_bwd(qkv, out, dout, lse, gate, table, klen, B, H, T, 0.125))
    fl = 4.0 * B * H * T * T * 64
    print(f"B={B} H={H} T={T}: fwd {f * 1e3:.1f} us ({fl / f / 1e9:.1f} TFLOP/s algorithmic)  bwd {b * 1e3:.1f} us ({2.5 * fl / b / 1e9:.1f} TFLOP/s)")
