import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_attention_gpu import _inputs
from mtasr_b200 import kernels as Kn
cuda = torch.device("cuda:0")
def poison():
    junk = torch.full((64 * 1024 * 1024,), float("nan"), device=cuda); del junk
for (B, H, T) in [(2, 2, 64), (2, 3, 197)]:
    qkv, gate, table, klen = _inputs(B, H, T, cuda, seed=1)
    D = H * 64
    g = torch.Generator(device=cuda).manual_seed(5)
    dout = (torch.randn(B * T, D, device=cuda, generator=g) * 0.5).to(torch.bfloat16)
    for mode in ["clean", "poison_all", "poison_fwd_only", "poison_bwd_only"]:
        if mode in ("poison_all", "poison_fwd_only"): poison()
        out, lse = Kn.attn_fwd(qkv, gate, table, klen, B, H, T, 0.125)
        torch.cuda.synchronize()
        print(mode, (B,H,T), "fwd nan: out", torch.isnan(out.float()).sum().item(), "lse", torch.isnan(lse).sum().item(), "lse inf", torch.isinf(lse).sum().item())
        if mode in ("poison_all", "poison_bwd_only"): poison()
        dqkv, dgate, dtable = Kn.attn_bwd(qkv, out, dout, lse, gate, table, klen, B, H, T, 0.125)
        torch.cuda.synchronize()
        nz = torch.isnan(dgate).nonzero()
        print("   bwd nan: dqkv", torch.isnan(dqkv.float()).sum().item(), "dgate", torch.isnan(dgate).sum().item(), "dtable", torch.isnan(dtable).sum().item(), "first", nz[:6].tolist(), "klen", klen.tolist())
