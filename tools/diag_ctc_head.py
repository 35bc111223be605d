"""Where does the fp32-mode CTC-head gradient lose accuracy?  Stage-by-stage errors against float64 (GPU diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtasr_b200 import kernels as K, precise
from mtasr_b200.precise import _a, _b, R

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
torch.manual_seed(4)
V, D, T, B = 4099, 1024, 150, 4
lin = torch.nn.Linear(D, V).to(dev)
with torch.no_grad():
    lin.weight.mul_(3.0)
hs = torch.randn(B, T, D, device=dev)
hlens = torch.tensor([T, T - 7, T // 2, 5], device=dev)
Lmax = 14
ys = torch.randint(0, V - 2, (B, Lmax), device=dev); ys[:, 3] = ys[:, 2]
ylens = torch.tensor([Lmax, 9, 0, 12], device=dev)
up = torch.rand(B, device=dev) + 0.5
blank = V - 1
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

# float64 truth
h64 = hs.double().requires_grad_(True)
logits64 = torch.nn.functional.linear(h64, lin.weight.double(), lin.bias.double())
logits64.retain_grad()
lp64 = logits64.transpose(0, 1).log_softmax(2)
tgt = torch.cat([ys[i, :l] for i, l in enumerate(ylens)])
nll64 = torch.nn.functional.ctc_loss(lp64, tgt, hlens, ylens, blank=blank, reduction="none", zero_infinity=True)
(nll64 * up.double()).sum().backward()
dl64 = logits64.grad                                     # (B,T,V) exact dlogits
lse64 = torch.logsumexp(logits64.detach(), -1)

hs2 = hs.view(B * T, D)
A, Bw = _a(hs2), _b(lin.weight)
bf = lin.bias.detach().float()
Vp = (V + 7) // 8 * 8
lg = torch.empty(B * T, Vp, device=dev)
K.gemm(K.Operand(A, R * D), K.Operand(Bw, R * D), B * T, V, R * D, K.Out(lg, Vp), bias=bf)
print("dense logits (split GEMM) vs f64: rel", rel(lg[:, :V], logits64.detach().view(B * T, V)), "max abs", (lg[:, :V].double() - logits64.detach().view(B * T, V)).abs().max().item())
print("torch fp32 logits vs f64: max abs", (lin(hs).double() - logits64.detach()).abs().max().item())
lse, _ = precise._head_lse_argmax(A, Bw, bf, B * T, V, D)
print("lse (mode-1 epilogue) vs f64: max abs", (lse.double() - lse64.view(-1)).abs().max().item(), " torch fp32 lse:", (torch.logsumexp(lin(hs), -1).double() - lse64).abs().max().item())
Lp = (Lmax + 1 + 63) // 64 * 64
ysc, hl, yl = ys.contiguous(), hlens.long().contiguous(), ylens.long().contiguous()
wg, bg = K.ctc_gather_rows(Bw, bf, ysc, yl, Lp, blank)
glog = torch.empty(B, T, Lp, device=dev)
K.gemm(K.Operand(A, R * D, sb0=T * R * D), K.Operand(wg, R * D, sb0=Lp * R * D), T, Lp, R * D, K.Out(glog, Lp, sb0=T * Lp), batch=(B, 1), bias=bg, bias_sb0=Lp)
g64 = K.ctc_gather_cols(logits64.detach().float().contiguous(), ysc, yl, Lp, blank)   # fp32 rounding of the exact logits
print("lattice logits (gathered-row split GEMM) vs f64: max abs", (glog - g64).abs().max().item())

def run_ctc(glog_, lse_):
    nll, nll_raw, alpha, coff = K.ctc_alpha_fwd(glog_, lse_.view(B, T).contiguous(), ysc, hl, yl, Lmax)
    dG, rowscale = K.ctc_beta_bwd(glog_, lse_.view(B, T).contiguous(), ysc, hl, yl, Lmax, alpha, coff, nll_raw, up)
    return nll, dG, rowscale

# occupancy term in isolation: exact (fp32-rounded) inputs vs our inputs
P64 = torch.softmax(logits64.detach(), -1)
valid = (torch.arange(T, device=dev)[None] < hlens[:, None]) & (nll64.detach() > 0)[:, None]
dense64 = P64 * (up.double()[:, None, None] * valid[..., None])
sparse64 = dl64 - dense64                                  # = -gamma * up on lattice columns
for name, gl, ls in (("exact inputs", g64, lse64.float()), ("our inputs", glog, lse)):
    nll, dG, rowscale = run_ctc(gl.contiguous(), ls.contiguous())
    dense = torch.zeros(B, T, V, device=dev)
    K.ctc_scatter_cols(dG, ysc, yl, dense, blank)
    print(f"[{name}] nll rel {rel(nll, nll64.detach())}  sparse (occupancy) term rel {rel(dense, sparse64)}")
dl = lg.clone()
K.softmax_scale_f32_(dl, lse, rowscale.view(-1), V)
print("dense softmax term rel", rel(dl[:, :V], dense64.view(B * T, V)))
K.ctc_scatter_cols(dG, ysc, yl, dl.view(B, T, Vp), blank)
print("full dlogits rel", rel(dl.view(B, T, Vp)[..., :V], dl64), " db rel", rel(dl[:, :V].sum(0), dl64.sum((0, 1))))
dw64 = dl64.view(B * T, V).t() @ hs2.double()
print("dW from OUR dl with exact f64 GEMM:", rel(dl[:, :V].double().t() @ hs2.double(), dw64), " with split GEMM:", rel(precise._wgrad32(dl, hs2)[:V], dw64))
print("dW from EXACT dl (fp32-rounded) with split GEMM:", rel(precise._wgrad32(torch.nn.functional.pad(dl64.view(B * T, V).float(), (0, Vp - V)).contiguous(), hs2)[:V], dw64))
