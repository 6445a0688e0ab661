"""Launches one of the bench-shape tcgen05 kernels a few times (for ncu captures):  python tools/one_gemm.py kvq|mlp_up|unpool_out|mlp_fused"""
import math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "kvq"
dev = torch.device("cuda:0")
B, Np, C, H = 64, 2048, 384, 768
g = torch.Generator("cpu").manual_seed(0)
a = torch.randn(B * Np, C, generator=g).to(dev).bfloat16()
stats = torch.zeros(B, C // 12, 2, dtype=torch.float64, device=dev)
x = torch.randn(B * Np, C, device=dev)
xb = torch.empty(B * Np, C, device=dev, dtype=torch.bfloat16)
if which == "kvq":
    w = (torch.randn(B * 1152, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
    b = torch.randn(B, 1152, generator=g).to(dev)
    o = torch.empty(B * Np, 1152, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, bias=b, bias_stride=1152, out_bf16=o, rows_per_cloud=Np, valid_rows=Np, w_rows_per_cloud=1152, n_out=1152)
elif which == "mlp_up":
    w = (torch.randn(B * H, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
    b = torch.randn(B, H, generator=g).to(dev)
    o = torch.empty(B * Np, H, device=dev, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, bias=b, bias_stride=H, act_alpha=1.3, out_bf16=o, rows_per_cloud=Np, valid_rows=Np, w_rows_per_cloud=H, n_out=H)
elif which == "unpool_out":
    w = (torch.randn(C, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
    b = torch.randn(C, generator=g).to(dev)
    fn = lambda: ops.gemm(a, w, bias=b, res=x, out_f32=x, out_bf16=xb, stats=stats, rows_per_cloud=Np, valid_rows=Np)
else:
    w1 = (torch.randn(B * H, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
    b1 = torch.randn(B, H, generator=g).to(dev)
    w2 = (torch.randn(C, H, generator=g) / math.sqrt(H)).to(dev).bfloat16()
    b2 = torch.randn(C, generator=g).to(dev)
    fn = lambda: ops.mlp(a, w1, b1, 1.3, w2, b2, x, out_f32=x, out_bf16=xb, stats=stats, rows_per_cloud=Np, valid_rows=Np,
                         w1_rows_per_cloud=H, b1_stride=H)
for _ in range(4):
    fn()
torch.cuda.synchronize()
