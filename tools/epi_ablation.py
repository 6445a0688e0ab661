"""Epilogue ablations (development aid): time the pair GEMMs and the fused MLP with parts of the shared epilogue
switched off through gecco_set_option("epi_skip", mask): 1 no output stores, 2 no residual, 4 no proxy fence,
8 no statistics atomics.  Results are wrong when mask != 0; only the timing is of interest."""
import ctypes, math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import _abi, ops

dev = torch.device("cuda:0")
lib = _abi.init(0)
B, Np, C, H = 64, 2048, 384, 768
g = torch.Generator("cpu").manual_seed(0)
a = torch.randn(B * Np, C, generator=g).to(dev).bfloat16()
stats = torch.zeros(B, C // 12, 2, dtype=torch.float64, device=dev)
x = torch.randn(B * Np, C, device=dev)
xb = torch.empty(B * Np, C, device=dev, dtype=torch.bfloat16)
w1 = (torch.randn(B * H, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
b1 = torch.randn(B, H, generator=g).to(dev)
w2 = (torch.randn(C, H, generator=g) / math.sqrt(H)).to(dev).bfloat16()
b2 = torch.randn(C, generator=g).to(dev)
wkvq = (torch.randn(B * 1152, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
bkvq = torch.randn(B, 1152, generator=g).to(dev)
okvq = torch.empty(B * Np, 1152, device=dev, dtype=torch.bfloat16)
wo = (torch.randn(C, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
hid = torch.empty(B * Np, H, device=dev, dtype=torch.bfloat16)


def t_us(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n


cases = {
    "kvq": lambda: ops.gemm(a, wkvq, bias=bkvq, bias_stride=1152, out_bf16=okvq, rows_per_cloud=Np, valid_rows=Np,
                            w_rows_per_cloud=1152, n_out=1152),
    "mlp_up": lambda: ops.gemm(a, w1, bias=b1, bias_stride=H, act_alpha=1.3, out_bf16=hid, rows_per_cloud=Np, valid_rows=Np,
                               w_rows_per_cloud=H, n_out=H),
    "unpool_out+stats": lambda: ops.gemm(a, wo, bias=b2, res=x, out_f32=x, out_bf16=xb, stats=stats, rows_per_cloud=Np,
                                         valid_rows=Np),
    "mlp_fused": lambda: ops.mlp(a, w1, b1, 1.3, w2, b2, x, out_f32=x, out_bf16=xb, stats=stats, rows_per_cloud=Np,
                                 valid_rows=Np, w1_rows_per_cloud=H, b1_stride=H),
}
masks = [int(m) for m in sys.argv[1:]] or [0, 1, 2, 4, 8, 3, 15]
print("us/launch by epi_skip mask:", masks)
for name, fn in cases.items():
    row = []
    for m in masks:
        lib.gecco_set_option(ctypes.c_char_p(b"epi_skip"), m)
        row.append(t_us(fn))
        x.normal_()
    lib.gecco_set_option(ctypes.c_char_p(b"epi_skip"), 0)
    print(f"{name:18s}", "  ".join(f"{v:7.1f}" for v in row))
