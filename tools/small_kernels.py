"""Stand-alone timings of the non-GEMM kernels at the bench shape (development aid): with a hot L2 (back to back) and
after a 512 MB streaming write that evicts L2 (the state the kernels see inside an evaluation)."""
import math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import ops

dev = torch.device("cuda:0")
B, N, C, H = 64, 2048, 384, 768
g = torch.Generator("cpu").manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def t_us(fn, n=20, cold=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if cold:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot * 1000 / n


x = torch.randn(B * N, C, device=dev)
stats = ops.group_stats(x, N, N, 12)
t = torch.randn(B, device=dev)
sw, sb, bw, bb = (torch.randn(C, 1, device=dev), torch.randn(C, device=dev), torch.randn(C, 1, device=dev), torch.randn(C, device=dev))
w_kvq = torch.randn(3 * C, C, device=dev) / math.sqrt(C)
b_kvq = torch.randn(3 * C, device=dev)
w_up = torch.randn(H, C, device=dev) / math.sqrt(C)
b_up = torch.randn(H, device=dev)
cases = {
    "fold_adagn kvq (1152 rows)": lambda: ops.fold_adagn(w_kvq, b_kvq, stats, t, sw, sb, bw, bb, clouds=B, valid_rows=N),
    "fold_adagn mlp.0 (768 rows)": lambda: ops.fold_adagn(w_up, b_up, stats, t, sw, sb, bw, bb, clouds=B, valid_rows=N),
}
w_out = torch.randn(3, C, device=dev) / math.sqrt(C)
b_out = torch.randn(3, device=dev)
xin = torch.randn(B, N, 3, device=dev)
sig = torch.full((B,), 2.0, device=dev)
out = torch.empty(B, N, 3, device=dev)
cases["head (GroupNorm16, mode 1)"] = lambda: ops.head(x, w_out, b_out, clouds=B, rows_per_cloud=N, valid_rows=N, norm=2, groups=16,
                                                       stats=stats, xin=xin, sigma=sig, mode=1, out=out)
h = torch.randn(B * 64, C, device=dev)
hst = ops.group_stats(h, 64, 64, 12)
hb = torch.empty(B * 64, C, device=dev, dtype=torch.bfloat16)
cases["adagn (inducers)"] = lambda: ops.adagn(h, hst, 12, t, sw, sb, bw, bb, rows_per_cloud=64, valid_rows=64, out_bf16=hb)
levels = [torch.randn(B, s, s, c, device=dev).bfloat16() for s, c in ((34, 96), (17, 192), (8, 384))]
K = torch.tensor([[1.0859, 0, 0.4964], [0, 1.0859, 0.4964], [0, 0, 1]], device=dev).expand(B, 3, 3).contiguous()
lo = torch.empty(B * N, 672, device=dev, dtype=torch.bfloat16)
lst = torch.zeros(B, 16, 2, dtype=torch.float64, device=dev)
cases["lookup (137^2 pyramid)"] = lambda: ops.lookup(xin, levels, K, reparam_kind=1, mean=[0, 0, 1], sigma_r=[0.15] * 3, sigma=sig,
                                                     rows_per_cloud=N, out_bf16=lo, stats=lst)
for name, fn in cases.items():
    print(f"{name:32s} hot L2 {t_us(fn):7.1f} us   cold L2 {t_us(fn, cold=True):7.1f} us")
