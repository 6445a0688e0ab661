"""Stand-alone timing of the pool attention core at the bench shape (64 clouds x 2048 points, 8 heads x 48): mma.sync
split-KV kernel + combine (GECCO_POOL_TC=0) against the tcgen05 / TMEM kernel.  GB/s on the compulsory traffic (k, v once)."""
import math, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import ops

dev = torch.device("cuda:0")
B, N, H, D, I = 64, 2048, 8, 48, 64
C = H * D
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


kv = torch.randn(B * N, 3 * C, device=dev).bfloat16()
qs = (torch.randn(H, I, D, device=dev) * (D**-0.5 * math.log2(math.e))).bfloat16().contiguous()
out = torch.empty(B * I, C, device=dev, dtype=torch.bfloat16)
nbytes = B * N * 2 * C * 2
ref = None
for tc, splits in (("0", 3), ("1", 3), ("1", 1)):
    os.environ["GECCO_POOL_TC"] = tc
    fn = lambda: ops.pool_attention(kv, qs, clouds=B, rows_per_cloud=N, valid_rows=N, heads=H, head_dim=D, k_off=0, v_off=C,
                                    splits=splits, out=out)
    fn()
    torch.cuda.synchronize()
    if ref is None:
        ref = out.float().clone()
    err = (out.float() - ref).abs().max().item()
    hot, cold = t_us(fn), t_us(fn, cold=True)
    print(f"pool attention tc={tc} splits<={splits}: hot L2 {hot:7.1f} us ({nbytes / hot / 1e3:7.1f} GB/s)   cold L2 {cold:7.1f} us "
          f"({nbytes / cold / 1e3:7.1f} GB/s)   max |diff vs mma.sync| {err:.2e}")
