"""Achieved HBM bandwidth of the training-path element-wise kernels (csrc/train_ops.cu) at the config-5 shape."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import ops

dev = torch.device("cuda:0")
B, N = 32, 2048


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


for C in (384, 768):
    x = torch.randn(B, N, C, device=dev)
    dy = torch.randn(B, N, C, device=dev)
    p, q, r = (torch.randn(B, C, device=dev) for _ in range(3))
    alpha = torch.tensor(1.3, device=dev)
    nbytes = x.numel() * 4
    rows = [("affine (norm forward)", lambda: ops.train_affine(x, None, p, None, r), 2 * nbytes),
            ("affine2 (norm backward dx)", lambda: ops.train_affine(dy, x, p, q, r), 3 * nbytes),
            ("colsum2 (norm backward sums)", lambda: ops.train_colsum2(dy, x), 2 * nbytes),
            ("group_stats (norm forward sums)", lambda: ops.group_stats(x.view(B * N, C), N, N, 12), nbytes),
            ("gauss_act forward", lambda: ops.train_gauss_act_fwd(x, alpha), 2 * nbytes),
            ("gauss_act backward", lambda: ops.train_gauss_act_bwd(x, dy, alpha), 3 * nbytes)]
    for name, fn, by in rows:
        t = timed(fn)
        print(f"C={C:4d} {name:34s} {t * 1e6:7.1f} us  {by / t / 1e9:7.0f} GB/s")
