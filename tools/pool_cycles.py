"""Per-role cycle counters of the tcgen05 pool attention kernel at the bench shape (needs a GECCO_DEBUG_COUNTERS=1 build:
GECCO_DEBUG_COUNTERS=1 python -m gecco_b200.build)."""
import ctypes, math, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import ops, _abi

dev = torch.device("cuda:0")
lib = _abi.init(0)
B, N, H, D, I = 64, 2048, 8, 48, 64
C = H * D
kv = torch.randn(B * N, 3 * C, device=dev).bfloat16()
qs = (torch.randn(H, I, D, device=dev) * (D**-0.5 * math.log2(math.e))).bfloat16().contiguous()
out = torch.empty(B * I, C, device=dev, dtype=torch.bfloat16)
run = lambda: ops.pool_attention(kv, qs, clouds=B, rows_per_cloud=N, valid_rows=N, heads=H, head_dim=D, k_off=0, v_off=C, splits=1, out=out)
for _ in range(3):
    run()
torch.cuda.synchronize()
dbg = torch.zeros(148, 32, dtype=torch.int64, device=dev)
lib.gecco_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
run()
torch.cuda.synchronize()
lib.gecco_set_debug_buffer(ctypes.c_void_p(0))
d = dbg.cpu().double()[:128]
for w in range(2):
    print(f"softmax wg{w}: " + ", ".join(f"{n}={d[:, w * 8 + i].mean().item():.0f}" for i, n in enumerate(
        ["wait_S", "softmax+P_store", "wait_O", "O_read", "loop_total", "units"])))
print("mma warp : " + ", ".join(f"{n}={d[:, 16 + i].mean().item():.0f}" for i, n in enumerate(["issue_S(incl waits)", "issue_PV(incl waits)", "loop_total"])))
