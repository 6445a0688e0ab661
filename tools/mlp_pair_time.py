"""Per-role wait breakdown of the CTA-pair MLP (mlp_pair.cu) on the bench shape (development aid)."""
import ctypes, math, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import _abi, ops

dev = torch.device("cuda:0")
lib = _abi.init(0)
B, Np, C, HID = int(os.environ.get("CLOUDS", "64")), 2048, 384, 768
g = torch.Generator("cpu").manual_seed(0)
xf = (torch.randn(B * Np, C, generator=g)).to(dev)
xb = xf.bfloat16()
w1 = (torch.randn(HID, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
w2 = (torch.randn(C, HID, generator=g) / math.sqrt(HID)).to(dev).bfloat16()
b1, b2 = torch.randn(HID, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
t = torch.randn(B, generator=g).to(dev)
nw = [torch.randn(C, generator=g).to(dev) for _ in range(4)]
stats_in = ops.group_stats(xf, Np, Np, 12)
an = dict(stats=stats_in, t=t, scale_w=nw[0], scale_b=nw[1], bias_w=nw[2], bias_b=nw[3], groups=32)
stats = torch.zeros(B, C // 12, 2, dtype=torch.float64, device=dev)
x = xf.clone()
xo = torch.empty(B * Np, C, device=dev, dtype=torch.bfloat16)
def run():
    ops.mlp(xb, w1, b1, 1.3, w2, b2, x, out_f32=x, out_bf16=xo, stats=stats, rows_per_cloud=Np, valid_rows=Np, anorm=an)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f"== mlp_pair {B} x {Np}: {us:.1f} us/launch, {4 * B * Np * C * HID / us / 1e6:.0f} TFLOP/s")
dbg = torch.zeros(148, 32, dtype=torch.int64, device=dev)
lib.gecco_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
run()
torch.cuda.synchronize()
lib.gecco_set_debug_buffer(ctypes.c_void_p(0))
d = dbg.cpu().double()
names = ["prod_total", "prod_wait_W1slot", "prod_wait_Aslot", "prod_wait_hidden_ready", "prod_wait_Hslot", "prod_wait_W2slot",
         "mma_total", "mma_wait_accU", "mma_wait_W1", "mma_wait_A", "mma_wait_accD", "mma_wait_H", "mma_wait_W2",
         "epi_total", "epi_wait_accU", "epi_fast", "epi_store_complete", "epi_wait_accD", "epi_panel"]
lead, peer = d[0::2], d[1::2]
print("   leader:", {n: int(lead[:, i].mean().item()) for i, n in enumerate(names)})
print("   peer  :", {n: int(peer[:, i].mean().item()) for i, n in enumerate(names) if not n.startswith("mma")})
