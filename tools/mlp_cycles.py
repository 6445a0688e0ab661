"""Per-role cycle breakdown of the fused MLP kernel on the bench shape (development aid)."""
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
w1 = (torch.randn(B * H, C, generator=g) / math.sqrt(C)).to(dev).bfloat16()
b1 = torch.randn(B, H, generator=g).to(dev)
w2 = (torch.randn(C, H, generator=g) / math.sqrt(H)).to(dev).bfloat16()
b2 = torch.randn(C, generator=g).to(dev)
x = torch.randn(B * Np, C, device=dev)
xb = torch.empty(B * Np, C, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(B, C // 12, 2, dtype=torch.float64, device=dev)
enames = {16: "w0_tmem_ld_wait", 17: "w0_res_wait", 18: "w0_store_read_wait", 19: "w0_stage_sts", 20: "w0_stats"}
names = ["prod_total", "prod_wait_Afree", "prod_wait_W1free", "prod_wait_W2free", "mma_total", "mma_wait_W1", "mma_wait_A",
         "mma_wait_hready", "mma_wait_yempty", "mma_wait_W2", "act0_total", "act0_wait_hfull", "act0_wait_hcempty",
         "act0_wait_yfull", "act0_epilogue", "-", "w0_tmem_ld_wait", "w0_res_wait", "w0_store_read_wait", "w0_stage_sts",
         "w0_stats", "mma_wait_hfree", "-", "-", "-", "act0_tmem_ld_release", "act0_ld_plus_math", "-", "-", "w0_fence",
         "w0_tma_store_issue"]


def run():
    ops.mlp(a, w1, b1, 1.3, w2, b2, x, out_f32=x, out_bf16=xb, stats=stats, rows_per_cloud=Np, valid_rows=Np,
            w1_rows_per_cloud=H, b1_stride=H)


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
dbg = torch.zeros(148, 32, dtype=torch.int64, device=dev)
lib.gecco_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
run()
torch.cuda.synchronize()
lib.gecco_set_debug_buffer(ctypes.c_void_p(0))
d = dbg.cpu().double()
lead, peer = d[0::2], d[1::2]
print(f"== mlp_fused: {us:.1f} us/launch, {4 * B * Np * C * H / us / 1e6:.0f} TFLOP/s, {B * Np * C * 12 / us / 1e3:.0f} GB/s algorithmic")
print("   leader:", {n: int(lead[:, i].mean().item()) for i, n in enumerate(names) if n != "-"})
print("   peer  :", {n: int(peer[:, i].mean().item()) for i, n in enumerate(names) if n != "-"})
