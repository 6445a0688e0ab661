"""Per-role cycle breakdown of the CTA-pair GEMM on the bench shapes (development aid)."""
import ctypes, math, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import _abi, ops

dev = torch.device("cuda:0")
lib = _abi.init(0)
B, Np, K = 64, 2048, 384
g = torch.Generator("cpu").manual_seed(0)
a = torch.randn(B * Np, K, generator=g).to(dev).bfloat16()
names = ["prod_total", "prod_wait_Afree", "prod_wait_Wfree", "mma_total", "mma_wait_acc", "mma_wait_A", "mma_wait_W", "tiles",
         "epi0_total", "epi0_wait_acc", "epi1_total", "epi1_wait_acc", "mma_issue", "chunk_cycles", "chunks", "prefetch",
         "w0_tmem_ld_wait", "w0_res_wait", "w0_store_read_wait", "w0_stage_sts", "w0_stats", "-", "xf_total", "xf_wait_Aempty",
         "xf_wait_stage", "xf_norm"] + ["-"] * 3 + ["w0_fence", "w0_tma_store_issue"]
SKIP = int(os.environ.get("EPI_SKIP", "0"))
lib.gecco_set_option(ctypes.c_char_p(b"epi_skip"), SKIP)
print("epi_skip", SKIP)
xf = torch.randn(B * Np, K, generator=g).to(dev)
stats = ops.group_stats(xf, Np, Np, 12)
tn = torch.randn(B, generator=g).to(dev)
nw = [torch.randn(K, 1, generator=g).to(dev), torch.randn(K, generator=g).to(dev), torch.randn(K, 1, generator=g).to(dev), torch.randn(K, generator=g).to(dev)]
AN = dict(stats=stats, t=tn, scale_w=nw[0], scale_b=nw[1], bias_w=nw[2], bias_b=nw[3], groups=32)
CASES = [("kvq", 1152, True, None, False, False), ("kvq_anorm", 1152, False, None, False, True), ("mlp_up", 768, True, 1.3, False, False),
         ("mlp_up_anorm", 768, False, 1.3, False, True), ("unpool_out", 384, False, None, True, False)]
only = os.environ.get("CASES")
for label, n_out, percloud, act, res, anorm in CASES:
    if only and label not in only.split(","):
        continue
    w = (torch.randn((B if percloud else 1) * n_out, K, generator=g) / math.sqrt(K)).to(dev).bfloat16()
    bias = torch.randn(B if percloud else 1, n_out, generator=g).to(dev)
    if not percloud:
        bias = bias[0].contiguous()
    out = torch.empty(B * Np, n_out, device=dev, dtype=torch.bfloat16)
    x = torch.randn(B * Np, n_out, device=dev) if res else None
    dbg = torch.zeros(148, 32, dtype=torch.int64, device=dev)
    def run():
        ops.gemm(xf.bfloat16() if anorm else a, w, bias=bias, bias_stride=n_out if percloud else 0, act_alpha=act, out_bf16=out, rows_per_cloud=Np,
                 valid_rows=Np, w_rows_per_cloud=n_out if percloud else 0, n_out=n_out, res=x, out_f32=x, anorm=AN if anorm else None)
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
    lib.gecco_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.gecco_set_debug_buffer(ctypes.c_void_p(0))
    d = dbg.cpu().double()
    lead, peer = d[0::2], d[1::2]
    print(f"== {label}: {us:.1f} us/launch, {2 * B * Np * n_out * K / us / 1e6:.0f} TFLOP/s")
    print("   leader:", {n: int(lead[:, i].mean().item()) for i, n in enumerate(names) if n != '-'})
    print("   peer  :", {n: int(peer[:, i].mean().item()) for i, n in enumerate(names) if i < 3 or i >= 8})
