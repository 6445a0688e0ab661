"""Stage timeline of the inducer chain kernel (development aid): globaltimer stamps of CTA 0..3 of the first cluster."""
import ctypes, math, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import _abi, ops

dev = torch.device("cuda:0")
lib = _abi.init(0)
C, HID = 384, 768
for clouds, splits in ((64, 3), (64, 1), (4, 1)):
    g = torch.Generator("cpu").manual_seed(0)
    bf = torch.bfloat16
    r = lambda *s, k=1.0: (torch.randn(*s, generator=g) * k).to(dev)
    pooled = r(clouds * 64, C).to(bf)
    w = [r(C, C, k=C**-0.5).to(bf), r(HID, C, k=C**-0.5).to(bf), r(C, HID, k=HID**-0.5).to(bf), r(2 * C, C, k=C**-0.5).to(bf)]
    n1 = [r(C), r(C), r(C), r(C)]
    n2 = [r(C), r(C), r(C), r(C)]
    t = r(clouds)
    partial = torch.rand(clouds * 8 * max(splits, 1) * 64 * 50, generator=g).to(dev) if splits > 1 else None
    def run():
        return ops.inducer_chain(pooled, w[0], n1, w[1], r(HID), 1.3, w[2], r(C), n2, w[3], r(2 * C), t, partial=partial, splits=splits)
    b0, b2, bkv = r(HID), r(C), r(2 * C)
    def run():
        return ops.inducer_chain(pooled, w[0], n1, w[1], b0, 1.3, w[2], b2, n2, w[3], bkv, t, partial=partial, splits=splits)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"== clouds {clouds} splits {splits}: {e0.elapsed_time(e1) * 50:.1f} us / launch (incl. torch.empty of the outputs)")
    dbg = torch.zeros(148, 32, dtype=torch.int64, device=dev)
    lib.gecco_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.gecco_set_debug_buffer(ctypes.c_void_p(0))
    d = dbg.cpu()
    names = {0: "start", 1: "combine done"}
    for s in range(4):
        names.update({2 + 4 * s: f"s{s} barrier passed", 20 + s: f"s{s} first k-block landed (MMA)", 3 + 4 * s: f"s{s} accumulator full",
                      4 + 4 * s: f"s{s} epilogue stored", 5 + 4 * s: f"s{s} proxy fence"})
    for cta in (0, 3):
        t0 = d[cta, 0].item()
        ev = sorted((d[cta, k].item() - t0, n) for k, n in names.items() if d[cta, k].item() > 0)
        print(f"   CTA {cta}: " + "; ".join(f"{n} {dt / 1000:.2f}" for dt, n in ev))
