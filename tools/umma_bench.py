"""Cycles per tcgen05.mma (M128 x N x K16, bf16) on one SM for the operand sources the attention kernels can use
(gecco_debug_umma_bench): SS / TS mode, K-major / MN-major B.  Floor: N / 2 cycles (8192 FLOP per cycle per SM)."""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from gecco_b200 import _abi

lib = _abi.init(0)
lib.gecco_debug_umma_bench.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda:0")
names = {0: "SS  B K-major ", 1: "SS  B MN-major", 2: "TS  B K-major ", 3: "TS  B MN-major"}
for batch in (8, 64):
    for mode in (0, 1, 2, 3):
        row = []
        for n in (48, 64, 96, 128, 192, 256):
            if (mode & 1) and n > 128:
                continue
            _abi.check(lib.gecco_debug_umma_bench(mode, n, batch, 20, C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            torch.cuda.synchronize()
            best = out[0].item()
            row.append(f"N={n}: {best / batch:6.1f} (floor {n / 2:.0f})")
        print(f"batch {batch:3d} {names[mode]}  cycles/MMA incl. commit+wake: " + "  ".join(row))
