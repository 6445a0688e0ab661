"""Weight-gradient GEMMs of the training step (dW = dY^T X, reduction over M = 65536 rows): library mm vs a manual
split over row slabs with bmm (more CTAs in flight for the 18..36 output tiles)."""
import torch
dev = torch.device("cuda:0")
M = 65536
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for N, K in ((768, 384), (384, 384), (384, 768), (1152, 384)):
    dy = torch.randn(M, N, device=dev).bfloat16(); x = torch.randn(M, K, device=dev).bfloat16()
    ref = torch.mm(dy.t(), x, out_dtype=torch.float32)
    t0 = timed(lambda: torch.mm(dy.t(), x, out_dtype=torch.float32))
    row = [f"N={N} K={K}: mm {t0:.1f} us ({2*M*N*K/t0/1e6:.0f} TF/s)"]
    for S in (4, 8, 16, 32):
        f = lambda: torch.bmm(dy.view(S, M // S, N).transpose(1, 2), x.view(S, M // S, K), out_dtype=torch.float32).sum(0)
        try:
            out = f(); err = (out - ref).abs().max().item() / ref.abs().max().item()
            row.append(f"S={S}: {timed(f):.1f} us (err {err:.1e})")
        except Exception as e:
            row.append(f"S={S}: {type(e).__name__} {str(e)[:60]}")
    print("; ".join(row))
