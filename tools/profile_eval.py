"""A few denoiser evaluations of the bench workload (B = 64 conditional clouds, 2048 points) for ncu captures.

    ncu --set full --import-source on -k regex:gemm_tc -s 50 -c 8 -o gpurun_out/gemm python tools/profile_eval.py
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
import gecco_b200 as G  # noqa: E402


def main(evals: int = 2, config: int = 2):
    dev = torch.device("cuda:0")
    cfg = bench.CONFIGS[config]
    clouds = cfg["clouds"]
    model = bench.build_model(dev, cfg)
    g = torch.Generator("cpu").manual_seed(123)
    ctx = G.Context3d(image=torch.rand(clouds, 3, cfg["image"], cfg["image"], generator=g).to(dev),
                      K=torch.tensor(cfg["K"]).expand(clouds, 3, 3).contiguous().to(dev))
    post = model.conditioner(ctx)
    x = torch.randn(clouds, bench.POINTS, 3, generator=g).to(dev)
    sigma = torch.full((clouds,), 1.5, device=dev)
    for _ in range(evals):
        out = model(x, sigma, ctx, post)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 2)
