"""Host cost of one kernel launch through the engine: a 2-step sampler call (3 evaluations, < 1024 launches, so the
CUDA launch queue never fills and the call returns at host speed) timed until it returns and until the GPU is done."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
import gecco_b200 as G
from gecco_b200 import _abi

dev = torch.device("cuda:0")
model = bench.build_model(dev)
B = bench.CLOUDS_PER_GPU
g = torch.Generator("cpu").manual_seed(123)
ctx = G.Context3d(image=torch.rand(B, 3, bench.IMAGE, bench.IMAGE, generator=g).to(dev),
                  K=torch.tensor(bench.K_CAM).expand(B, 3, 3).contiguous().to(dev))
lib = _abi.init(0)
for steps in (2, 2, 2, 64):
    torch.cuda.synchronize()
    n0 = lib.gecco_launch_count(0)
    t0 = time.perf_counter()
    out = model.sample_stochastic((B, bench.POINTS, 3), ctx, rng=torch.Generator(dev).manual_seed(42), num_steps=steps)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n = lib.gecco_launch_count(0) - n0
    print(f"num_steps={steps}: {n} engine launches, call returned after {1e3 * (t1 - t0):.1f} ms "
          f"({1e6 * (t1 - t0) / max(n, 1):.1f} us per launch incl. conditioner + python), GPU done after {1e3 * (t2 - t0):.1f} ms")
