#!/usr/bin/env python
"""Benchmark of the gecco sampling hot path (BASELINE.json: point clouds/sec, 2048 points, full EDM sampler).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): ShapeNet-vol style image-conditional model — RayNetwork + SetTransformer
(6 layers, C=384, 64 inducers, 8 heads) + GaussianReparam, ConvNeXt-T conditioner, synthetic 3x137x137 images and
cameras, 2048 points, 64 clouds per GPU, random-init weights.  One "step" = one `Diffusion.sample_stochastic`
call on the batch: conditioner once + 64 stochastic EDM steps = 127 denoiser evaluations.

Own arm: `value` has the context resident in HBM; `e2e` starts from pinned host images / cameras and ends with the
sampled clouds back on the host.  Under torchrun every rank samples its own 64 clouds (weak scaling, no data-path
collective); time = max over ranks.  `roofline` is the tensor roofline of the dominant kernel class (the tcgen05
projection GEMMs), timed with CUDA events inside this script by the engine's per-kernel-class profiler.
`--impl reference`: the CPU restatement of the reference (oracle/gecco_oracle.py; the Python reference itself
cannot travel to the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

CLOUDS_PER_GPU = 64
POINTS = 2048
IMAGE = 137
NUM_STEPS = 64
EVALS = 2 * NUM_STEPS - 1
# algorithmic FLOPs of one conditional denoiser evaluation per cloud: 33.266 GFLOP
C_, I_, L_, CTX_ = 384, 64, 6, 672
FLOP_PER_EVAL = L_ * (16 * POINTS * C_ * C_ + 8 * POINTS * I_ * C_ + 14 * I_ * C_ * C_) + 4 * 3 * POINTS * C_ + 2 * POINTS * CTX_ * C_
WORKLOAD = ("ShapeNet-vol image-conditional (RayNetwork+SetTransformer L6 C384 I64 H8, GaussianReparam, ConvNeXt-T conditioner), "
            f"{POINTS} points, synthetic 3x{IMAGE}x{IMAGE} images + cameras, {CLOUDS_PER_GPU} clouds per GPU, "
            f"{NUM_STEPS}-step stochastic EDM sampler ({EVALS} denoiser evaluations), random-init weights")
REPARAM = dict(mean=[0.0, 0.0, 1.0], sigma=[0.15, 0.15, 0.15])
K_CAM = [[1.0859, 0.0, 0.4964], [0.0, 1.0859, 0.4964], [0.0, 0.0, 1.0]]
SIGMA_MAX = 165.0


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ own arm
def build_model(device):
    import gecco_b200 as G
    from gecco_b200.models import ConvNeXtExtractor, GaussianActivation, RayNetwork, SetTransformer
    from gecco_b200.reparam import GaussianReparam

    torch.manual_seed(0)
    rp = GaussianReparam(torch.tensor(REPARAM["mean"]), torch.tensor(REPARAM["sigma"]))
    st = SetTransformer(n_layers=L_, num_inducers=I_, feature_dim=C_, t_embed_dim=1, num_heads=8, activation=GaussianActivation)
    net = RayNetwork(backbone=st, reparam=rp, context_dims=(96, 192, 384))
    model = G.Diffusion(backbone=G.EDMPrecond(model=net), conditioner=ConvNeXtExtractor(pretrained=False), reparam=rp,
                        loss=G.EDMLoss(schedule=G.LogUniformSchedule(max=SIGMA_MAX)))
    return model.to(device).eval()


def run_own(args):
    import torch.distributed as dist

    import gecco_b200 as G
    from gecco_b200 import engine as E

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    model = build_model(device)
    B = CLOUDS_PER_GPU
    g = torch.Generator("cpu").manual_seed(123 + rank)
    images_h = torch.rand(B, 3, IMAGE, IMAGE, generator=g).pin_memory()
    K_h = torch.tensor(K_CAM).expand(B, 3, 3).contiguous().pin_memory()
    images_d, K_d = images_h.to(device), K_h.to(device)
    out_h = torch.empty(B, POINTS, 3, dtype=torch.float64).pin_memory()
    rng = torch.Generator(device).manual_seed(42 + rank)
    ctx_d = G.Context3d(image=images_d, K=K_d)

    def step_resident():
        return model.sample_stochastic((B, POINTS, 3), ctx_d, rng=rng)

    def step_e2e():
        ctx = G.Context3d(image=images_h.to(device, non_blocking=True), K=K_h.to(device, non_blocking=True))
        out = model.sample_stochastic((B, POINTS, 3), ctx, rng=rng)
        out_h.copy_(out, non_blocking=True)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return ms.item()

    for _ in range(args.warmup):
        out = step_resident()
    assert torch.isfinite(out).all(), "non-finite samples"
    clocks = ClockSampler(local) if rank == 0 else None
    E.launch_count(reset=True)
    ms = timed(step_resident, args.steps)
    launches = E.launch_count(reset=True)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clk = clocks.stop() if clocks else None

    # host time to enqueue one step (no synchronisation inside): if it approaches ms_per_step the run is launch-bound
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step_resident()
    host_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()

    # per-kernel-class device time of one more step, CUDA events on the launching stream
    E.profile_start()
    step_resident()
    prof = E.profile_stop()
    total_prof_ms = sum(p["ms"] for p in prof)
    gemm = [p for p in prof if p["name"].startswith("gemm_") or p["name"].endswith("_fused")]
    gemm_ms, gemm_flops = sum(p["ms"] for p in gemm), sum(p["flops"] for p in gemm)
    pk = peaks()
    # the dominant kernel: the CTA-pair tcgen05 GEMM of the k|v|q projection (most expensive launch of an evaluation)
    dom = next(p for p in prof if p["name"] == "gemm_kv_q")
    achieved_tf = dom["flops"] / (dom["ms"] * 1e-3) / 1e12 if dom["ms"] > 0 else 0.0
    look = next((p for p in prof if p["name"] == "lookup"), None)

    if rank == 0:
        clouds = world * B * args.steps
        value = clouds / (ms * 1e-3)
        traffic = None
        tfile = ROOT / "profiles" / "roofline_traffic.json"
        lookup_traffic = None
        if tfile.exists():
            tj = json.loads(tfile.read_text())
            traffic = tj.get("gemm_dram_bytes_per_launch")
            lookup_traffic = tj.get("lookup_dram_bytes_per_launch")
        line = {
            "metric": "point clouds/sec (2048 pts, full EDM sampler)", "value": value, "unit": "clouds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clouds_per_gpu": B, "points": POINTS, "num_steps": NUM_STEPS,
                       "l2": "no flush needed: every evaluation streams a >600 MB working set per GPU (L2 is 126 MB)",
                       "parallelism": f"dp{world} (independent clouds, no data-path collective)"},
            "e2e": {"value": clouds / (ms_e2e * 1e-3), "unit": "clouds/s",
                    "h2d_bytes_per_step": images_h.numel() * 4 + K_h.numel() * 4, "d2h_bytes_per_step": out_h.numel() * 8},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_ms,
            "clocks": clk,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved_tf / pk["tf_sustained"], "traffic": traffic,
                         "kernel": "gemm_pair_kernel (tcgen05 cta_group::2, k|v|q projection M=B*2048 N=1152 K=384, AdaGN folded "
                                   "into per-cloud weights); algorithmic FLOPs 2*M*N*K per launch",
                         "launches_timed": dom["launches"], "us_per_launch": dom["ms"] * 1e3 / dom["launches"],
                         "share_of_step": dom["ms"] / total_prof_ms if total_prof_ms else None,
                         "all_tcgen05_gemms": {"tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None,
                                               "share_of_step": gemm_ms / total_prof_ms if total_prof_ms else None},
                         "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                         "whole_path_tflops": world * B * EVALS * FLOP_PER_EVAL * args.steps / (ms * 1e-3) / 1e12,
                         "whole_path_frac_of_tensor_peak": B * EVALS * FLOP_PER_EVAL * args.steps / (ms * 1e-3) / 1e12 / pk["tf_sustained"],
                         "lookup_hbm": None if look is None else {
                             "achieved_gbs": look["bytes"] / (look["ms"] * 1e-3) / 1e9, "peak_gbs": pk["hbm"],
                             "frac": look["bytes"] / (look["ms"] * 1e-3) / 1e9 / pk["hbm"],
                             "kernel": "lookup_staged_kernel (shared-memory staged pyramid slices); algorithmic bytes = "
                                       "pyramid + bf16 output + coordinates per launch",
                             "us_per_launch": look["ms"] * 1e3 / look["launches"], "traffic": lookup_traffic}},
            "kernel_classes": [{"name": p["name"], "launches": p["launches"], "ms": round(p["ms"], 3),
                                "tflops": round(p["flops"] / (p["ms"] * 1e-3) / 1e12, 1) if p["ms"] > 0 else None,
                                "gbs": round(p["bytes"] / (p["ms"] * 1e-3) / 1e9, 1) if p["ms"] > 0 else None} for p in prof],
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference(budget_s=25.0)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def cpu_reference(budget_s: float, steps: int = 1, warmup: int = 0) -> dict:
    """The oracle restatement of the reference path on the host cores, bounded sample of the bench workload."""
    from oracle import gecco_oracle as O
    from tests import synth
    import torchvision.models as tvm

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleConfig(kind="cond", reparam="gaussian", sigma_max=SIGMA_MAX)
    sd = synth.full_state_dict("cond", "gaussian", REPARAM["mean"], REPARAM["sigma"], 1234)
    torch.manual_seed(0)
    feats_net = tvm.convnext_tiny(weights=None).features[:6].eval()  # stages 0-2 (models/feature_pyramid.py:46-53)
    B = 1
    img = torch.rand(B, 3, IMAGE, IMAGE, generator=torch.Generator().manual_seed(123))
    K = torch.tensor(K_CAM).expand(B, 3, 3).contiguous()

    def pyramid():
        with torch.no_grad():
            x, out = img, []
            for i in range(0, 6, 2):
                x = feats_net[i + 1](feats_net[i](x))
                out.append(x)
        return out

    # calibrate: one evaluation
    feats = pyramid()
    x = torch.randn(B, POINTS, 3)
    sg = torch.full((B,), 1.0)
    with torch.no_grad():
        O.denoise(cfg, sd, x, sg, feats, K)
        t0 = time.perf_counter()
        O.denoise(cfg, sd, x, sg, feats, K)
        t_eval = time.perf_counter() - t0
    n_steps = NUM_STEPS
    total_calls = steps + warmup
    while n_steps > 2 and total_calls * (2 * n_steps - 1) * t_eval > budget_s:
        n_steps //= 2
    times = []
    for i in range(total_calls):
        t0 = time.perf_counter()
        feats = pyramid()
        O.sample_stochastic(cfg, sd, (B, POINTS, 3), feats, K, rng=torch.Generator().manual_seed(42), num_steps=n_steps)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    scale = EVALS / (2 * n_steps - 1)  # extrapolation to the full 127 evaluations when the sample was shortened
    return {"value": B / (t * scale), "unit": "clouds/s", "cores": cores, "kind": "port",
            "sample": (f"{B} cloud(s) x {POINTS} points, conditioner + {n_steps}-step sampler ({2 * n_steps - 1} evaluations"
                       + ("" if n_steps == NUM_STEPS else f", extrapolated x{scale:.2f} to {EVALS}") + f"), fp32 torch CPU, {t:.1f} s per call"),
            "seconds_per_call": t}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    base = cpu_reference(budget_s=150.0, steps=steps, warmup=warmup)
    world = int(os.environ.get("WORLD_SIZE", 1))
    line = {"impl": "reference", "metric": "point clouds/sec (2048 pts, full EDM sampler)", "value": base["value"],
            "unit": "clouds/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": base["seconds_per_call"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference path (oracle port) on the host cores"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; gecco_b200 has no CPU path (use --impl reference for the CPU baseline)")
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", "29531", __file__, "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
               str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
        raise SystemExit(subprocess.call(cmd))
    run_own(args)


if __name__ == "__main__":
    main()
