"""Builds gecco_b200/lib/libgecco_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Run as ``python -m gecco_b200.build`` or through ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU.  Objects are rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OBJ = ROOT / "csrc" / "build"
LIB = ROOT / "lib" / "libgecco_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return nvcc


def _newest_header() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list((ROOT.parent / "include").glob("*.h"))
    return max(h.stat().st_mtime for h in hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
    flags = list(NVCC_FLAGS)
    marker = OBJ / ".debug_counters"
    debug = os.environ.get("GECCO_DEBUG_COUNTERS", "0") == "1"
    OBJ.mkdir(parents=True, exist_ok=True)
    if debug:
        flags.append("-DGECCO_DEBUG_COUNTERS=1")
    if debug != marker.exists():  # switching between the production and the instrumented build recompiles everything
        force = True
        marker.touch() if debug else marker.unlink()
    LIB.parent.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(CSRC.glob("*.cu"))
    hdr_t = _newest_header()
    jobs = []
    objs = []
    for src in srcs:
        obj = OBJ / (src.stem + ".o")
        objs.append(obj)
        if force or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, hdr_t):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, res

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, res in ex.map(compile_one, jobs):
                log = OBJ / (src.stem + ".log")
                log.write_text(res.stdout + res.stderr)
                if verbose or res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                if res.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src.name}")
    if jobs or not LIB.exists():
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link of libgecco_b200.so failed")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
