#!/bin/bash
# round-2 first GPU call: baseline tests, sanitizer evidence, fused-MLP cycle counters
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
# sanitizer (racecheck / synccheck) over the tcgen05 / mbarrier kernels, bounded
for tool in racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gemm_gpu.py tests/test_mlp_fused_gpu.py "tests/test_ops_gpu.py::test_pool_attention" "tests/test_ops_gpu.py::test_unpool_attention" -m gpu -x -q > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  tail -5 gpurun_out/r2a_sanitizer_$tool.log
done
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_denoiser_gpu.py -m gpu -x -q -k "uncond or cond_gaussian" > gpurun_out/r2a_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2a_sanitizer_memcheck.log
# fused MLP per-role cycles (instrumented build on the box)
GECCO_DEBUG_COUNTERS=1 python -m gecco_b200.build > /dev/null 2>&1
python tools/mlp_cycles.py > gpurun_out/r2a_mlp_cycles.log 2>&1; cat gpurun_out/r2a_mlp_cycles.log
python tools/gemm_cycles.py > gpurun_out/r2a_gemm_cycles.log 2>&1; tail -20 gpurun_out/r2a_gemm_cycles.log
