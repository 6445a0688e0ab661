#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/r2m_pytest.log | tail -6
GECCO_DEBUG_COUNTERS=1 python -m gecco_b200.build > /dev/null 2>&1
CASES=kvq_anorm,mlp_up_anorm python tools/gemm_cycles.py > gpurun_out/r2m_gemm_cycles.log 2>&1; cat gpurun_out/r2m_gemm_cycles.log | cut -c1-900
