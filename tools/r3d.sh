#!/bin/bash
timeout 600 python -m pytest tests/test_mlp_fused_gpu.py -x -q -m gpu > gpurun_out/r3d_mlp.log 2>&1; echo "mlp tests rc=$?"; tail -25 gpurun_out/r3d_mlp.log
