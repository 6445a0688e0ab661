#!/bin/bash
timeout 150 python -m pytest tests/test_mlp_fused_gpu.py -q -m gpu -x -k "pair" 2>&1 | tail -8
timeout 100 python tools/mlp_pair_time.py 2>&1 | tail -6
