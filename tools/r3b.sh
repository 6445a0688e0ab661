#!/bin/bash
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q -m gpu > gpurun_out/r3b_chain.log 2>&1; echo "chain tests rc=$?"; tail -5 gpurun_out/r3b_chain.log
python tools/chain_time.py 2>&1 | tail -12
