#!/bin/bash
timeout 900 python -m pytest tests/test_training_gpu.py -x -q -s > gpurun_out/r4c_train.log 2>&1; echo "train tests rc=$?"; grep -E "loss|norm|cosine|rel rms|passed|failed|Error|error" gpurun_out/r4c_train.log | tail -30
