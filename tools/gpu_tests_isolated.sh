#!/bin/bash
# Runs every GPU test node in its own process so that one CUDA context failure does not
# cascade into the following tests.  Usage: tools/gpu_tests_isolated.sh tests/test_ops_gpu.py [...]
ids=$(python -m pytest "$@" -m gpu --collect-only -q 2>/dev/null | grep '::')
fail=0
for id in $ids; do
  out=$(timeout 300 python -m pytest "$id" -x -q -m gpu 2>&1)
  rc=$?
  if [ $rc -ne 0 ]; then
    fail=$((fail+1))
    echo "=== FAIL $id"
    echo "$out" | grep -E "^E |Error|assert" | head -12
  else
    echo "ok   $id"
  fi
done
echo "failed: $fail"
