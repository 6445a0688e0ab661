#!/bin/bash
for v in 1 5 9 1 5 9; do echo "respf=$v"; GECCO_MLP_RESPF=$v timeout 100 python tools/mlp_pair_time.py 2>&1 | grep "==" ; done
