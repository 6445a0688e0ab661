python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k lookup 2>&1 | tail -5
python tools/lookup_time.py 2>&1 | tail -12
for cfg in "s2 5" "s4 51"; do set -- $cfg
ncu --set full --clock-control none --import-source on -k regex:lookup_staged --launch-skip $2 --launch-count 1 -o /tmp/lk_$1 -f python tools/lookup_time.py > gpurun_out/ncu_lk_$1.log 2>&1
ncu -i /tmp/lk_$1.ncu-rep --page raw --csv > gpurun_out/lk_$1_raw.csv 2>/dev/null
ncu -i /tmp/lk_$1.ncu-rep --page source --csv > gpurun_out/lk_$1_source.csv 2>/dev/null
done
ls -la gpurun_out/lk_*
