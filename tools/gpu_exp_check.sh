#!/bin/bash
# quick validation of the optimiser tests + committed vectors + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optimize_gpu.py tests/test_pipeline_golden.py tests/test_class_layer.py -x -q -m gpu 2>&1 | tail -4
ncu --clock-control none --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/exp_chk_launches.csv python tools/profile_step.py ba 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/exp_chk_launches.csv 2>/dev/null | grep -E "k_ba_pair|launches"
