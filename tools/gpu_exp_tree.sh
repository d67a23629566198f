#!/bin/bash
# validation of a front-end change: extractor parity tests + committed vectors + launch list of one extract
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_extractor_gpu.py tests/test_pipeline_golden.py -x -q -m gpu 2>&1 | tail -4
ncu --clock-control none --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/exp_tree_launches.csv python tools/profile_step.py extract 3 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/exp_tree_launches.csv 2>/dev/null | head -12
