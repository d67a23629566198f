#!/bin/bash
# validation of a Schur-kernel change: optimiser parity tests + committed vectors, launch list of one BA, one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optimize_gpu.py tests/test_pipeline_golden.py tests/test_class_layer.py -x -q -m gpu 2>&1 | tail -3
ncu --clock-control none --metrics gpu__time_duration.sum -c 160 --csv --log-file gpurun_out/exp_schur_launches.csv python tools/profile_step.py ba 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/exp_schur_launches.csv 2>/dev/null | grep -E "k_ba_schur|k_ba_chol|launches"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/exp_schur_bench.json 2> gpurun_out/exp_schur_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/exp_schur_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ba ms", d["value_stage_ms_per_frame_stream0"]["local_ba"], "schur", d["roofline_schur"]["avg_launch_us"], d["roofline_schur"]["frac"])
PY
