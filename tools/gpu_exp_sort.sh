#!/bin/bash
# experiment: hand-written stable radix sort of the graph preparation (cub removed), parallel pair counts
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_optimize_gpu.py tests/test_class_layer.py tests/test_cabi.py -x -q 2>&1 | tail -8
ncu --clock-control none --metrics gpu__time_duration.sum -c 120 --csv --log-file gpurun_out/exp_sort_launches.csv python tools/profile_step.py ba 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/exp_sort_launches.csv 2>/dev/null | grep -E "k_sort|k_ba_pair|k_ba_emit|k_ba_seg|k_ba_chunk|k_ba_free|k_ba_landmark_index|launches"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_sort_bench.json 2> gpurun_out/exp_sort_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/exp_sort_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["value_stage_ms_per_frame_stream0"], d["single_stream_latency"])
PY
