#!/bin/bash
# experiment: fused trial / linearisation pass + single-barrier cluster reduction in k_pose_optimize
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_optimize_gpu.py tests/test_class_layer.py -x -q -m gpu 2>&1 | tail -5
ncu --clock-control none --metrics gpu__time_duration.sum -c 40 --csv --log-file gpurun_out/exp_pose_launches.csv python tools/profile_step.py pose 4 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/exp_pose_launches.csv 2>/dev/null | head -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_pose_bench.json 2> gpurun_out/exp_pose_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/exp_pose_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["value_stage_ms_per_frame_stream0"], d["single_stream_latency"], d["stage_us_per_frame"]["pose_optimizer_kernel"])
PY
