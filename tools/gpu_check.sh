#!/bin/bash
# usage (on the GPU box, through gpurun): bash tools/gpu_check.sh <tag> [pytest args]
tag=$1; shift
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu "$@" 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 1800 gpurun_out/${tag}_pytest.log
tail -c 600 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("stage_ms_per_frame_stream0"))
    print(d.get("value_stage_ms_per_frame_stream0")); print(d["stage_us_per_frame"]); print(d["single_stream_latency"], d["gpu_launches"]); print(d["roofline"])
except Exception as e:
    print("bench parse failed", e)
PY
