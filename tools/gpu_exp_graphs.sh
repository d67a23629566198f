#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency"
for v in "" "--graphs" "--graphs --streams 4" "--streams 4"; do
  timeout 300 $B $v > gpurun_out/exp_graphs.json 2> gpurun_out/exp_graphs.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/exp_graphs.json"))
    print("[$v] value", d["value"], "e2e", d["e2e"]["value"], "ba ms", d["value_stage_ms_per_frame_stream0"]["local_ba"])
except Exception as e:
    print("[$v] failed", e)
PY
done
