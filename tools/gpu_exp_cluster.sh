#!/bin/bash
# experiment: solver cluster width / stream count at the bench's operating point
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency"
for v in "2 8" "4 8" "8 8" "1 8" "2 6" "2 10"; do
  set -- $v
  timeout 300 $B --cluster $1 --streams $2 > gpurun_out/exp_cl_$1_s$2.json 2> gpurun_out/exp_cl_$1_s$2.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/exp_cl_$1_s$2.json"))
    r = d["roofline"]; s = d.get("roofline_schur", {})
    print("cluster $1 streams $2: value", d["value"], "e2e", d["e2e"]["value"], "ba ms", d["value_stage_ms_per_frame_stream0"]["local_ba"], "chol us", r["avg_launch_us"], "frac", r["frac"], "schur us", s.get("avg_launch_us"), "schur frac", s.get("frac"))
except Exception as e:
    print("cluster $1 $2 failed", e)
PY
done
