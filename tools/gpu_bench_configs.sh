#!/bin/bash
# bench lines of the other BASELINE configurations (configs[1], configs[2], configs[4])
mkdir -p gpurun_out
for c in 2 3 5; do
  timeout 200 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/r2_bench_config$c.json 2> gpurun_out/r2_bench_config$c.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_config$c.json"))
    print("config $c:", d["config"]["workload"][:60], "| value", d["value"], d["unit"], "e2e", d["e2e"]["value"], "streams", d["config"]["streams_per_gpu"], d["value_stage_ms_per_frame_stream0"])
except Exception as e:
    print("config $c failed", e); print(open("gpurun_out/r2_bench_config$c.err").read()[-600:])
PY
done
