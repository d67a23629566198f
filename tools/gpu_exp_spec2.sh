#!/bin/bash
# experiment: static second trial batch (ovs_optimizer_set_second_batch) vs the 4-wide single batch, same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_optimize_gpu.py -x -q -m gpu 2>&1 | tail -5
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency"
for v in "4 0 8" "2 2 8" "1 3 8" "2 0 8" "2 2 12" "3 1 8"; do
  set -- $v
  timeout 300 $B --spec $1 --spec2 $2 --streams $3 > gpurun_out/exp_spec_$1_$2_s$3.json 2> gpurun_out/exp_spec_$1_$2_s$3.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/exp_spec_$1_$2_s$3.json"))
    print("spec $1 spec2 $2 streams $3: value", d["value"], "e2e", d["e2e"]["value"], "ba ms", d["value_stage_ms_per_frame_stream0"]["local_ba"], "trials", d["roofline"].get("lm_trials_per_frame"), "systems/launch", d["roofline"].get("systems_per_launch"), "launches/frame", d["roofline"].get("launches_per_frame"))
except Exception as e:
    print("spec $1 $2 $3 failed", e)
PY
done
NCU="ncu --clock-control none"
timeout 240 $NCU --set full --import-source on -k regex:k_hamming_topk --launch-skip 1 -c 1 -o gpurun_out/r2_full_k_hamming_topk -f python tools/profile_step.py match 3 > /dev/null 2>&1
ncu -i gpurun_out/r2_full_k_hamming_topk.ncu-rep --page raw --csv > gpurun_out/r2_full_k_hamming_topk.csv 2>/dev/null
rm -f gpurun_out/r2_full_k_hamming_topk.ncu-rep
python tools/ncu_extract.py gpurun_out/r2_full_k_hamming_topk.csv | head -12
