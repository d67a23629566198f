#!/bin/bash
# usage (on the GPU box, through gpurun): bash tools/gpu_round.sh <tag> [full]
# One call that brings back everything a round needs: GPU parity tests, the bench line, the ncu launch list of the bench
# command, and (with "full") --set full captures of the heavy kernels.  Every leg has its own timeout.
tag=$1; full=$2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log
tail -c 600 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 400 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("stage_ms_per_frame_stream0"))
    print(d.get("value_stage_ms_per_frame_stream0")); print(d.get("stage_us_per_frame")); print(d.get("single_stream_latency"), d.get("gpu_launches")); print(d.get("roofline")); print(d.get("cpu_baseline")); print(d.get("clocks"))
except Exception as e:
    print("bench parse failed", e)
PY
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/${tag}_launches_bench_steps2_warmup1.csv \
    python bench.py --steps 2 --warmup 1 --streams 1 --frames-per-step 1 --no-cpu-baseline --no-latency > gpurun_out/${tag}_launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches_bench_steps2_warmup1.csv | head -40
if [ "$full" = "full" ]; then
  for k in k_ba_cholesky_solve k_ba_schur_chunk k_ba_linearize k_ba_update; do
    timeout 240 $NCU --set full --import-source on -k regex:$k --launch-skip 3 -c 1 -o gpurun_out/${tag}_full_$k -f python tools/profile_step.py ba 1 > /dev/null 2>&1
    ncu -i gpurun_out/${tag}_full_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_full_$k.csv 2>/dev/null
  done
  for k in k_hamming_topk k_fast_score k_cell_nms k_orient_describe k_pyramid_group k_tree_distribute; do
    timeout 240 $NCU --set full --import-source on -k regex:$k --launch-skip 1 -c 1 -o gpurun_out/${tag}_full_$k -f python tools/profile_step.py match 3 > /dev/null 2>&1
    ncu -i gpurun_out/${tag}_full_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_full_$k.csv 2>/dev/null
  done
  for k in k_ba_linearize k_ba_update k_hamming_topk k_cell_nms k_orient_describe k_pyramid_group; do rm -f gpurun_out/${tag}_full_$k.ncu-rep; done
  timeout 120 python tools/profile_step.py ba 2 2>&1 | tail -22 > gpurun_out/${tag}_cholesky_phase_clocks.txt
  python tools/ncu_extract.py --json gpurun_out/${tag}_dram_traffic.json gpurun_out/${tag}_full_*.csv > gpurun_out/${tag}_ncu_full_summary.md 2>&1
  grep -E "^###|duration|dram" gpurun_out/${tag}_ncu_full_summary.md | head -60
fi
