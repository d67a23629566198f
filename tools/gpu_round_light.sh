#!/bin/bash
# usage: bash tools/gpu_round_light.sh <tag>  -- the bench line, the ncu launch list of the bench command and a --set full capture of the
# Schur kernel (for a build whose GPU tests have already been run)
tag=$1
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 300 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["single_stream_latency"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d["roofline_schur"]["avg_launch_us"], d["roofline_schur"]["frac"], d["clocks"])
PY
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/${tag}_launches_bench_steps2_warmup1.csv \
    python bench.py --steps 2 --warmup 1 --streams 1 --frames-per-step 1 --no-cpu-baseline --no-latency > gpurun_out/${tag}_launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/${tag}_launches_bench_steps2_warmup1.csv | head -8
timeout 240 $NCU --set full --import-source on -k regex:k_ba_schur_chunk --launch-skip 3 -c 1 -o gpurun_out/${tag}_full_k_ba_schur_chunk -f python tools/profile_step.py ba 1 > /dev/null 2>&1
ncu -i gpurun_out/${tag}_full_k_ba_schur_chunk.ncu-rep --page raw --csv > gpurun_out/${tag}_full_k_ba_schur_chunk.csv 2>/dev/null
rm -f gpurun_out/${tag}_full_k_ba_schur_chunk.ncu-rep
python tools/ncu_extract.py gpurun_out/${tag}_full_k_ba_schur_chunk.csv | head -16
