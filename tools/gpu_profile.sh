#!/bin/bash
# usage (on the GPU box): bash tools/gpu_profile.sh <tag>   -- ncu launch lists + --set full captures of the heavy kernels
tag=$1
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/${tag}_launches_ba.csv python tools/profile_step.py ba 2 > gpurun_out/${tag}_prof_ba.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/${tag}_launches_front.csv python tools/profile_step.py match 2 > gpurun_out/${tag}_prof_front.log 2>&1
for k in k_ba_schur_chunk k_ba_linearize k_ba_update k_ba_errors k_ba_cholesky_solve k_ba_pose_accum_chunk; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 3 -c 1 -o gpurun_out/${tag}_full_$k -f python tools/profile_step.py ba 1 > /dev/null 2>&1
  ncu -i gpurun_out/${tag}_full_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_full_$k.csv 2>/dev/null
done
for k in k_hamming_topk k_fast_score k_cell_nms k_orient_describe k_pyramid_group k_tree_distribute; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 1 -c 1 -o gpurun_out/${tag}_full_$k -f python tools/profile_step.py match 1 > /dev/null 2>&1
  ncu -i gpurun_out/${tag}_full_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_full_$k.csv 2>/dev/null
done
python tools/summarize_launches.py gpurun_out/${tag}_launches_ba.csv | head -40
python tools/profile_step.py ba 2 2>&1 | tail -22
