#!/bin/bash
# The measurement set behind profiles/<tag>_*: bench line, reference arm, ncu launch lists of the bench command and of
# one LM solve, one --set full capture of every hot kernel (summarised by scripts/ncu_summary.py).
#   bash scripts/final_measurements.sh <tag> [part]     part 1: bench + launch lists, 2: full capture, 3: 16-camera capture
T=${1:-r02_z}; P=${2:-1}
if [ "$P" = 1 ]; then
  python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
  python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_steps6.csv python bench.py --steps 6 --warmup 3 > /dev/null 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_lm_launches.csv python scripts/lm_launches.py > gpurun_out/${T}_lm.log 2>&1
elif [ "$P" = 2 ]; then
  ncu --set full --clock-control none --import-source on -k regex:'k2p_kernel|k2c_|k2_syrk|finalize_kernel|solve_reduced|backsub_kernel|sum_scalars|residual_chunks|triangulate_kernel|project_points_multi|homography_transfer' -c 48 -o gpurun_out/${T}_full python scripts/ncu_targets.py > gpurun_out/${T}_ncu.log 2>&1
else
  ncu --set full --clock-control none -k regex:'k2c_|k2_syrk|k2p_kernel|solve_reduced' -c 8 -o gpurun_out/${T}_full_16cam python scripts/ncu_targets.py 16 25000 > gpurun_out/${T}_ncu16.log 2>&1
fi
ls -la gpurun_out | tail -8
