#!/bin/bash
# Runs on the GPU box (under gpurun): per-launch time list + one full ncu capture of one step.
# usage: tools/profile.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
K='regex:preprocess_kernel|tile_scan_kernel|scatter_kernel|render_fwd_kernel|render_bwd_kernel|preprocess_bwd_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 24 -c 30 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ops > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch-list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 16 -c 4 -f \
    -o $OUT/fwd_step_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ops > $OUT/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
ls -la $OUT
