#!/bin/bash
# ncu captures of the cost-volume, PTF and raster-backward kernels (one launch each).  usage: tools/profile_ops.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cost_volume_fwd_tc -s 3 -c 1 -f -o $OUT/cv_fwd_$TAG python tools/bench_cv.py > $OUT/ncu_cv_$TAG.log 2>&1; echo "cv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:ptf_compact_kernel|ptf_match_kernel|ptf_project_kernel" -s 9 -c 3 -f -o $OUT/ptf_$TAG python tools/bench_ptf_only.py > $OUT/ncu_ptf_$TAG.log 2>&1; echo "ptf rc=$?"
ls -la $OUT | tail -8
