#!/bin/bash
# ncu captures of the cost-volume, PTF and raster-backward kernels (one launch each).  usage: tools/profile_ops.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cost_volume_fwd -s 3 -c 1 -f -o $OUT/cv_fwd_$TAG python tools/bench_cv.py > $OUT/ncu_cv_$TAG.log 2>&1; echo "cv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cost_volume_bwd -s 0 -c 1 -f -o $OUT/cv_bwd_$TAG python tools/bench_ops.py > $OUT/ncu_cvb_$TAG.log 2>&1; echo "cvb rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:render_bwd_kernel|preprocess_bwd_kernel" -s 6 -c 2 -f -o $OUT/raster_bwd_$TAG python tools/bench_ops.py > $OUT/ncu_rb_$TAG.log 2>&1; echo "rb rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:ptf_" -c 60 --csv --log-file $OUT/launches_ptf_$TAG.csv python tools/bench_ops.py > $OUT/ncu_ptf_$TAG.log 2>&1; echo "ptf rc=$?"
ls -la $OUT | tail -12
