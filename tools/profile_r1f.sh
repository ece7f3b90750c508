#!/bin/bash
# ncu captures of the kernels that had no committed capture yet: cost-volume backward (tcgen05), PTF GRU (tcgen05),
# render backward, depth head.  usage: tools/profile_r1f.sh <tag>
TAG=${1:-r1f}
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
CV_N=1 timeout 300 $NCU -k regex:cost_volume_bwd_tc -s 1 -c 1 -o $OUT/cv_bwd_$TAG python tools/bench_cv_bwd.py > $OUT/ncu_cvb_$TAG.log 2>&1; echo "cv bwd rc=$?"
timeout 300 $NCU -k regex:cost_volume_fwd_tc -s 3 -c 1 -o $OUT/cv_fwd_$TAG python tools/bench_cv.py > $OUT/ncu_cvf_$TAG.log 2>&1; echo "cv fwd rc=$?"
timeout 300 $NCU -k regex:ptf_gru_tc_kernel -s 2 -c 1 -o $OUT/ptf_gru_$TAG python tools/bench_ptf_only.py > $OUT/ncu_gru_$TAG.log 2>&1; echo "gru rc=$?"
timeout 300 $NCU -k "regex:render_bwd_kernel|preprocess_bwd_kernel" -s 4 -c 2 -o $OUT/raster_bwd_$TAG python tools/bench_raster_bwd.py > $OUT/ncu_rb_$TAG.log 2>&1; echo "raster bwd rc=$?"
timeout 300 $NCU -k regex:depth_head -s 2 -c 1 -o $OUT/dh_$TAG python tools/bench_depth_head.py 2 > $OUT/ncu_dh_$TAG.log 2>&1; echo "dh rc=$?"
ls -la $OUT | tail -12
