#!/bin/bash
# Round-2 evidence: launch lists (C2 full, C3 400 frames) and ncu --set full captures of the hot kernels on the C2 shape.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
bash scripts/gpu_launches.sh 1000 r2_launches_c2 c2
bash scripts/gpu_launches.sh 400 r2_launches_c3 c3
WL=c2 bash scripts/gpu_prof.sh 1000 score_walk r2_score_walk ransac_count r2_ransac_count eight_point_qr r2_eight_point_qr eight_point_list r2_eight_point_list eight_point_warp r2_eight_point_warp bucket_select r2_bucket_select klt_quad r2_klt_quad klt_lane r2_klt_border pyr_down r2_pyr pose_kernel r2_pose
