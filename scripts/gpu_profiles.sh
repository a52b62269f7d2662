#!/bin/bash
# Round profile set on the C2 shape: ncu launch list + one --set full capture per hot kernel (gpurun_out/, scratch).
mkdir -p gpurun_out
bash scripts/gpu_launches.sh 999 r1_launches > /dev/null
NCU_COUNT=1 bash scripts/gpu_prof.sh 999 klt_quad_kernel r1_klt_quad radix_sort_frame r1_radix nms_kernel r1_nms pyr_down r1_pyr ransac_count r1_ransac
NCU_COUNT=2 bash scripts/gpu_prof.sh 999 score_tile_kernel r1_score 'klt_lane_kernel' r1_klt_border
