#!/bin/bash
# Round-2 measurement set on one B200 (run under gpurun): bench lines of configs 2 / 3 / 5, the ncu launch list of the
# config-2 step and `--set full` captures of the kernels named in profiles/r2_summary.md.  Outputs: gpurun_out/r2_final_*.
set -u
O=gpurun_out
python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err
python bench.py --config 3 > $O/r2_final_bench_c3.json 2> $O/r2_final_bench_c3.err
python bench.py --config 5 --steps 30 > $O/r2_final_bench_c5.json 2> $O/r2_final_bench_c5.err
python bench.py --impl reference --steps 10 --warmup 1 > $O/r2_final_bench_reference.json 2> $O/r2_final_bench_reference.err
STEP="python bench.py --steps 6 --warmup 3 --no-cpu --no-aux --no-graph --streams 1"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/r2_final_launches.csv $STEP > /dev/null 2>&1
for k in roi_tile_tma_kernel head_tc_kernel decode_select_cluster_kernel decode_collect_kernel nms_scan_kernel nms_mask_kernel roi_prep_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $O/r2_prof_$k $STEP > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:tail_conv_collect -s 2 -c 1 -f -o $O/r2_prof_tail_conv_collect python tools/tail_probe.py > /dev/null 2>&1
C3="python bench.py --config 3 --steps 3 --warmup 3 --no-cpu"
for k in focal_fwd_bwd_kernel focal_render_fwd_bwd_kernel render_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $O/r2_prof_$k $C3 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:roi_tile_bwd -s 1 -c 1 -f -o $O/r2_prof_roi_tile_bwd python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la $O | grep r2_ | tail -30
