#!/bin/bash
# Round-2 profile capture (run on the GPU box through gpurun; outputs under gpurun_out/).
# Every command first runs WITHOUT ncu and must exit 0; numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-strawman"
for wl in c2 c3; do
  $B --workload $wl > gpurun_out/r02_plain_$wl.json 2> gpurun_out/r02_plain_$wl.err || { echo "plain $wl failed"; tail -5 gpurun_out/r02_plain_$wl.err; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_$wl.csv \
      $B --workload $wl > gpurun_out/r02_ncu_list_$wl.log 2>&1
done
# full captures of the dominant / changed kernels (one launch each, after the warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:search_dual_kernel -s 3 -c 1 -f -o gpurun_out/r02_dual_c2 \
    $B --workload c2 > gpurun_out/r02_ncu_full_dual.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_tc16_kernel -s 3 -c 1 -f -o gpurun_out/r02_tc16_c3 \
    $B --workload c3 > gpurun_out/r02_ncu_full_tc16.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:gather_loss_st_pipe_kernel|backward_pipe_kernel|rescore_groups_kernel|split16_tokens_kernel" -s 8 -c 4 -f \
    -o gpurun_out/r02_helpers_c3 $B --workload c3 > gpurun_out/r02_ncu_full_helpers.log 2>&1
ls -la gpurun_out | grep r02_
