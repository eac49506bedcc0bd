#!/bin/bash
# Round-2 (second session) profile capture: the kernels added after scripts/capture_profiles.sh.
# Every command first runs WITHOUT ncu and must exit 0; numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
python scripts/conv_dw_bench.py > gpurun_out/r02b_plain_dw.txt 2>&1 || { echo "plain dw failed"; exit 1; }
python scripts/norm_bench.py > gpurun_out/r02b_plain_norm.txt 2>&1 || { echo "plain norm failed"; exit 1; }
python scripts/pruned_tier_profile.py > gpurun_out/r02b_plain_pruned.txt 2>&1 || { echo "plain pruned failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv1x1_dw_kernel -s 3 -c 1 -f -o gpurun_out/r02b_conv_dw \
    python scripts/conv_dw_bench.py > gpurun_out/r02b_ncu_dw.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:groupnorm_silu_bwd_reg_kernel|groupnorm_silu_fwd_kernel" -s 40 -c 2 -f -o gpurun_out/r02b_norm \
    python scripts/norm_bench.py > gpurun_out/r02b_ncu_norm.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:pruned_search_kernel|pruned_select_kernel" -s 2 -c 2 -f -o gpurun_out/r02b_pruned \
    python scripts/pruned_tier_profile.py > gpurun_out/r02b_ncu_pruned.log 2>&1
B="python bench.py --steps 2 --warmup 3 --no-strawman"
$B --workload c3 > gpurun_out/r02b_plain_c3.json 2> gpurun_out/r02b_plain_c3.err || { echo "plain c3 failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches_c3.csv \
    $B --workload c3 > gpurun_out/r02b_ncu_list_c3.log 2>&1
ls -la gpurun_out | grep r02b_
