#!/bin/bash
# Round-2 GPU pass B (one GPU): ncu --set full captures of the fill kernels (both substitution variants), of the
# minimizer kernel on the judged read batch and of the mapper kernels.
set -u
out=gpurun_out
mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
for lds in 0 1; do
  B200_SUBST_LDS=$lds $NCU -k regex:fill_short_kernel -c 2 -o $out/r02b_k1_lds$lds -f python bench.py --device-only --steps 1 --warmup 1 > $out/r02b_ncu_k1_lds$lds.log 2>&1
  B200_SUBST_LDS=$lds $NCU -k regex:fill_long16_kernel -c 2 -o $out/r02b_k3_lds$lds -f python tools/bench_long.py --pairs 2048 --type 2 --steps 1 --check 0 > $out/r02b_ncu_k3_lds$lds.log 2>&1
done
$NCU -k regex:minimize_kernel -c 3 -o $out/r02b_minimize -f python tools/prof_mapper.py minimize > $out/r02b_ncu_min.log 2>&1
$NCU -k regex:'chain_kernel|seed_count_kernel|seed_emit_kernel|dedup_flag_kernel|dedup_scatter_kernel|region_kernel' --launch-skip 12 -c 14 -o $out/r02b_mapper -f python tools/prof_mapper.py map > $out/r02b_ncu_map.log 2>&1
ls -la $out/*.ncu-rep
