#!/bin/bash
# Round-2 GPU pass B (one GPU): ncu --set full captures of the fill kernels (both substitution variants), of the
# minimizer kernel on the judged read batch and of the mapper kernels. The reports are turned into CSV pages on the
# box (raw metrics + per-SASS-instruction page) and removed: gpurun brings back at most 64 MiB.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_minimize.py "tests/test_gpu_fullsize.py::test_config3_minimizer_tuples_of_2000_reads" -q > $out/r02b_pytest_min.log 2>&1; tail -2 $out/r02b_pytest_min.log
PROF_TIME=1 python tools/prof_mapper.py minimize > $out/r02b_min_time.log 2>&1; tail -1 $out/r02b_min_time.log
NCU="ncu --set full --clock-control none --import-source on"
export_rep() {   # $1 = report stem
  ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2>/dev/null
  ncu -i $out/$1.ncu-rep --page source --print-source sass --csv 2>/dev/null | gzip > $out/$1_sass.csv.gz
  rm -f $out/$1.ncu-rep
}
for lds in 0 1; do
  B200_SUBST_LDS=$lds $NCU -k regex:fill_short_kernel -c 1 -o $out/r02b_k1_lds$lds -f python bench.py --device-only --steps 1 --warmup 1 > $out/r02b_ncu_k1_lds$lds.log 2>&1
  export_rep r02b_k1_lds$lds
  B200_SUBST_LDS=$lds $NCU -k regex:fill_long16_kernel -c 1 -o $out/r02b_k3_lds$lds -f python tools/bench_long.py --pairs 2048 --type 2 --steps 1 --check 0 > $out/r02b_ncu_k3_lds$lds.log 2>&1
  export_rep r02b_k3_lds$lds
done
$NCU -k regex:minimize_kernel -c 1 --launch-skip 1 -o $out/r02b_minimize -f python tools/prof_mapper.py minimize > $out/r02b_ncu_min.log 2>&1
export_rep r02b_minimize
$NCU -k regex:'chain_kernel|seed_count_kernel|seed_emit_kernel|dedup_flag_kernel|dedup_scatter_kernel|region_kernel' --launch-skip 8 -c 8 -o $out/r02b_mapper -f python tools/prof_mapper.py map > $out/r02b_ncu_map.log 2>&1
ncu -i $out/r02b_mapper.ncu-rep --page raw --csv > $out/r02b_mapper_raw.csv 2>/dev/null; rm -f $out/r02b_mapper.ncu-rep
du -sh $out; ls -la $out | tail -20
