#!/bin/bash
# Profiling pass of a round (run under gpurun, one GPU): the ncu launch list of the bench command and one
# `--set full` capture of the dominant kernels. Outputs land in gpurun_out/; tools/ncu_summarize.py turns the
# .ncu-rep into the JSON summaries committed under profiles/.
set -u
tag=${1:-r01b}
out=gpurun_out
mkdir -p $out
python bench.py --steps 2 --warmup 3 > $out/bench_${tag}_plain.json 2> $out/bench_${tag}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_${tag}_bench.csv \
    python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > $out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fill_short_kernel|walk_kernel' -c 4 \
    -o $out/prof_${tag}_short -f python bench.py --steps 1 --warmup 1 --cpu-seconds 1 > $out/ncu_full_${tag}_short.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'fill_long16_kernel|walk_tile' -c 8 \
    -o $out/prof_${tag}_long -f python bench.py --steps 1 --warmup 1 --cpu-seconds 1 > $out/ncu_full_${tag}_long.log 2>&1
ls -la $out | tail -8
