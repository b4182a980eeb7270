#!/bin/bash
# Round-2 GPU pass H (one GPU): every GPU test, the bench line, launch list + ncu captures of the shipping kernels, CLI.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q --timeout 900 > $out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02h_pytest.log; tail -4 $out/r02h_pytest.log
PROF_TIME=1 python tools/prof_mapper.py minimize 2>&1 | tail -1 | tee $out/r02h_min_time.log
python bench.py > $out/r02h_bench.json 2> $out/r02h_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/r02h_bench.json')); e=d['e2e']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(e['value']), round(e['ms_per_step'],3), 'ptr', round(e['pointer_api']['ms_per_step'],2), 'fill', round(d['roofline']['fill_ms_per_step'],3))
print('c4', d['extra']['c4_strong'].get('reads_per_s'), 'c5', d['extra']['c5_strong'].get('gcups'))
print({k:(round(v.get('fill_gcups',0)) if 'fill_gcups' in v else round(v.get('roofline_frac_hbm',0),3)) for k,v in d['extra']['single_gpu'].items()})"
python bench.py --impl reference --steps 2 --warmup 1 > $out/r02h_bench_reference.json 2>> $out/r02h_bench.err; cut -c1-300 $out/r02h_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/launches_r02_bench.csv python bench.py --steps 2 --warmup 3 --no-strong --no-extra --cpu-seconds 1 > $out/r02h_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
export_rep() { ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2>/dev/null; ncu -i $out/$1.ncu-rep --page source --print-source sass --csv 2>/dev/null | gzip > $out/$1_sass.csv.gz; rm -f $out/$1.ncu-rep; }
$NCU -k regex:'fill_short_kernel|walk_kernel|pack_kernel|emit_kernel' -c 4 -o $out/r02h_k1 -f python bench.py --device-only --steps 1 --warmup 1 > $out/r02h_ncu_k1.log 2>&1; export_rep r02h_k1
$NCU -k regex:'fill_long16_kernel|walk_tile_wait_kernel' -c 2 -o $out/r02h_k3 -f python tools/bench_long.py --pairs 2048 --type 2 --steps 1 --check 0 > $out/r02h_ncu_k3.log 2>&1; export_rep r02h_k3
$NCU -k regex:minimize_kernel -c 1 --launch-skip 1 -o $out/r02h_minimize -f python tools/prof_mapper.py minimize > $out/r02h_ncu_min.log 2>&1; export_rep r02h_minimize
$NCU -k regex:'chain_kernel|seed_count_kernel|seed_emit_kernel|dedup_flag_kernel|dedup_sentinel_kernel|dedup_scatter_kernel|region_kernel' --launch-skip 9 -c 9 -o $out/r02h_mapper -f python tools/prof_mapper.py map > $out/r02h_ncu_map.log 2>&1
ncu -i $out/r02h_mapper.ncu-rep --page raw --csv > $out/r02h_mapper_raw.csv 2>/dev/null; rm -f $out/r02h_mapper.ncu-rep
python tools/bench_cli.py 100000 1 > $out/r02h_cli_n1.jsonl 2> $out/r02h_cli.err; cut -c1-330 $out/r02h_cli_n1.jsonl
du -sh $out
