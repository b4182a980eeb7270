#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
for sf in 0 1 2; do
  B200_STREAM_FILL=$sf python bench.py --no-extra --no-strong --no-cpu-baseline > $out/r02e_bench_sf$sf.json 2> $out/r02e_bench_sf$sf.err
  python -c "
import json; d=json.load(open('$out/r02e_bench_sf$sf.json')); print('stream_fill=$sf value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'ptr ms', round(d['e2e']['pointer_api']['ms_per_step'],3))"
done
python -m pytest tests/test_gpu_host_paths.py tests/test_gpu_mapper.py tests/test_abi.py "tests/test_gpu_fullsize.py::test_config4_paf_of_64_reads_matches_the_reference_mapper" -q --timeout 900 > $out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02e_pytest.log; tail -4 $out/r02e_pytest.log
python tools/bench_cli.py 100000 1 > $out/r02e_cli_n1.jsonl 2> $out/r02e_cli.err; cat $out/r02e_cli_n1.jsonl | cut -c1-900
