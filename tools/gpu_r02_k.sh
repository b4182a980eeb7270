#!/bin/bash
# Round-2 GPU pass K (one GPU): K1 local with the packed first-row search, CLI with huge-page buffers.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_align.py tests/test_gpu_host_paths.py tests/test_gpu_mapper.py "tests/test_gpu_fullsize.py::test_config2_every_pair_against_the_reference" -q --timeout 900 > $out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02k_pytest.log; tail -3 $out/r02k_pytest.log
python tools/bench_k1_types.py > $out/r02k_k1_types.json 2> $out/r02k_k1_types.err; cat $out/r02k_k1_types.json | cut -c1-600
python tools/bench_cli.py 100000 1 > $out/r02k_cli_n1.jsonl 2> $out/r02k_cli.err; python -c "
import json
for l in open('$out/r02k_cli_n1.jsonl'):
    d=json.loads(l); print(d['argv'], round(d['wall_s'],2), round(d['reads_per_s']), d['trace'][-1])"
