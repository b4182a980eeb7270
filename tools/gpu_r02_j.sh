#!/bin/bash
# Round-2 GPU pass J (one GPU): every GPU test with the final defaults, the final N=1 bench line, phase trace of the pointer-array path.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q --timeout 900 > $out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02j_pytest.log; tail -3 $out/r02j_pytest.log
python bench.py > $out/r02j_bench.json 2> $out/r02j_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/r02j_bench.json')); e=d['e2e']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(e['value']), round(e['ms_per_step'],3), 'ptr', round(e['pointer_api']['ms_per_step'],2), 'traffic src', d['roofline']['traffic_source'])
print('c4', d['extra']['c4_strong'].get('reads_per_s'), 'c5', d['extra']['c5_strong'].get('gcups'))"
B200_TRACE=1 python bench.py --no-extra --no-strong --no-cpu-baseline --steps 3 > /dev/null 2> $out/r02j_trace.err; grep "ptr:" $out/r02j_trace.err | tail -3 | cut -c1-600
