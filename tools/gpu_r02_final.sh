#!/bin/bash
# Round-2 final pass (one GPU): what the driver runs at round end -- every GPU test, smoke, both bench arms -- plus the CLI.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q --timeout 900 > $out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02f_pytest.log; tail -3 $out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference > $out/r02f_bench_reference.json 2> $out/r02f_bench_reference.err; cut -c1-160 $out/r02f_bench_reference.json
python bench.py > $out/r02f_bench.json 2> $out/r02f_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/r02f_bench.json')); e=d['e2e']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(e['value']), round(e['ms_per_step'],3), 'ptr', round(e['pointer_api']['ms_per_step'],2))
print('c4', d['extra']['c4_strong'].get('reads_per_s'), 'c5', d['extra']['c5_strong'].get('gcups'), d['extra']['c5_strong'].get('e2e_gcups'))"
python tools/bench_cli.py 100000 1 > $out/r02f_cli_n1.jsonl 2> $out/r02f_cli.err; python -c "
import json
for l in open('$out/r02f_cli_n1.jsonl'):
    d=json.loads(l); print(d['argv'], round(d['wall_s'],2), round(d['reads_per_s']), d['trace'][-1])"
