#!/bin/bash
# Round-2 GPU pass L (one GPU): the final state -- every GPU test, smoke, the N=1 bench line, the reference arm, the CLI.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q --timeout 900 > $out/r02l_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02l_pytest.log; tail -3 $out/r02l_pytest.log
python __graft_entry__.py smoke > $out/r02l_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > $out/r02l_bench.json 2> $out/r02l_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/r02l_bench.json')); e=d['e2e']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(e['value']), round(e['ms_per_step'],3), 'ptr', round(e['pointer_api']['ms_per_step'],2), 'traffic src', d['roofline']['traffic_source'])
print('c4', d['extra']['c4_strong'].get('reads_per_s'), 'c5', d['extra']['c5_strong'].get('gcups'), d['extra']['c5_strong'].get('e2e_gcups'))"
python bench.py --impl reference --steps 2 --warmup 1 > $out/r02l_bench_reference.json 2>> $out/r02l_bench.err; cut -c1-200 $out/r02l_bench_reference.json
for rep in 1 2; do python tools/bench_cli.py 100000 1 > $out/r02l_cli_n1_$rep.jsonl 2> $out/r02l_cli.err; python -c "
import json
for l in open('$out/r02l_cli_n1_$rep.jsonl'):
    d=json.loads(l); print(d['argv'], round(d['wall_s'],2), round(d['reads_per_s']), d['trace'][-1])"; done
