#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_mapper.py -q --timeout 600 2>&1 | tail -2
for G in 1 2; do
python tools/bench_cli.py 100000 $G > $out/r02n2b_cli_n$G.jsonl 2> $out/r02n2b_cli.err; python -c "
import json
for l in open('$out/r02n2b_cli_n$G.jsonl'):
    d=json.loads(l); print($G, d['argv'], round(d['wall_s'],2), round(d['reads_per_s']), d['paf_lines'], d['paf_md5'], d['trace'][-1][:150])"
done
