#!/bin/bash
# Round-2 GPU pass I (one GPU): host-pipeline schedule A/B, CLI after the write(2) change, gzip test, final N=1 line.
set -u
out=gpurun_out
mkdir -p $out
for taper in 0 1; do for sf in 0 1; do
  B200_TAPER_TAIL=$taper B200_STREAM_FILL=$sf python bench.py --no-extra --no-strong --no-cpu-baseline > $out/r02i_bench_t${taper}_s$sf.json 2> $out/r02i_bench.err
  python -c "
import json; d=json.load(open('$out/r02i_bench_t${taper}_s$sf.json')); print('taper=$taper stream=$sf e2e ms', round(d['e2e']['ms_per_step'],3), 'ptr', round(d['e2e']['pointer_api']['ms_per_step'],2))"
done; done
python -m pytest tests/test_gpu_mapper.py -q --timeout 600 2>&1 | tail -2
python tools/bench_cli.py 100000 1 > $out/r02i_cli_n1.jsonl 2> $out/r02i_cli.err; python -c "
import json
for l in open('$out/r02i_cli_n1.jsonl'):
    d=json.loads(l); print(d['argv'], round(d['wall_s'],2), round(d['reads_per_s']), d['trace'][-1])"
