#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
python __graft_entry__.py smoke 2>&1 | tail -1
for t in 0 2 0 2; do
  B200_TAPER_TAIL=$t python bench.py --no-extra --no-strong --no-cpu-baseline > $out/r02m_bench_t$t.json 2> $out/r02m_bench.err
  python -c "
import json; d=json.load(open('$out/r02m_bench_t$t.json')); print('taper_tail=$t e2e ms', round(d['e2e']['ms_per_step'],3), 'ptr', round(d['e2e']['pointer_api']['ms_per_step'],2))"
done
