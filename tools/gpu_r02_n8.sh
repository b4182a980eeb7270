#!/bin/bash
# 8-GPU pass: bench.py under torchrun at N=8 exactly as the driver launches it, the CLI over 8 and 4 devices.
set -u
out=gpurun_out
mkdir -p $out
for N in 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > $out/r02n_bench_n$N.json 2> $out/r02n_bench_n$N.err; echo "bench N=$N rc=$?"
  python -c "
import json; d=json.load(open('$out/r02n_bench_n$N.json')); e=d['e2e']
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(e['value']), 'ms', round(e['ms_per_step'],2), 'h2d ceiling', round(e['h2d_ceiling_gbs'],1), 'frac', round(e['frac_of_h2d_bound'],3), 'ptr', round(e['pointer_api']['ms_per_step'],2))
c4=d['extra']['c4_strong']; c5=d['extra']['c5_strong']
print(' c4', c4.get('reads_per_s'), c4.get('map_s'), c4.get('rank_seconds_min_max'), c4.get('batch_reads'), c4.get('error'))
print(' c5', c5.get('gcups'), c5.get('s_per_pass'), c5.get('e2e_gcups'), c5.get('error'))"
done
for G in 8 4; do
python tools/bench_cli.py 100000 $G > $out/r02n_cli_n$G.jsonl 2> $out/r02n_cli_n$G.err; python -c "
import json
for l in open('$out/r02n_cli_n$G.jsonl'):
    d=json.loads(l); print($G, d['argv'], round(d['wall_s'],2), round(d['reads_per_s']), d['paf_lines'], d['trace'][-1])"
done
