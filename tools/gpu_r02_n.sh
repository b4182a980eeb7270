#!/bin/bash
# multi-GPU pass: bench.py under torchrun at N ranks (as the driver launches it) + the CLI over N devices
set -u
N=${1:-2}
out=gpurun_out
mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $out/r02n_bench_n$N.json 2> $out/r02n_bench_n$N.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/r02n_bench_n$N.json')); e=d['e2e']
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(e['value']), 'ms', round(e['ms_per_step'],2), 'h2d ceiling', round(e['h2d_ceiling_gbs'],1), 'frac', round(e['frac_of_h2d_bound'],3), 'ptr', round(e['pointer_api']['ms_per_step'],2))
print('c4', d['extra']['c4_strong']); print('c5', d['extra']['c5_strong'])"
tail -3 $out/r02n_bench_n$N.err
python tools/bench_cli.py 100000 $N > $out/r02n_cli_n$N.jsonl 2> $out/r02n_cli_n$N.err; cut -c1-400 $out/r02n_cli_n$N.jsonl
