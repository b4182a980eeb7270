#!/bin/bash
# Round-2 GPU pass P (one GPU): the pointer-array entry point with the host-side 2-bit packing gather.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_host_paths.py tests/test_gpu_align.py -q --timeout 900 > $out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02p_pytest.log; tail -4 $out/r02p_pytest.log
for hp in 1 0; do
  B200_HOST_PACK=$hp B200_TRACE=1 python bench.py --no-extra --no-strong --no-cpu-baseline --steps 5 > $out/r02p_bench_hp$hp.json 2> $out/r02p_bench_hp$hp.err
  python -c "
import json; d=json.load(open('$out/r02p_bench_hp$hp.json')); print('host_pack=$hp e2e ms', round(d['e2e']['ms_per_step'],3), 'ptr ms', round(d['e2e']['pointer_api']['ms_per_step'],2), 'x', round(d['e2e']['pointer_api']['vs_packed'],2))"
  grep "ptr:" $out/r02p_bench_hp$hp.err | tail -2 | cut -c1-400
done
