#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
env B200_STREAM_FILL=1 python tools/trace_e2e.py > $out/r02d_trace_stream1.log 2>&1
grep "stall\|timeline\|^step" $out/r02d_trace_stream1.log | cut -c1-2600 | tail -5
env B200_STREAM_FILL=0 python tools/trace_e2e.py > $out/r02d_trace_stream0.log 2>&1
grep "^step" $out/r02d_trace_stream0.log | tail -3
python -m pytest tests/test_gpu_host_paths.py tests/test_gpu_align.py "tests/test_gpu_fullsize.py::test_config2_every_pair_against_the_reference" -q --timeout 600 > $out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02d_pytest.log
tail -6 $out/r02d_pytest.log
python bench.py --no-extra --no-strong > $out/r02d_bench.json 2> $out/r02d_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/r02d_bench.json')); print(d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['pointer_api'])"
