#!/bin/bash
# (historical: B200_LONG_PIPE selected the software-pipelined K3 variant, removed after this measurement -- profiles/variants_r02.json)
# Round-2 GPU pass C (one GPU): streaming fill (gate kernels), pipelined K3, per-warp minimizer staging.
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_host_paths.py tests/test_gpu_minimize.py tests/test_gpu_mapper.py "tests/test_gpu_fullsize.py::test_config2_every_pair_against_the_reference" "tests/test_gpu_fullsize.py::test_config3_minimizer_tuples_of_2000_reads" -q --timeout 600 > $out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02c_pytest.log
tail -6 $out/r02c_pytest.log
PROF_TIME=1 python tools/prof_mapper.py minimize 2>&1 | tail -1 | tee $out/r02c_min_time.log
for lds in 0 2; do for pipe in 0 1; do
  echo "K3 lds=$lds pipe=$pipe"
  B200_SUBST_LDS=$lds B200_LONG_PIPE=$pipe python tools/bench_long.py --pairs 2048 --type 2 --steps 3 | tee $out/r02c_k3_semi_lds${lds}_pipe$pipe.json | python -c "import json,sys; d=json.load(sys.stdin); print(' semi', round(d['fill_ms'],2), round(d['ms_per_step'],2), d['parity_ok'])"
  B200_SUBST_LDS=$lds B200_LONG_PIPE=$pipe python tools/bench_long.py --pairs 512 --fixed 10000 --type 1 --steps 2 | tee $out/r02c_k3_local_lds${lds}_pipe$pipe.json | python -c "import json,sys; d=json.load(sys.stdin); print(' local', round(d['fill_ms'],2), round(d['ms_per_step'],2), d['parity_ok'])"
done; done
for sf in 1 0; do
  B200_STREAM_FILL=$sf python tools/trace_e2e.py > $out/r02c_trace_stream$sf.log 2>&1
  grep "^step" $out/r02c_trace_stream$sf.log | tail -3
done
B200_STREAM_FILL=1 B200_TAPER_TAIL=0 python tools/trace_e2e.py > $out/r02c_trace_stream1_notaper.log 2>&1; grep "^step" $out/r02c_trace_stream1_notaper.log | tail -2
python bench.py --no-extra --no-strong > $out/r02c_bench.json 2> $out/r02c_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$out/r02c_bench.json')); print(d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['pointer_api'])"
