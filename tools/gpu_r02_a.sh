#!/bin/bash
# Round-2 GPU pass A (one GPU): every GPU test, kernel-variant A/B, the bench line, the e2e device timeline.
set -u
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/r02a_gpu.txt; nproc >> $out/r02a_gpu.txt
python -m pytest tests -m gpu -q --timeout 600 > $out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02a_pytest.log
tail -5 $out/r02a_pytest.log
for lds in 1 0; do
  B200_SUBST_LDS=$lds python bench.py --device-only --steps 10 --warmup 3 > $out/r02a_k1_lds$lds.json 2> $out/r02a_k1_lds$lds.err
  B200_SUBST_LDS=$lds python tools/bench_long.py --pairs 2048 --type 2 --steps 3 > $out/r02a_k3_semi_lds$lds.json 2>> $out/r02a_k1_lds$lds.err
  B200_SUBST_LDS=$lds python tools/bench_long.py --pairs 512 --fixed 10000 --type 1 --steps 2 > $out/r02a_k3_local_lds$lds.json 2>> $out/r02a_k1_lds$lds.err
done
python tools/trace_e2e.py > $out/r02a_trace.log 2>&1
B200_TAPER_TAIL=0 python tools/trace_e2e.py > $out/r02a_trace_notaper.log 2>&1
python bench.py > $out/r02a_bench.json 2> $out/r02a_bench.err; echo "bench rc=$?"
tail -c 600 $out/r02a_bench.json
