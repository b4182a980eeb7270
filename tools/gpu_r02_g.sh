#!/bin/bash
# (historical: B200_SHORT_SEL_SMEM / B200_LONG_SEL_SMEM selected the selectors-in-shared-memory variants, removed after this measurement -- profiles/variants_r02.json)
# Round-2 GPU pass G (one GPU): per-row selectors in shared memory (more warps per scheduler) for K1 and K3.
set -u
out=gpurun_out
mkdir -p $out
for ssel in 0 1; do for pipe in 0 1; do
  B200_SHORT_SEL_SMEM=$ssel B200_FILL_PIPE=$pipe python bench.py --device-only --steps 10 --warmup 3 > $out/r02g_k1_sel${ssel}_pipe$pipe.json 2> $out/r02g_k1.err
  python -c "
import json; d=json.load(open('$out/r02g_k1_sel${ssel}_pipe$pipe.json')); print('K1 sel_smem=$ssel pipe=$pipe step', round(d['ms_per_step'],3), 'fill', round(d['roofline']['fill_ms_per_step'],3))"
done; done
for lsel in 0 1; do for lds in 0 2; do
  echo "K3 sel_smem=$lsel lds=$lds"
  B200_LONG_SEL_SMEM=$lsel B200_SUBST_LDS=$lds python tools/bench_long.py --pairs 2048 --type 2 --steps 3 | tee $out/r02g_k3_semi_sel${lsel}_lds$lds.json | python -c "import json,sys; d=json.load(sys.stdin); print(' semi fill', round(d['fill_ms'],2), 'step', round(d['ms_per_step'],2), d['parity_ok'])"
  B200_LONG_SEL_SMEM=$lsel B200_SUBST_LDS=$lds python tools/bench_long.py --pairs 512 --fixed 10000 --type 1 --steps 2 | tee $out/r02g_k3_local_sel${lsel}_lds$lds.json | python -c "import json,sys; d=json.load(sys.stdin); print(' local fill', round(d['fill_ms'],2), 'step', round(d['ms_per_step'],2), d['parity_ok'])"
done; done
python -m pytest tests/test_gpu_host_paths.py tests/test_gpu_align.py -q --timeout 600 > $out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> $out/r02g_pytest.log; tail -5 $out/r02g_pytest.log
