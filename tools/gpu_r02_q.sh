#!/bin/bash
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_gpu_host_paths.py -q --timeout 900 -k "pointer" 2>&1 | tail -2
python tools/fuzz_gpu.py 100 21 2>&1 | tail -2
python tools/fuzz_gpu.py 60 22 2>&1 | tail -2
python tools/fuzz_gpu_minimize.py 30 5 2>&1 | tail -1
python tools/fuzz_gpu_mapper.py 40 6 2>&1 | tail -1
