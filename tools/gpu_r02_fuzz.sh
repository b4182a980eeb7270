#!/bin/bash
set -u
python tools/fuzz_gpu.py 150 101 2>&1 | tail -1
python tools/fuzz_gpu.py 150 102 2>&1 | tail -1
python tools/fuzz_gpu_minimize.py 60 103 2>&1 | tail -1
python tools/fuzz_gpu_mapper.py 90 104 2>&1 | tail -1
