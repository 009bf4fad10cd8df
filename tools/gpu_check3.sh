#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
bash tools/ncu_all.sh
