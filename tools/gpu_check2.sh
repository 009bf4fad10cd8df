#!/bin/bash
# One gpurun call: GPU test suite, smoke, bench line with the extra records, ncu pass over every kernel.
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none -k regex:hdrnet_b200 -c 260 -f -o gpurun_out/r02_all_kernels python tools/prof_all_kernels.py > gpurun_out/r02_all_kernels.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r02_all_kernels.log
