#!/bin/bash
# One gpurun call: GPU test suite, smoke, bench line, same-box A/B of the issuer-warp CTA shapes.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
L=hdrnet_b200/lib/libhdrnet_b200.so
AB_ROUNDS=5 timeout 300 python tools/ab_lib.py $L:7:HDRNET_ASYNC_THREADS=512 $L:7:HDRNET_ASYNC_THREADS=352 $L:7:HDRNET_ASYNC_THREADS=352,HDRNET_TEX_CHUNKS=4 $L:4 $L:2 > gpurun_out/ab_lib.txt 2>&1; grep -v bursts gpurun_out/ab_lib.txt | tail -8
