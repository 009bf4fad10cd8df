#!/bin/bash
# One gpurun call for the issuer-warp kernel: its parity tests, the interleaved A/B, one full ncu capture.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.csv 2>&1
timeout 600 python -m pytest tests/test_slice_apply_gpu.py -q --timeout 120 -p no:cacheprovider \
    -k "async or bitwise or 4k_frame or oracle" -x > gpurun_out/pytest_async.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_async.log
tail -8 gpurun_out/pytest_async.log
timeout 300 python tools/ab_bench.py "tex t512" async > gpurun_out/ab_stdout.txt 2>&1; echo "ab exit $?"; cat gpurun_out/ab_stdout.txt | tail -12
HDRNET_TEX_CHUNKS=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:slice_apply_rows_async -s 2 -c 1 \
    -f -o gpurun_out/prof_async python tools/prof_variant.py 7 > gpurun_out/ncu_async.log 2>&1
echo "ncu exit $?"
