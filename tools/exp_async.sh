#!/bin/bash
# One gpurun call: full GPU test suite + smoke + bench (tools/gpu_check.sh quick), then the same-box A/B of the
# programmatic-dependent-launch switch.
set -u
bash tools/gpu_check.sh quick
L=hdrnet_b200/lib/libhdrnet_b200.so
AB_ROUNDS=7 timeout 300 python tools/ab_lib.py $L:4:HDRNET_TEX_CHUNKS=4 $L:7:HDRNET_ASYNC_PDL=1 $L:7:HDRNET_ASYNC_PDL=0 $L:0 \
  $L:7:HDRNET_ASYNC_PDL=1,HDRNET_TEX_CHUNKS=4 $L:7:HDRNET_ASYNC_PDL=0,HDRNET_TEX_CHUNKS=4 \
  > gpurun_out/ab_lib_stdout.txt 2>&1; echo "ab exit $?"; cat gpurun_out/ab_lib_stdout.txt | tail -16 | cut -c1-200
