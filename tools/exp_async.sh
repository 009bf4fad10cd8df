#!/bin/bash
# One gpurun call: parity of the CTA shapes, same-box A/B of knobs, launch-level timing of pre-pass + main kernel.
set -u
mkdir -p gpurun_out
L=hdrnet_b200/lib/libhdrnet_b200.so; V1=tools/abtmp/lib_v1.so
for cfg in "352 2" "224 3" "224 4"; do set -- $cfg
  HDRNET_ASYNC_THREADS=$1 HDRNET_ASYNC_OCC=$2 timeout 300 python -m pytest tests/test_slice_apply_gpu.py -q --timeout 120 -p no:cacheprovider \
    -k "tex_async or bitwise" -x 2>&1 | tail -2
done > gpurun_out/pytest_shapes.log 2>&1; cat gpurun_out/pytest_shapes.log
AB_ROUNDS=5 timeout 400 python tools/ab_lib.py $L:4:HDRNET_TEX_CHUNKS=4 $V1:7:HDRNET_TEX_CHUNKS=5 \
  $L:7:HDRNET_TEX_CHUNKS=5 $L:7:HDRNET_TEX_CHUNKS=5,HDRNET_ASYNC_LEAN=0 \
  $L:7:HDRNET_TEX_CHUNKS=4,HDRNET_ASYNC_THREADS=352 $L:7:HDRNET_TEX_CHUNKS=5,HDRNET_ASYNC_THREADS=352 $L:7:HDRNET_TEX_CHUNKS=6,HDRNET_ASYNC_THREADS=352 \
  $L:7:HDRNET_TEX_CHUNKS=4,HDRNET_ASYNC_THREADS=224 $L:7:HDRNET_TEX_CHUNKS=5,HDRNET_ASYNC_THREADS=224 $L:7:HDRNET_TEX_CHUNKS=6,HDRNET_ASYNC_THREADS=224 \
  $L:7:HDRNET_TEX_CHUNKS=4,HDRNET_ASYNC_THREADS=224,HDRNET_ASYNC_OCC=4 $L:7:HDRNET_TEX_CHUNKS=5,HDRNET_ASYNC_THREADS=224,HDRNET_ASYNC_OCC=4 $L:7:HDRNET_TEX_CHUNKS=6,HDRNET_ASYNC_THREADS=224,HDRNET_ASYNC_OCC=4 \
  > gpurun_out/ab_lib_stdout.txt 2>&1; echo "ab exit $?"; grep -v bursts gpurun_out/ab_lib_stdout.txt | tail -16
for cfg in "512 2 5" "352 2 5" "224 3 5" "224 4 5"; do set -- $cfg
HDRNET_ASYNC_THREADS=$1 HDRNET_ASYNC_OCC=$2 HDRNET_TEX_CHUNKS=$3 timeout 200 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"slice_apply_rows_async|yblend" -s 4 -c 2 --csv \
    --log-file gpurun_out/launch_$1_$2_$3.csv python tools/prof_variant.py 7 > /dev/null 2>&1; echo "ncu $cfg exit $?"
done
