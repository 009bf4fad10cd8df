#!/bin/bash
set -u
mkdir -p gpurun_out
B=tools/ubench/bin
V=${1:-10}
{
  timeout 30 $B/cabi_check $V 1 16 3840 16 16 8 5;   echo "cabi_check $V small exit $?"
  timeout 30 $B/cabi_check $V 2 37 1920 16 16 8 5;   echo "cabi_check $V 1920 exit $?"
  timeout 30 $B/cabi_check $V 3 50 640 8 5 4 5;      echo "cabi_check $V 640 gw5 gd4 exit $?"
  timeout 30 $B/cabi_check $V 1 100 3840 16 16 16 5; echo "cabi_check $V 16x16x16 exit $?"
  timeout 60 $B/cabi_check $V 8 2160 3840 16 16 8 20; echo "cabi_check $V 8x4K exit $?"
} > gpurun_out/r2_mma_check.txt 2>&1
cat gpurun_out/r2_mma_check.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"slice_apply_rows_mma" -s 1 -c 1 \
      -f -o gpurun_out/r2_prof_mma $B/cabi_check $V 8 2160 3840 16 16 8 2 > gpurun_out/r2_ncu_mma.log 2>&1; echo "ncu exit $?"
