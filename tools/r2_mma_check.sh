#!/bin/bash
# correctness + time of the tensor-core gather form (variants 10 / 11) through the C-ABI, torch-free
set -u
mkdir -p gpurun_out
B=tools/ubench/bin
{
for v in 10 11; do
  timeout 30 $B/cabi_check $v 1 16 3840 16 16 8 5;   echo "cabi_check $v small exit $?"
  timeout 30 $B/cabi_check $v 2 37 1920 16 16 8 5;   echo "cabi_check $v 1920 exit $?"
  timeout 30 $B/cabi_check $v 1 64 4032 16 16 8 5;   echo "cabi_check $v 4032 exit $?"
  timeout 30 $B/cabi_check $v 3 50 640 8 5 4 5;      echo "cabi_check $v 640 gw5 gd4 exit $?"
  timeout 30 $B/cabi_check $v 1 100 3840 32 32 8 5;  echo "cabi_check $v 32x32x8 exit $?"
  timeout 30 $B/cabi_check $v 1 100 3840 16 16 16 5; echo "cabi_check $v 16x16x16 exit $?"
  timeout 60 $B/cabi_check $v 8 2160 3840 16 16 8 20; echo "cabi_check $v 8x4K exit $?"
done
timeout 60 $B/cabi_check 7 8 2160 3840 16 16 8 20; echo "cabi_check 7 8x4K exit $?"
} > gpurun_out/r2_mma_check.txt 2>&1
cat gpurun_out/r2_mma_check.txt
timeout 120 compute-sanitizer --tool memcheck $B/cabi_check 10 1 8 1920 16 16 8 1 2>&1 | tail -15
