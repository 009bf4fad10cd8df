#!/bin/bash
# First GPU call of round 2 (one gpurun, ~3 min): everything needed to decide where the tensor-core
# forms of slice-apply stand -- written at the end of round 1, when the GPU budget was spent.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
# Build tools/ubench/bin/{tmem_paths,cabi_check} first (commands in their headers); they travel
# with the snapshot.
set -u
mkdir -p gpurun_out
B=tools/ubench/bin
# 1. torch-free: correctness + time of the issuer-warp form (7), the tensor-core form (8) and the
#    tensor-core gather form (9, never run) at a small and at the headline shape
for v in 7 8 9; do
  timeout 30 $B/cabi_check $v 1 16 3840 16 16 8 5;  echo "cabi_check $v small exit $?"
  timeout 60 $B/cabi_check $v 8 2160 3840 16 16 8 20; echo "cabi_check $v 8x4K exit $?"
done > gpurun_out/r2_cabi_check.txt 2>&1
cat gpurun_out/r2_cabi_check.txt
# 2. tensor-memory rates without the spill artefact (tcgen05.ld x32, tcgen05.st x16, MMA issue)
timeout 60 $B/tmem_paths > gpurun_out/r2_tmem_paths.txt 2>&1; echo "tmem_paths exit $?"; tail -22 gpurun_out/r2_tmem_paths.txt
# 3. the gated pytest cases of both tensor-core forms (ragged tiles, borders, gw = 3, 8 x 4K)
HDRNET_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_slice_apply_gpu.py -q --timeout 120 -p no:cacheprovider \
    -k "tensor_core" 2>&1 | tail -15 > gpurun_out/r2_pytest_tc.txt; cat gpurun_out/r2_pytest_tc.txt
# 4. one full ncu capture of each tensor-core kernel (where do the issue slots go?)
for v in 8 9; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"slice_apply_rows_tc" -s 2 -c 1 \
      -f -o gpurun_out/r2_prof_tc$v python tools/prof_variant.py $v > gpurun_out/r2_ncu_tc$v.log 2>&1; echo "ncu $v exit $?"
done
# 5. same-box A/B of the issuer-warp form's opt-in knobs that have not been measured yet:
#    texture fetches one pixel ahead (HDRNET_ASYNC_PIPE), in the 64- and the 80-register shape
L=hdrnet_b200/lib/libhdrnet_b200.so
HDRNET_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_slice_apply_gpu.py -q --timeout 120 -p no:cacheprovider \
    -k "pipelined" 2>&1 | tail -5
AB_ROUNDS=5 timeout 300 python tools/ab_lib.py $L:7 $L:7:HDRNET_ASYNC_PIPE=1 $L:7:HDRNET_ASYNC_THREADS=352 \
  $L:7:HDRNET_ASYNC_THREADS=352,HDRNET_ASYNC_PIPE=1 $L:7:HDRNET_TEX_CHUNKS=4,HDRNET_ASYNC_THREADS=352,HDRNET_ASYNC_PIPE=1 \
  $L:7:HDRNET_TEX_CHUNKS=6,HDRNET_ASYNC_THREADS=352 > gpurun_out/r2_ab_pipe.txt 2>&1; grep -v bursts gpurun_out/r2_ab_pipe.txt | tail -8
# 6. fused-guide (model path) forms under the issuer-warp control flow (HDRNET_FUSED_ASYNC=1, never run)
HDRNET_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_px_gpu.py -q --timeout 120 -p no:cacheprovider \
    -k "issuer_warp" 2>&1 | tail -5
timeout 200 python tools/ab_fused.py > gpurun_out/r2_ab_fused.txt 2>&1; tail -20 gpurun_out/r2_ab_fused.txt
