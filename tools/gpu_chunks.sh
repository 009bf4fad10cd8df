#!/bin/bash
# Texture / LSU split of the corner chunks per grid shape (issuer-warp kernel, tools/ab_lib.py).
set -u
mkdir -p gpurun_out
L=hdrnet_b200/lib/libhdrnet_b200.so
for G in 32,32,16 16,16,16 32,32,8; do
  AB_GRID=$G AB_ROUNDS=3 timeout 200 python tools/ab_lib.py $L:0:HDRNET_TEX_CHUNKS=5 $L:0:HDRNET_TEX_CHUNKS=4 $L:0:HDRNET_TEX_CHUNKS=3 $L:0:HDRNET_TEX_CHUNKS=2 $L:0:HDRNET_TEX_CHUNKS=1 $L:2 $L:4 > /dev/null 2>gpurun_out/ab_err.txt
  echo "== grid $G"; grep -v bursts gpurun_out/ab_lib.txt | cut -c30-; cp gpurun_out/ab_lib.txt gpurun_out/ab_chunks_low_${G//,/x}.txt
done
