#!/bin/bash
set -u
mkdir -p gpurun_out
B=tools/ubench/bin
timeout 60 $B/cabi_check 7 8 2160 3840 16 16 8 20
timeout 900 python -m pytest tests/test_slice_apply_gpu.py -x -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -6
L=hdrnet_b200/lib/libhdrnet_b200.so
AB_ROUNDS=5 timeout 300 python tools/ab_lib.py $L:7:HDRNET_ASYNC_SLAB=1 $L:7:HDRNET_ASYNC_SLAB=0 $L:7:HDRNET_ASYNC_SLAB=1,HDRNET_TEX_CHUNKS=4 > gpurun_out/ab_slab.txt 2>&1; grep -v bursts gpurun_out/ab_slab.txt | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/bench_slab.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_slab.json')); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], 'sustained', d['sustained']['frac'], d['config']['kernel']['threads'])"
