#!/bin/bash
# Round 2, late: model tests, chain timing by batch, ncu --set full of the current headline kernel
# (slab-warp issuer-warp form) and of the coefficient-network chain; CSV exports only.
set -u
mkdir -p gpurun_out /tmp/ncu
python -m pytest tests/test_models.py -x -q -m gpu 2>&1 | tail -3
python tools/time_cnn.py
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'slice_apply_rows_async' -s 2 -c 1 -f -o /tmp/ncu/headline python bench.py --steps 2 --warmup 3 --no-extra > /dev/null 2>&1; echo "ncu headline exit $?"
ncu -i /tmp/ncu/headline.ncu-rep --page raw --csv > gpurun_out/r02_slabwarp_raw.csv 2>/dev/null
ncu -i /tmp/ncu/headline.ncu-rep --page details --csv > gpurun_out/r02_slabwarp_details.csv 2>/dev/null
ncu -i /tmp/ncu/headline.ncu-rep --page source --csv > gpurun_out/r02_slabwarp_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:'^(conv|fc_|fuse)' -s 24 -c 8 -f -o /tmp/ncu/chain python tools/prof_cnn.py > /dev/null 2>&1; echo "ncu chain exit $?"
ncu -i /tmp/ncu/chain.ncu-rep --page raw --csv > gpurun_out/r02_cnn_chain_raw.csv 2>/dev/null
python tools/ncu_kernel_table.py gpurun_out/r02_cnn_chain_raw.csv | cut -c1-220
python tools/ncu_kernel_table.py gpurun_out/r02_slabwarp_raw.csv | cut -c1-220
python tools/time_any.py
python -m pytest tests/test_slice_apply_gpu.py -x -q -m gpu -k "any or shape or offset" 2>&1 | tail -3
ls -la gpurun_out | head -20
