#!/bin/bash
# ncu pass over every kernel of the library (tools/prof_all_kernels.py); only the CSV export and
# the per-kernel table travel back (the .ncu-rep of ~150 captures exceeds gpurun's 64 MiB).
set -u
mkdir -p gpurun_out /tmp/ncu
timeout 1200 ncu --set full --clock-control none -k regex:'^(conv|fc_|fuse|lowres|resize|slice|yblend|guide)' -c 240 -f -o /tmp/ncu/r02_all_kernels python tools/prof_all_kernels.py > gpurun_out/r02_all_kernels.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/ncu/r02_all_kernels.ncu-rep --page raw --csv > gpurun_out/r02_all_kernels_raw.csv 2>/dev/null
python tools/ncu_kernel_table.py gpurun_out/r02_all_kernels_raw.csv > gpurun_out/r02_all_kernels_ncu.md; head -50 gpurun_out/r02_all_kernels_ncu.md | cut -c1-260
ls -la gpurun_out /tmp/ncu
