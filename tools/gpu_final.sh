#!/bin/bash
# One GPU call that refreshes the round's evidence set on the code as it stands: smoke(), the whole
# GPU suite, the driver's bench command (with the extra records), the reference arm once, and the
# ncu launch list of the bench command.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
set -u
R=${1:-r02_final}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/${R}_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/${R}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc $?"; tail -3 gpurun_out/${R}_bench.err
python - <<EOF
import json
d = json.load(open('gpurun_out/${R}_bench.json'))
print('bench', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'sustained', d.get('sustained', {}).get('frac'),
      'e2e', d['e2e']['value'], 'u8', d['e2e'].get('u8_image_path', {}).get('value'), 'clocks', d.get('clocks'))
print(json.dumps(d.get('extra', {}))[:3000])
EOF
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null; echo "reference rc $?"; cut -c1-400 gpurun_out/${R}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_ncu_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/${R}_ncu_bench.log 2>&1; echo "ncu rc $?"
grep -c . gpurun_out/${R}_ncu_launches.csv
