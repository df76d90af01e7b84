#!/bin/bash
# Prefetch-distance sweep of the line smoothers (run on the GPU box):
#   tools/tune_line.sh "0 3 6 12"
cd "$(dirname "$0")/../emg3d_b200/csrc" || exit 1
for PF in ${1:-0 3 6 12}; do
    touch gs_line.cu
    make EXTRA="-DEMG_LINE_PREFETCH=$PF" > /dev/null 2>&1 || { echo "build failed"; exit 1; }
    echo "== EMG_LINE_PREFETCH=$PF"
    (cd ../.. && python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:round(v.get('ms_nu2',v.get('ms')),3) for k,v in d['kernels'].items()})" ; python tools/run_configs.py config2:128 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('config2:128 warm cycles', d['cycle_s_warm'][1:4], 'it', d['it_mg'])")
done
touch gs_line.cu; make > /dev/null 2>&1
