#!/bin/bash
cd "$(dirname "$0")/../emg3d_b200/csrc"
for v in "8 2 16" "8 3 16" "8 4 16" "4 4 16" "4 6 16" "8 2 64" "8 3 257" "4 4 257" "16 1 16"; do
  set -- $v
  rm -f build/amat.o
  make -s EXTRA="-DEMG_RZ_BY=$1 -DEMG_RZ_MINB=$2 -DEMG_RZ_KZ=$3 -Xptxas -v" 2>&1 | grep -A2 "residual_zmarch_kernelINS_4cplx" | grep -E "registers|spill" | tr '\n' ' '
  echo
  (cd ../.. && python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('BY MINB KZ = $v:', 'V-cycle ms', round(d['ms_per_step'],2), 'residual ms', round(d['kernels']['residual']['ms'],4))")
done
