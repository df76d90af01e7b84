#!/bin/bash
# residual_zmarch_kernel: rows per block (BY), blocks per SM (MINB), planes per block (KZ)
cd "$(dirname "$0")/../emg3d_b200/csrc"
while read -r a b c; do
  [ -z "$a" ] && continue
  rm -f build/amat.o
  make -s EXTRA="-DEMG_RZ_BY=$a -DEMG_RZ_MINB=$b -DEMG_RZ_KZ=$c -Xptxas -v" 2>&1 | grep -A2 "residual_zmarch_kernelINS_4cplx" | grep -E "registers|spill" | tr '\n' ' '
  echo
  (cd ../.. && python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('BY MINB KZ = $a $b $c:', 'V-cycle ms', round(d['ms_per_step'],2), 'residual ms', round(d['kernels']['residual']['ms'],4))")
done <<LIST
${1:-8 4 16
8 3 16
8 2 16
4 6 16
8 3 32}
LIST
