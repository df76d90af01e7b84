#!/bin/bash
# y-marching point smoother: warps per SM (register cap) and the 8-colour tile kernel for comparison.
cd "$(dirname "$0")/../emg3d_b200/csrc"
while read -r a b; do
  [ -z "$a" ] && continue
  rm -f build/gs_point.o
  make -s EXTRA="-DEMG_PT_MARCH=$a -DEMG_MARCH_MINB=$b -Xptxas -v" 2>&1 | grep -A2 "gs_point_march_kernelINS_4cplx" | grep -E "registers|spill" | tr '\n' ' '
  echo
  (cd ../.. && python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('march minb = $a $b:', 'V-cycle ms', round(d['ms_per_step'],2), 'launch_ms', round(d['roofline']['launch_ms'],4), 'frac', round(d['roofline']['frac'],3))")
done <<LIST
${1:-1 8
1 10
1 12
0 8}
LIST
