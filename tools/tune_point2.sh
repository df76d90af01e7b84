#!/bin/bash
# Variants of the tile-fused point smoother's latency hiding (run on the GPU box):
# EMG_PT_RING = stages of the cp.async coefficient ring, EMG_PT_PFE = L2 prefetch of E.
cd "$(dirname "$0")/../emg3d_b200/csrc"
while read -r a b c; do
  [ -z "$a" ] && continue
  rm -f build/gs_point.o
  make -s EXTRA="-DEMG_PT_RING=$a -DEMG_PT_PFE=$b -DEMG_TILE_MINB=$c -Xptxas -v" 2>&1 | grep -A2 "gs_point_tile_kernelINS_4cplx" | grep -E "registers|spill" | tr '\n' ' '
  echo
  (cd ../.. && python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ring pfe minb = $a $b $c:', 'V-cycle ms', round(d['ms_per_step'],2), 'launch_ms', round(d['roofline']['launch_ms'],4), 'frac', round(d['roofline']['frac'],3))")
done <<LIST
${1:-0 0 3
0 1 3
1 0 3
2 0 3
1 1 3
1 0 2}
LIST
