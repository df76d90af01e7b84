#!/bin/bash
# Build variants of the tile-fused point smoother and time them (run on the GPU box).
cd "$(dirname "$0")/../emg3d_b200/csrc"
for v in "16 8 8 2" "16 8 8 3" "16 8 8 4" "32 8 8 2" "32 8 4 3" "16 16 8 2" "8 8 8 4" "32 4 4 4"; do
  set -- $v
  rm -f build/gs_point.o
  make -s EXTRA="-DEMG_TILE_X=$1 -DEMG_TILE_Y=$2 -DEMG_TILE_Z=$3 -DEMG_TILE_MINB=$4 -Xptxas -v" 2>&1 | grep -A2 "gs_point_tile_kernelINS_4cplx" | grep -E "registers|spill" | tr '\n' ' '
  echo
  (cd ../.. && python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tile $v:', 'V-cycle ms', round(d['ms_per_step'],2), 'launch_ms', round(d['roofline']['launch_ms'],4), 'frac', round(d['roofline']['frac'],3))")
done
