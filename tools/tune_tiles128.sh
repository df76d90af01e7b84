#!/bin/bash
# Tile shape of the tile-fused point smoother on the 128^3 level (256 tiles per launch
# with 16x8x8 tiles: fewer blocks than the GPU holds).
cd "$(dirname "$0")/../emg3d_b200/csrc"
while read -r a b c d; do
  [ -z "$a" ] && continue
  rm -f build/gs_point.o
  make -s EXTRA="-DEMG_TILE_X=$a -DEMG_TILE_Y=$b -DEMG_TILE_Z=$c -DEMG_TILE_MINB=$d" 2>&1 | grep -i error
  for n in 128 256; do
  (cd ../.. && python bench.py --size $n --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tile $a $b $c minb $d size $n:', 'V-cycle ms', round(d['ms_per_step'],3), 'launch_ms', round(d['roofline']['launch_ms'],4))")
  done
done <<LIST
${1:-16 8 8 3
8 8 8 6
16 8 4 6
16 4 4 8}
LIST
