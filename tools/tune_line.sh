#!/bin/bash
# Line smoothers: factor stream through TMA bulk copies (EMG_LINE_BULK, ring depth) vs plain loads.
cd "$(dirname "$0")/../emg3d_b200/csrc"
while read -r a b; do
  [ -z "$a" ] && continue
  rm -f build/gs_line.o
  make -s EXTRA="-DEMG_LINE_BULK=$a -DEMG_LINE_BULK_STAGES=$b -Xptxas -v" 2>&1 | grep -A2 "gs_line_color_kernelINS_4cplxELi0" | grep -E "registers|spill" | tr '\n' ' '
  echo
  (cd ../.. && python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bulk stages = $a $b:', {k: round(v.get('ms_nu2', v.get('ms')), 3) for k, v in d['kernels'].items()})")
done <<LIST
${1:-1 4
1 8
1 2
0 4}
LIST
