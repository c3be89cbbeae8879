#!/bin/bash
# A/B of environment knobs on ONE box (box-to-box variation is ~4 %): tools/ab_bench.sh "PXL_X=0" "PXL_X=1" ...
# prints value / ms_per_step of `bench.py --no-gpu-torch-baseline --no-alt` for every setting, twice (ABAB).
for rep in 1 2; do
  for kv in "$@"; do
    env $kv python bench.py --no-gpu-torch-baseline --no-alt --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('%-40s %8.2f img/s  %7.3f ms/step  e2e %8.2f  conv %.1f TF/s  wgrad %.1f TF/s' % ('$kv', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline_wgrad']['achieved']))"
  done
done
