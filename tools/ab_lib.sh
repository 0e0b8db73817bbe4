#!/bin/bash
# A/B of two builds of the library on the GPU box (bring-up tool): parity cases + bench kernel times for each LSR_LIB
for so in "$@"; do
  echo "=== $so"
  LSR_LIB=$PWD/$so timeout 300 python tools/gpu_check.py replica_color_mapper replica_color_tracker tum_color_mapper_dynr 2>&1 | grep -v Warn | tail -4 | cut -c1-400
  LSR_LIB=$PWD/$so python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('value',round(d['value']),'ms',round(d['ms_per_step'],3),d['kernels'])"
done
