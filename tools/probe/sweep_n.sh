#!/bin/bash
# usage: tools/probe/sweep_n.sh IMAGES "ENV1=a" "ENV2=b" ...  -- headline kernel path with IMAGES images on this GPU (what one rank of an N-GPU run sees)
n=$1; shift
for cfg in "$@"; do
  env $cfg python bench.py --images $n --steps 20 --warmup 3 --no-configs --no-e2e --cpu-seconds 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('images $n', '$cfg', 'ms', round(d['ms_per_step'], 2), 'top', d['roofline']['kernel'], round(d['roofline']['kernel_ms_per_launch'], 2))"
done
