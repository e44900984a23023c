#!/bin/bash
# usage: tools/probe/sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   -- headline only (device-resident), 20 steps each
for cfg in "$@"; do
  env $cfg python bench.py --steps 20 --warmup 3 --no-configs --no-e2e --cpu-seconds 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'ms', round(d['ms_per_step'], 2), 'top', d['roofline']['kernel'], round(d['roofline']['kernel_ms_per_launch'], 2), 'parity', d['run']['parity_vs_oracle_digests'])"
done
