#!/bin/bash
# usage: tools/probe/sweep_e2e.sh "ENV1=a" "ENV2=b" ...   -- headline + end-to-end leg only
for cfg in "$@"; do
  env $cfg python bench.py --steps 5 --warmup 3 --no-configs --cpu-seconds 0 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'ms', round(d['ms_per_step'], 2), 'e2e Mpix/s', round(d['e2e']['value']), 'parity', d['run']['parity_vs_oracle_digests'])"
done
