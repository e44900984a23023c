timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(j['ms_per_step'],3), j['kernel_profile_ms']['k_attribute'], j['kernel_profile_ms']['k_tournament'])"; }
run evprefetch_plus_tournament
