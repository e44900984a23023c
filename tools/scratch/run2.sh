timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/r1n_bench.json 2> gpurun_out/r1n_bench.err; tail -c 900 gpurun_out/r1n_bench.json
