timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo rc=$?
