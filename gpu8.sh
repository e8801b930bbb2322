timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo rc=$?
