set -x
python __graft_entry__.py --smoke 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo rc=$?
tail -5 gpurun_out/bench1.err
