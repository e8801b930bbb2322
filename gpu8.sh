timeout 240 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$?
grep -E "per-step|phases" gpurun_out/bench_1gpu.err | cut -c1-700
