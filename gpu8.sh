timeout 200 python -m pytest tests/test_cubify_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 100 python scripts/cubify_stress.py
