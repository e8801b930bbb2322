timeout 120 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python scripts/time_gemm.py
