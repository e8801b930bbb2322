timeout 120 python scripts/time_gemm.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 3 -c 1 -o gpurun_out/gemm_v1 python scripts/time_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo rc=$?
