timeout 120 python scripts/time_knn.py 10 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_nn -s 6 -c 1 -o gpurun_out/knn_v6 python scripts/time_knn.py 10 > gpurun_out/ncu_knn.log 2>&1
echo rc=$?
